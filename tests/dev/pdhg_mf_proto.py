"""Development driver for the matrix-free PDHG reference (tests/mf_reference.py): iterate equality with the
CSR iteration, LP optimum against HiGHS, convergence.   python tests/dev/pdhg_mf_proto.py [N F cores]"""
import sys
import time

sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np

from mf_reference import Generic, MatrixFree, pc_diag, solve, strengthened  # noqa: F401

if __name__ == "__main__":
    from helpers import arrays_of
    from neptune_mip_b200 import synth
    from oracle import mip as omip
    N, F = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (12, 5)
    cores = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    a = arrays_of(synth.random_payload(N, F, 1, node_cores=cores))
    m = strengthened(a)
    # (b) iterate equivalence with the generic CSR PDHG, same preconditioner
    T, S = pc_diag(m)
    g, mf = Generic(m, T, S), MatrixFree(a)
    print("omega0", g.omega, mf.omega, "norms", g.nb, mf.nb, g.nc, mf.nc)
    for k in range(200):
        g.step(); mf.step()
    xm, ym = mf.pack()
    print("iterate diff after 200 steps: x", np.abs(xm - g.x).max(), "y", np.abs(ym - g.y).max())
    print("kkt generic", g.kkt(g.x, g.y)); print("kkt mf     ", mf.kkt(mf.state()))
    # (c) LP optimum
    lp = omip.solve_model(m, relax=True)
    print("highs", lp["objective"])
    t = time.time()
    out = solve(MatrixFree(a), verbose=True)
    print(out, time.time() - t)
