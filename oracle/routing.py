"""TEST INFRASTRUCTURE ONLY -- exact routing LP for a fixed placement (HiGHS).

With c fixed, the reference's step-1 model (`constraints_step1.py:47-65`, objective
`objectives.py:4-11`) is an LP in x over the open pods only:
    min sum d[i,j] w[f,i] x[i,f,j]   s.t.  sum_j x[i,f,j] = 1,   sum_{f,i} w[f,i] r[f,j] x[i,f,j] <= K_j,  x >= 0,
and, with `c1b=True`, the lower half of the c <-> x link for every open pod (`constraints_step1.py:12-15`):
    sum_i x[i,f,j] >= 1 - eps.
Used to check `neptune_route_lp` (csrc/route_lp.cu) and `neptune_route_capacitated` (csrc/route_cap.cuh)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


def lp_routing(a, c, c1b=False):
    N, F = a["N"], a["F"]
    cb = np.asarray(c) > 0
    cols = [(f, i, j) for f in range(F) for i in range(N) for j in np.flatnonzero(cb[f])]
    if any(not cb[f].any() for f in range(F)):
        return None
    fi = np.array([f * N + i for f, i, j in cols])
    jj = np.array([j for f, i, j in cols])
    ff = np.array([f for f, i, j in cols])
    ii = np.array([i for f, i, j in cols])
    cost = a["d"][ii, jj] * a["w"][ff, ii]
    A_eq = sp.csr_matrix((np.ones(len(cols)), (fi, np.arange(len(cols)))), shape=(F * N, len(cols)))
    A_ub = sp.csr_matrix((a["w"][ff, ii] * a["r"][ff, jj], (jj, np.arange(len(cols)))), shape=(N, len(cols)))
    b_ub = a["Kj"]
    if c1b:
        pods = [(f, j) for f in range(F) for j in np.flatnonzero(cb[f])]
        pid = {p: k for k, p in enumerate(pods)}
        rows = np.array([pid[(f, j)] for f, i, j in cols])
        A_c = sp.csr_matrix((-np.ones(len(cols)), (rows, np.arange(len(cols)))), shape=(len(pods), len(cols)))
        A_ub = sp.vstack([A_ub, A_c]).tocsr()
        b_ub = np.r_[a["Kj"], -(1.0 - 1e-6) * np.ones(len(pods))]
    res = linprog(cost, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=np.ones(F * N), bounds=(0, None), method="highs")
    if res.status != 0:
        return None
    x = np.zeros((N, F, N))
    x[ii, ff, jj] = res.x
    return float(res.fun), x
