"""(b) SpMV / SpMV^T against scipy, PDHG optimum against HiGHS on the same matrix."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import arrays_of, cuda_batch
from neptune_mip_b200 import synth
from oracle import mip as omip
from oracle import model as omodel

pytestmark = pytest.mark.gpu


def _csr(mdl, b=0):
    return sp.csr_matrix((mdl.val[b].cpu().numpy(), mdl.col_idx.cpu().numpy(), mdl.row_ptr.cpu().numpy()),
                         shape=(mdl.rows, mdl.cols))


@pytest.mark.parametrize("shape", [(8, 4), (20, 5), (50, 10), (120, 30)])
@pytest.mark.parametrize("kind", ["min_delay", "min_delay_util"])
def test_spmv_matches_scipy(shape, kind):
    import torch
    from neptune_mip_b200 import device
    payloads = [synth.random_payload(shape[0], shape[1], s, node_cores=100) for s in range(2)]
    mdl = device.assemble(cuda_batch(payloads), kind, 0.5)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((mdl.B, mdl.cols), dtype=torch.float64, device="cuda", generator=g)
    y = torch.rand((mdl.B, mdl.rows), dtype=torch.float64, device="cuda", generator=g) - 0.5
    ax = device.spmv(mdl, x).cpu().numpy()
    aty = device.spmv_t(mdl, y).cpu().numpy()
    for b in range(mdl.B):
        A = _csr(mdl, b)
        ref = A @ x[b].cpu().numpy()
        assert np.allclose(ax[b], ref, rtol=1e-12, atol=1e-9 * np.abs(ref).max())
        reft = A.T @ y[b].cpu().numpy()
        assert np.allclose(aty[b], reft, rtol=1e-12, atol=1e-9 * np.abs(reft).max())


def test_spmv_long_rows():
    """C4-type rows (F*N non-zeros) take the block-per-row path when F*N > 2048."""
    import torch
    from neptune_mip_b200 import device
    p = synth.random_payload(100, 25, 0, node_cores=100)
    mdl = device.assemble(cuda_batch([p]), "min_delay")
    x = torch.rand((1, mdl.cols), dtype=torch.float64, device="cuda")
    ref = _csr(mdl) @ x[0].cpu().numpy()
    got = device.spmv(mdl, x).cpu().numpy()[0]
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-9 * np.abs(ref).max())


@pytest.mark.parametrize("shape,cores,kind,flags", [
    ((8, 4), 30, "min_delay", 0), ((8, 4), 12, "min_delay", 0), ((12, 5), 25, "min_delay", 1),
    ((8, 4), 30, "min_delay_util", 1), ((10, 3), 20, "min_delay", 1)])
def test_pdhg_reaches_highs_lp_optimum(shape, cores, kind, flags):
    """LP relaxation value: PDHG vs HiGHS on the SAME matrix (as written, and strengthened).
    Tolerance: |obj_pdhg - obj_highs| <= 1e-4 * (1 + |obj_highs|)  (BASELINE.json: 1e-4 relative)."""
    from neptune_mip_b200 import device
    p = synth.random_payload(shape[0], shape[1], 1, node_cores=cores)
    mdl = device.assemble(cuda_batch([p]), kind, 0.5, flags=flags)
    ref = dict(A=_csr(mdl), lo=mdl.lo[0].cpu().numpy(), hi=mdl.hi[0].cpu().numpy(),
               obj=mdl.obj[0].cpu().numpy(), lb=mdl.col_lb[0].cpu().numpy(), ub=mdl.col_ub[0].cpu().numpy(),
               integ=mdl.col_int.cpu().numpy())
    lp = omip.solve_model(ref, relax=True)
    assert lp["optimal"]
    x, y, res = device.pdhg_solve(mdl, max_iters=60000, eps_rel=1e-6, eps_abs=1e-9)
    r = res[0]
    tol = 1e-4 * (1.0 + abs(lp["objective"]))
    assert abs(r["primal_obj"] - lp["objective"]) <= tol, (r, lp["objective"])
    assert abs(r["dual_obj"] - lp["objective"]) <= 10 * tol, (r, lp["objective"])
    # primal feasibility of the returned iterate, in the original (unscaled) space
    ax = _csr(mdl) @ x[0].cpu().numpy()
    viol = np.maximum(ref["lo"] - ax, 0) + np.maximum(ax - ref["hi"], 0)
    assert viol.max() <= 1e-4 * (1 + np.abs(ref["hi"][np.isfinite(ref["hi"])]).max())
