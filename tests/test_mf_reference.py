"""CPU: the matrix-free statement of the PDHG iteration (tests/mf_reference.py, what csrc/pdhg_mf.cu implements)
IS the CSR iteration on the oracle's strengthened matrix, and it converges to the HiGHS LP optimum."""
import numpy as np
import pytest

from helpers import arrays_of, float_payload
from mf_reference import Generic, MatrixFree, MatrixFreeN, node_cut_bigm, pc_diag, run_fixed, solve, strengthened
from neptune_mip_b200 import synth
from oracle import mip as omip

CASES = [("C1", lambda: synth.test_py_payload()), ("r8x4", lambda: synth.random_payload(8, 4, 0, node_cores=30)),
         ("r12x5", lambda: synth.random_payload(12, 5, 1, node_cores=25)), ("r7x3float", lambda: float_payload(7, 3, 5))]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_matrix_free_iteration_equals_csr_iteration(name, make):
    a = arrays_of(make())
    m = strengthened(a)
    T, S = pc_diag(m)
    g, mf = Generic(m, T, S), MatrixFree(a)
    assert abs(g.omega - mf.omega) <= 1e-12 * g.omega          # same initial primal weight
    assert abs(g.nb - mf.nb) <= 1e-12 * (1 + g.nb) and abs(g.nc - mf.nc) <= 1e-12 * (1 + g.nc)
    for _ in range(150):
        g.step(); mf.step()
    x, y = mf.pack()
    assert np.abs(x - g.x).max() <= 1e-10 * (1 + np.abs(g.x).max())
    assert np.abs(y - g.y).max() <= 1e-10 * (1 + np.abs(g.y).max())
    kg, km = g.kkt(g.x, g.y), mf.kkt(mf.state())
    for u, v in zip(kg, km):                                    # primal/dual residuals, objectives in closed form
        assert abs(u - v) <= 1e-9 * (1 + abs(u))


def test_closed_form_step_sizes_are_the_pock_chambolle_sums():
    a = arrays_of(synth.random_payload(8, 4, 2, node_cores=30))
    m = strengthened(a)
    T, S = pc_diag(m)
    mf = MatrixFree(a)
    N, F = a["N"], a["F"]; X, C = F * N * N, F * N
    assert np.allclose(T[:X], mf.Tx.reshape(-1), rtol=1e-14) and np.allclose(T[X:X + C], mf.Tc.reshape(-1), rtol=1e-14)
    rows = np.concatenate([np.stack([np.ones(C), np.full(C, mf.S1)], 1).reshape(-1), np.full(N, mf.S2), np.full(C, mf.S3),
                           mf.S4, np.full(X, mf.SS)])
    assert np.allclose(S, rows, rtol=1e-14)


def test_matrix_free_pdhg_reaches_the_highs_lp_optimum():
    a = arrays_of(synth.random_payload(8, 4, 1, node_cores=30))
    lp = omip.solve_model(strengthened(a), relax=True)
    out = solve(MatrixFree(a), max_iters=40000, check=64, eps=1e-6)
    assert out["converged"]
    assert abs(out["primal"] - lp["objective"]) <= 1e-4 * (1 + abs(lp["objective"]))
    assert abs(out["dual"] - lp["objective"]) <= 1e-4 * (1 + abs(lp["objective"]))


def test_run_fixed_reports_the_better_candidate():
    a = arrays_of(synth.random_payload(8, 4, 0, node_cores=30))
    x, y, info = run_fixed(a, 64)
    assert info["pick"] == (1 if info["kkt"][1] < info["kkt"][0] else 0)
    N, F = a["N"], a["F"]
    assert x.shape == (F * N * N + F * N,) and y.shape == (3 * F * N + 2 * N + F * N * N,)
    assert np.all(y[0:2 * F * N:2] == 0.0)                      # free C1a rows carry no multiplier


@pytest.mark.parametrize("seed", range(8))
def test_lp_value_sweep_against_highs(seed):
    """random shapes (odd / even N, slack and binding CPU rows): the restarted, averaged iteration in its
    matrix-free form converges to the HiGHS optimum of the oracle's strengthened matrix, primal and dual side"""
    rng = np.random.default_rng(100 + seed)
    N, F = int(rng.integers(3, 14)), int(rng.integers(1, 6))
    cores = int(rng.choice([8, 15, 30, 200]))
    a = arrays_of(synth.random_payload(N, F, seed, node_cores=cores))
    lp = omip.solve_model(strengthened(a), relax=True)
    if not lp["optimal"]:
        pytest.skip("relaxation infeasible for this draw")
    out = solve(MatrixFree(a), max_iters=80000, check=64, eps=1e-6)
    assert out["converged"], (N, F, cores, out)
    tol = 1e-4 * (1 + abs(lp["objective"]))
    assert abs(out["primal"] - lp["objective"]) <= tol and abs(out["dual"] - lp["objective"]) <= tol, (N, F, cores, out, lp["objective"])


@pytest.mark.parametrize("cut", [False, True], ids=["M=1e6", "node-cut"])
@pytest.mark.parametrize("kind", ["min_util", "min_delay_util"])
@pytest.mark.parametrize("name,make", CASES[:3], ids=[c[0] for c in CASES[:3]])
def test_matrix_free_iteration_with_node_variables_equals_csr_iteration(name, make, kind, cut):
    """the models with n[j] columns and C5a / C5b / C6 rows (reference neptune_step1.py:38-77): closed form == CSR,
    with the reference's M = 10^6 and with the per-node M of the node cut"""
    a = arrays_of(make())
    bigm = node_cut_bigm(a) if cut else None
    m = strengthened(a, kind, 0.5, bigm)
    T, S = pc_diag(m)
    g, mf = Generic(m, T, S), MatrixFreeN(a, kind, 0.5, bigm)
    assert abs(g.omega - mf.omega) <= 1e-12 * g.omega
    assert abs(g.nb - mf.nb) <= 1e-12 * (1 + g.nb) and abs(g.nc - mf.nc) <= 1e-12 * (1 + g.nc)
    for _ in range(150):
        g.step(); mf.step()
    x, y = mf.pack()
    assert np.abs(x - g.x).max() <= 1e-10 * (1 + np.abs(g.x).max())
    assert np.abs(y - g.y).max() <= 1e-10 * (1 + np.abs(g.y).max())
    for u, v in zip(g.kkt(g.x, g.y), mf.kkt(mf.state())):
        assert abs(u - v) <= 1e-9 * (1 + abs(u))


def test_node_cut_is_valid_and_tightens_the_relaxation():
    """M_j = pods node j can hold: the MIP optimum is unchanged (every integer point satisfies the tighter row), the LP
    bound moves from ~0 towards it"""
    a = arrays_of(synth.random_payload(12, 5, 1, node_cores=25))
    for kind in ("min_util", "min_delay_util"):
        lp_ref = omip.solve_model(strengthened(a, kind, 0.5), relax=True)["objective"]
        lp_cut = omip.solve_model(strengthened(a, kind, 0.5, node_cut_bigm(a)), relax=True)["objective"]
        mip = omip.solve_step1(a, kind, 0.5, time_limit=60)
        assert lp_ref <= lp_cut + 1e-9 <= mip["objective"] + 1e-6
        assert lp_cut >= 0.4 * mip["objective"] and lp_ref <= 0.1 * mip["objective"]
