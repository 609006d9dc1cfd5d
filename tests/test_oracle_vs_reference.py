"""Oracle vs the UNMODIFIED reference code (imported read-only through oracle.refshim).
Skipped where /root/reference does not exist (the GPU box)."""
import contextlib
import io
import sys

import numpy as np
import pytest

from helpers import data_of, small_payloads
from oracle import efttc as oefttc, model as omodel
from oracle.refshim import load_reference as L

pytestmark = pytest.mark.skipif(not L.reference_available(), reason="reference tree not mounted")

STEP1 = {"min_delay": "NeptuneStep1CPUMinDelay", "min_util": "NeptuneStep1CPUMinUtilization",
         "min_delay_util": "NeptuneStep1CPUMinDelayAndUtilization"}
EFTTC = {"min_delay": "EfttcStep1CPUMinDelay", "min_util": "EfttcStep1CPUMinUtilization",
         "min_delay_util": "EfttcStep1CPUMinDelayAndUtilization"}


def _ref_data(core, payload):
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        return core.data_to_solver_input(payload, with_db=False, workload_coeff=payload.get("workload_coeff", 1))


@pytest.mark.parametrize("name,payload,alpha", small_payloads(), ids=lambda v: v if isinstance(v, str) else None)
def test_input_adapter_equals_reference(name, payload, alpha):
    core = L.load_reference()
    rd, md = _ref_data(core, payload), data_of(payload)
    for k in ["node_memory_matrix", "function_memory_matrix", "node_delay_matrix", "workload_matrix",
              "max_delay_matrix", "node_cores_matrix", "cores_matrix", "old_allocations_matrix",
              "core_per_req_matrix", "node_costs"]:
        assert np.array_equal(np.asarray(getattr(rd, k)), np.asarray(getattr(md, k))), k
    assert rd.node_budget == md.node_budget and rd.nodes == md.nodes and rd.functions == md.functions


@pytest.mark.parametrize("name,payload,alpha", small_payloads(), ids=lambda v: v if isinstance(v, str) else None)
@pytest.mark.parametrize("kind", list(STEP1))
def test_model_equals_reference_built_matrix(name, payload, alpha, kind):
    core = L.load_reference()
    solvers = sys.modules["core.solvers"]
    data = _ref_data(core, payload)
    with contextlib.redirect_stdout(io.StringIO()):
        kw = dict(verbose=False)
        if kind == "min_delay_util":
            kw["alpha"] = alpha
        s = getattr(solvers, STEP1[kind])(**kw)
        s.load_data(data)
        s.init_objective()
    A, lo, hi, obj, lb, ub, integ = s.solver.export()
    m = omodel.build_step1(omodel.arrays_from_data(data), kind, alpha)
    assert np.array_equal(A.indptr, m["A"].indptr) and np.array_equal(A.indices, m["A"].indices)
    assert np.array_equal(A.data, m["A"].data)
    for k, v in dict(lo=lo, hi=hi, obj=obj, lb=lb, ub=ub, integ=integ).items():
        assert np.array_equal(v, m[k]), k


@pytest.mark.parametrize("name,payload,alpha", small_payloads(), ids=lambda v: v if isinstance(v, str) else None)
@pytest.mark.parametrize("kind", list(EFTTC))
@pytest.mark.parametrize("strict", [True, False])
def test_efttc_equals_reference(name, payload, alpha, kind, strict):
    core = L.load_reference()
    solvers = sys.modules["core.solvers"]
    L.enable_efttc_discard_patch(not strict)
    try:
        ref_err = None
        with contextlib.redirect_stdout(io.StringIO()):
            data = _ref_data(core, payload)
            kw = dict(verbose=False)
            if kind == "min_delay_util":
                kw["alpha"] = alpha
            s = getattr(solvers, EFTTC[kind])(**kw)
            s.load_data(data)
            try:
                s.solve()
                rx, rc = s.results()
                rs = s.score()
            except KeyError as e:
                ref_err = e
    finally:
        L.enable_efttc_discard_patch(False)
    a = omodel.arrays_from_data(data)
    if ref_err is not None:
        with pytest.raises(KeyError):
            oefttc.solve(a, kind, alpha, strict=strict)
        return
    res = oefttc.solve(a, kind, alpha, strict=strict)
    assert np.array_equal(rc, res.c) and np.array_equal(rx, res.x)
    got = oefttc.score(a, kind, alpha, res)
    assert got == rs or np.isclose(got, rs, rtol=1e-15, atol=0)
