"""(d) GPU EFTTC must return the SAME placement as the reference's greedy (via the oracle restatement,
which tests/test_oracle_vs_reference.py pins to the unmodified reference and to its Alibaba goldens)."""
import json
import os

import numpy as np
import pytest

from helpers import KIND_NAMES, arrays_of, cuda_batch, float_payload, small_payloads
from neptune_mip_b200 import synth
from oracle import efttc as oefttc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _run_gpu(payloads, kind, alpha):
    from neptune_mip_b200 import device
    c, n, info = device.efttc(cuda_batch(payloads), kind, alpha)
    return c.cpu().numpy(), n.cpu().numpy(), info.cpu().numpy()


@pytest.mark.parametrize("name,payload,alpha", small_payloads(), ids=lambda v: v if isinstance(v, str) else None)
@pytest.mark.parametrize("kind", KIND_NAMES)
def test_efttc_equals_oracle(name, payload, alpha, kind):
    a = arrays_of(payload)
    ref = oefttc.solve(a, kind, alpha, strict=False)
    c, n, info = _run_gpu([payload], kind, alpha)
    assert np.array_equal(c[0], ref.c.astype(np.uint8)), (info[0], ref.iterations)
    assert np.array_equal(n[0], ref.n.astype(np.uint8))
    assert info[0, 0] == ref.iterations
    assert bool(info[0, 2]) == ref.would_raise


@pytest.mark.parametrize("kind", KIND_NAMES)
def test_efttc_batched_sweep(kind):
    """C5-shaped sweep (20x5), 48 seeds in one launch."""
    payloads = [synth.random_payload(20, 5, s, node_cores=100) for s in range(48)]
    c, n, info = _run_gpu(payloads, kind, 0.5)
    for b, p in enumerate(payloads):
        ref = oefttc.solve(arrays_of(p), kind, 0.5, strict=False)
        assert np.array_equal(c[b], ref.c.astype(np.uint8)), b
        assert bool(info[b, 2]) == ref.would_raise, b


@pytest.mark.parametrize("kind", KIND_NAMES)
def test_efttc_c2(kind):
    p = synth.config_payload("C2")
    ref = oefttc.solve(arrays_of(p), kind, 0.5, strict=False)
    c, n, info = _run_gpu([p], kind, 0.5)
    assert np.array_equal(c[0], ref.c.astype(np.uint8))


@pytest.mark.parametrize("solver", ["EfttcMinDelay", "EfttcMinUtilization", "EfttcMinDelayAndUtilization"])
def test_efttc_alibaba_golden(solver):
    """The reference's shipped 100x25 outputs (testing/alibaba/alibaba_test/output_Efttc*_case0.json)."""
    with open(os.path.join(GOLD, "alibaba_case0.json")) as fh:
        gold = json.load(fh)
    payload = gold["input"]
    kind = {"EfttcMinDelay": "min_delay", "EfttcMinUtilization": "min_util",
            "EfttcMinDelayAndUtilization": "min_delay_util"}[solver]
    alpha = gold["outputs"][solver]["alpha"]
    c, n, info = _run_gpu([payload], kind, alpha)
    want = gold["outputs"][solver]["cpu_allocations"]
    got = {}
    for f, fname in enumerate(payload["function_names"]):
        for j, node in enumerate(payload["node_names"]):
            if c[0, f, j]:
                got.setdefault(fname, {})[node] = True
    assert got == want


def test_efttc_host_buffer_entry_point():
    """`neptune_efttc_host` (host pointers in, host pointers out) == device path + checkers."""
    import numpy as np
    from helpers import data_of
    from neptune_mip_b200 import device
    from oracle import checkers
    payloads = [synth.random_payload(12, 5, s, node_cores=25) for s in range(6)]
    host = device.InstanceBatch.host_arrays([data_of(p) for p in payloads])
    for kind in KIND_NAMES:
        c, n, info, flags, scores = device.efttc_host(host, kind, 0.3)
        cd, nd, infod = _run_gpu(payloads, kind, 0.3)
        assert np.array_equal(c, cd) and np.array_equal(n, nd) and np.array_equal(info, infod)
        for b, p in enumerate(payloads):
            a = arrays_of(p)
            ref = oefttc.solve(a, kind, 0.3, strict=False)
            assert flags[b] == checkers.flags_to_mask(checkers.check_all(a, ref.x, ref.c, ref.n))
            assert np.isclose(scores[b, 0], checkers.score_delay(a, ref.x), rtol=1e-12, atol=1e-12)
            assert scores[b, 1] == checkers.score_util(a, ref.n)
