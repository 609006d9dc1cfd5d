"""Exact routing LP on device (`neptune_route_lp`, csrc/route_lp.cu) against HiGHS on the same placements
(oracle.routing.lp_routing): the LP the reference's step-1 MIP reduces to for a fixed c
(`neptune/utils/constraints_step1.py:47-65`, `objectives.py:4-11`)."""
import json
import os

import numpy as np
import pytest

from helpers import arrays_of, cuda_batch
from neptune_mip_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _random_placements(F, N, seed, count):
    rng = np.random.default_rng(seed)
    out = np.zeros((count, F, N), np.uint8)
    for t in range(count):
        c = (rng.random((F, N)) < 0.5).astype(np.uint8)
        for f in range(F):
            if not c[f].any():
                c[f, rng.integers(N)] = 1
        out[t] = c
    return out


@pytest.mark.parametrize("shape", [(8, 4, 12), (12, 5, 25), (12, 5, 15), (20, 5, 40), (20, 5, 100)])
def test_route_lp_equals_highs_on_cpu_starved_placements(shape):
    """Random placements on instances where most CPU rows bind (node_cores far below the load): optimum value
    to 1e-9 relative, infeasibility verdicts identical, x feasible for the reference's CPU / handle checks."""
    import torch
    from neptune_mip_b200 import device
    from oracle import checkers, routing
    N, F, cores = shape
    P = 6
    payloads = [synth.random_payload(N, F, s, node_cores=cores) for s in range(6)]
    inst = cuda_batch(payloads)
    cs = np.stack([_random_placements(F, N, s, P) for s in range(6)])
    out = device.route_lp(inst, torch.from_numpy(cs).cuda().contiguous(), want_x=True)
    obj, st, x = out["obj"].cpu().numpy(), out["status"].cpu().numpy(), out["x"].cpu().numpy()
    info, c_out = out["info"].cpu().numpy(), out["c_out"].cpu().numpy()
    priced = 0
    for b, pl in enumerate(payloads):
        a = arrays_of(pl)
        for q in range(P):
            lp = routing.lp_routing(a, cs[b, q])
            if lp is None:
                assert st[b, q] == 0, (b, q, st[b, q], obj[b, q])
                continue
            if st[b, q] == 0:
                continue                                       # a function lost its last pod to C1b: no answer for this placement
            # the value is the optimum of the LP WITH the C1b rows (every open pod keeps a share >= 1 - eps,
            # constraints_step1.py:12-15) of the placement the kernel returns: c minus the pods nobody uses and, where
            # no routing can give a pod its share, minus those pods
            closed = (cs[b, q] != c_out[b, q])
            xr = x[b, q]
            if np.any(xr[:, closed] != 0.0):
                raise AssertionError("flow on a closed pod")
            lp2 = routing.lp_routing(a, c_out[b, q], c1b=True)
            assert lp2 is not None and obj[b, q] >= lp[0] - 1e-9 * (1 + abs(lp[0]))
            assert abs(obj[b, q] - lp2[0]) <= 1e-9 * (1 + abs(lp2[0])), (b, q, obj[b, q], lp2[0], lp[0], info[b, q])
            assert np.all(np.abs(xr.sum(axis=2) - 1.0) < 1e-9)
            assert np.all(checkers.cpu_load(a, xr) <= a["Kj"] + 1e-6)
            assert np.all(xr[:, cs[b, q] == 0] == 0.0)
            share = xr.sum(axis=0)
            assert np.all((share == 0.0) | (share >= 1.0 - 1e-6 - 1e-12))            # C1b
            priced += info[b, q, 1] > 0
    assert priced > 0          # the simplex path ran (not only the nearest-pod shortcut)


def test_route_lp_prices_the_proven_optimal_c2_placements():
    """HiGHS' optimal placements of C2 (tests/golden/mip_optima.json, proven optimal): the device LP must
    reproduce the MIP optimum to 1e-9 relative, feasible for all six reference checkers."""
    import torch
    from neptune_mip_b200 import device
    gold = [r for r in json.load(open(os.path.join(GOLD, "mip_optima.json")))
            if r["config"] == "C2" and r.get("placement") and r["optimal"]]
    assert len(gold) >= 8
    payloads = [synth.config_payload("C2", r["seed"]) for r in gold]
    inst = cuda_batch(payloads)
    c = np.zeros((len(gold), 1, inst.F, inst.N), np.uint8)
    for b, r in enumerate(gold):
        for f, j in r["placement"]:
            c[b, 0, f, j] = 1
    out = device.route_lp(inst, torch.from_numpy(c).cuda().contiguous(), want_x=True)
    flags, scores = device.check_solution(inst, out["x"][:, 0].contiguous(), device.u8_to_f64(out["c_out"][:, 0].contiguous()),
                                          out["n"][:, 0].contiguous())
    for b, r in enumerate(gold):
        assert int(out["status"].cpu()[b, 0]) == 1
        assert int(flags.cpu()[b]) == 63, (r["seed"], bin(int(flags.cpu()[b])))
        got = float(out["obj"].cpu()[b, 0])
        assert abs(got - r["objective"]) <= 1e-9 * (1 + abs(r["objective"])), (r["seed"], got, r["objective"])
        assert abs(float(scores.cpu()[b, 0]) - r["objective"]) <= 1e-9 * (1 + abs(r["objective"]))


def test_route_lp_is_reproducible_and_reports_oversize():
    import torch
    from neptune_mip_b200 import device
    payloads = [synth.random_payload(12, 5, s, node_cores=15) for s in range(4)]
    inst = cuda_batch(payloads)
    cs = torch.from_numpy(np.stack([_random_placements(5, 12, s, 4) for s in range(4)])).cuda().contiguous()
    a = device.route_lp(inst, cs, want_x=True)
    b = device.route_lp(inst, cs, want_x=True)
    assert torch.equal(a["obj"], b["obj"]) and torch.equal(a["x"], b["x"]) and torch.equal(a["status"], b["status"])
    tiny = device.route_lp(inst, cs, tableau_doubles=64)
    st = tiny["status"].cpu().numpy()
    assert ((st == 2) | (a["info"].cpu().numpy()[..., 1] == 0) | (a["status"].cpu().numpy() == 0)).all()
