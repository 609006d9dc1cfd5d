"""(c1) exact evaluation parity: the device checkers/scorers against the oracle restatement of
efttc/utils/constraints_step1.py + objectives.py -- flags bit-exact, scores to 1e-12 relative."""
import numpy as np
import pytest

from helpers import arrays_of, cuda_batch, float_payload, small_payloads
from neptune_mip_b200 import synth
from oracle import checkers, efttc as oefttc

pytestmark = pytest.mark.gpu


def _gpu_check(payload, x, c, n, alpha):
    import torch
    from neptune_mip_b200 import device
    inst = cuda_batch([payload])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()[None]  # noqa: E731
    flags, scores = device.check_solution(inst, t(x), t(c), t(n), alpha)
    return int(flags.cpu()[0]), scores.cpu().numpy()[0]


def _oracle(a, x, c, n, alpha):
    fl = checkers.flags_to_mask(checkers.check_all(a, x, c, n))
    sc = (checkers.score_delay(a, x), checkers.score_util(a, n), checkers.score_delay_util(a, n, x, alpha))
    return fl, np.array(sc, dtype=np.float64)


@pytest.mark.parametrize("name,payload,alpha", small_payloads(), ids=lambda v: v if isinstance(v, str) else None)
def test_check_on_efttc_solutions(name, payload, alpha):
    a = arrays_of(payload)
    for kind in ("min_delay", "min_util", "min_delay_util"):
        res = oefttc.solve(a, kind, alpha, strict=False)
        fl, sc = _oracle(a, res.x, res.c, res.n, alpha)
        gfl, gsc = _gpu_check(payload, res.x, res.c, res.n, alpha)
        assert gfl == fl, (kind, gfl, fl)
        assert np.allclose(gsc, sc, rtol=1e-12, atol=1e-12), (kind, gsc, sc)


@pytest.mark.parametrize("seed", range(6))
def test_check_flags_on_broken_solutions(seed):
    """Each checker must fire exactly like the oracle on deliberately damaged solutions."""
    rng = np.random.default_rng(seed)
    payload = synth.random_payload(9, 4, seed, node_cores=15) if seed % 2 else float_payload(9, 4, seed)
    a = arrays_of(payload)
    res = oefttc.solve(a, "min_util", 0.5, strict=False)
    N, F = a["N"], a["F"]
    variants = []
    x, c, n = res.x.copy(), res.c.copy(), res.n.copy()
    variants.append((x, c, n))
    x2 = x.copy(); x2[rng.integers(N), rng.integers(F), :] *= 0.8; variants.append((x2, c, n))       # handle_all
    c2 = c.copy(); c2[:, 0] = 1.0; variants.append((x, c2, n))                                       # c_x, memory
    n2 = 1.0 - n; variants.append((x, c, n2))                                                       # n_c
    n3 = np.ones(N); variants.append((x, c, n3))
    x3 = rng.random((N, F, N)); x3 /= x3.sum(axis=2, keepdims=True); variants.append((x3, np.ones((F, N)), n3))
    x4 = x.copy(); x4 *= 1.0 + 1e-7; variants.append((x4, c, n))
    x5 = np.round(x3, 3); variants.append((x5, np.ones((F, N)), n3))                               # response-rounded x
    for (vx, vc, vn) in variants:
        fl, sc = _oracle(a, vx, vc, vn, 0.37)
        gfl, gsc = _gpu_check(payload, vx, vc, vn, 0.37)
        assert gfl == fl, (gfl, fl)
        assert np.allclose(gsc, sc, rtol=1e-11, atol=1e-11)


def test_cpu_threshold_is_bit_exact():
    """A CPU row sitting exactly at K + 1e-6 and one ulp above it: verdicts must flip like the oracle's."""
    payload = float_payload(6, 3, 11)
    a = arrays_of(payload)
    res = oefttc.solve(a, "min_delay", 0.5, strict=False)
    load = checkers.cpu_load(a, res.x)
    j = int(np.argmax(load))
    for delta in (0.0, 1):
        K = np.array(a["Kj"], dtype=np.float64)
        thr = load[j] - 1e-6
        K[j] = np.nextafter(thr, np.inf) if delta else np.nextafter(thr, -np.inf)
        p2 = dict(payload); p2["node_cores"] = K.tolist()
        a2 = arrays_of(p2)
        fl, _ = _oracle(a2, res.x, res.c, res.n, 0.5)
        gfl, _ = _gpu_check(p2, res.x, res.c, res.n, 0.5)
        assert gfl == fl


def test_route_matches_change_x_one():
    import torch
    from neptune_mip_b200 import device
    rng = np.random.default_rng(3)
    payloads = [synth.random_payload(13, 4, s, node_cores=50) for s in range(3)]
    # force exact ties in the delay matrix
    for p in payloads:
        D = np.array(p["node_delay_matrix"]); D[D % 3 == 0] = 6; np.fill_diagonal(D, 0)
        p["node_delay_matrix"] = D.tolist()
    inst = cuda_batch(payloads)
    c = (rng.random((3, 4, 13)) < 0.3).astype(np.uint8)
    c[0, 2, :] = 0                                                   # a function with no pod at all
    x, n = device.route_placements(inst, torch.from_numpy(c).cuda())
    x, n = x.cpu().numpy(), n.cpu().numpy()
    for b, p in enumerate(payloads):
        a = arrays_of(p)
        xr = np.zeros((13, 4, 13))
        for f in range(4):
            oefttc._route_function(a, c[b].astype(bool), xr, f)
        assert np.array_equal(x[b], xr)
        assert np.array_equal(n[b], c[b].any(axis=0).astype(np.float64))


def test_eval_placements_matches_route_plus_check():
    import torch
    from neptune_mip_b200 import device
    rng = np.random.default_rng(7)
    payloads = [synth.random_payload(16, 5, s, node_cores=40) for s in range(2)]
    inst = cuda_batch(payloads)
    P = 24
    c = (rng.random((2, P, 5, 16)) < 0.2).astype(np.uint8)
    c[:, 0] = 0; c[:, 0, :, 0] = 1                                    # everything on node 0
    c[:, 1] = 1                                                       # everything everywhere (memory breaks)
    obj, flags, over = device.eval_placements(inst, torch.from_numpy(c).cuda(), alpha=0.5)
    obj, flags, over = obj.cpu().numpy(), flags.cpu().numpy(), over.cpu().numpy()
    for b, p in enumerate(payloads):
        a = arrays_of(p)
        for q in range(P):
            cb = c[b, q].astype(bool)
            xr = np.zeros((16, 5, 16))
            for f in range(5):
                oefttc._route_function(a, cb, xr, f)
            nn = cb.any(axis=0).astype(np.float64)
            fl = checkers.flags_to_mask(checkers.check_all(a, xr, cb.astype(np.float64), nn))
            load = checkers.cpu_load(a, xr)
            near = np.abs(load - a["Kj"] - 1e-6).min() < 1e-9       # skip knife-edge CPU verdicts
            if not near:
                assert flags[b, q] == fl, (b, q, flags[b, q], fl)
            assert np.isclose(obj[b, q, 0], checkers.score_delay(a, xr), rtol=1e-12, atol=1e-9)
            assert obj[b, q, 1] == checkers.score_util(a, nn)
            assert np.isclose(obj[b, q, 2], checkers.score_delay_util(a, nn, xr, 0.5), rtol=1e-6, atol=1e-9)
            assert np.isclose(over[b, q], np.maximum(load - a["Kj"], 0)[load > a["Kj"] + 1e-6].sum(), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("seed", range(8))
def test_route_capacitated_matches_routing_lp(seed):
    """CPU-binding placements: the device routing must be feasible for the reference's checkers and
    (1-2 binding nodes, the regime of the HiGHS optima) hit the exact LP value to 1e-9 relative."""
    import torch
    from neptune_mip_b200 import device
    from oracle import mip as omip, routing
    payload = synth.config_payload("C5", seed)
    a = arrays_of(payload)
    ref = omip.solve_step1(a, "min_delay")
    c = (ref["c"] > 0.5)
    lp = routing.lp_routing(a, c)
    assert lp is not None and abs(lp[0] - ref["objective"]) <= 1e-6 * (1 + abs(ref["objective"]))
    inst = cuda_batch([payload])
    c_out, x, n, obj, feas = device.route_capacitated(inst, torch.from_numpy(c.astype(np.uint8)).cuda()[None].contiguous())
    assert int(feas.cpu()[0]) == 1
    flags, scores = device.check_solution(inst, x, device.u8_to_f64(c_out), n)
    assert int(flags.cpu()[0]) == 63
    assert abs(float(obj.cpu()[0]) - lp[0]) <= 1e-9 * (1 + abs(lp[0])), (float(obj.cpu()[0]), lp[0])
    assert abs(float(scores.cpu()[0, 0]) - lp[0]) <= 1e-9 * (1 + abs(lp[0]))
