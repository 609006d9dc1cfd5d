"""(d2) round-robin delay-improvement greedy (`neptune_site_greedy`, csrc/site.cu): placements equal to the numpy
restatement (oracle/site.py) bit for bit; at BASELINE config 4 (2000 x 200) the placement is routed by the
capacity-aware router and passes every checker of the reference."""
import numpy as np
import pytest

from helpers import arrays_of, cuda_batch, float_payload
from neptune_mip_b200 import synth
from oracle import site as osite

CASES = [("C1-3x2", lambda s: synth.test_py_payload()), ("8x4", lambda s: synth.random_payload(8, 4, s, node_cores=30)),
         ("20x5", lambda s: synth.random_payload(20, 5, s, node_cores=100)), ("50x10", lambda s: synth.random_payload(50, 10, s, node_cores=200)),
         ("70x3", lambda s: synth.random_payload(70, 3, s, node_cores=60)), ("7x3-float", lambda s: float_payload(7, 3, 5 + s)),
         ("300x7", lambda s: synth.random_payload(300, 7, s, node_cores=None))]


@pytest.mark.parametrize("name,make", CASES[:4], ids=[c[0] for c in CASES[:4]])
def test_oracle_greedy_serves_every_function_within_memory(name, make):
    a = arrays_of(make(0))
    c, rounds, pods = osite.solve(a)
    assert (c.sum(axis=1) >= 1).all() and pods == int(c.sum()) and rounds >= 1
    assert ((a["m"][:, None] * c).sum(axis=0) <= a["Mj"] + 1e-9).all()
    c2, _, _ = osite.solve(a)
    assert np.array_equal(c, c2)
    # the greedy stops only when nothing gains or nothing fits: no function can still improve on a node with room
    cur = np.where(c[:, None, :] > 0, a["d"][None, :, :], np.inf).min(axis=2)          # [F, i]
    free = a["Mj"] - (a["m"][:, None] * c).sum(axis=0)
    for f in range(a["F"]):
        for j in np.flatnonzero((free + 1e-9 >= a["m"][f]) & (c[f] == 0)):
            assert np.sum(a["w"][f] * np.maximum(cur[f] - a["d"][:, j], 0.0)) <= 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_device_greedy_equals_the_oracle(name, make):
    from neptune_mip_b200 import device
    payloads = [make(s) for s in range(3)]
    inst = cuda_batch(payloads)
    c, info = device.site_greedy(inst)
    c2, _ = device.site_greedy(inst)
    assert np.array_equal(c.cpu().numpy(), c2.cpu().numpy())
    for b, p in enumerate(payloads):
        want, rounds, pods = osite.solve(arrays_of(p))
        assert np.array_equal(c[b].cpu().numpy(), want), (name, b)
        assert info[b].cpu().tolist() == [rounds, pods]


def test_oracle_two_choice_routing_keeps_the_rows():
    """numpy statement of the router: every source routed (rows of x sum to 1), CPU rows hold when it reports feasible"""
    for (N, F, s, k) in [(12, 5, 1, 25), (20, 5, 1, 100), (50, 10, 2, 200)]:
        a = arrays_of(synth.random_payload(N, F, s, node_cores=k))
        c, _, _ = osite.solve(a)
        c2, x, obj, feas, it = osite.route_two_choice(a, c)
        assert feas and np.abs(x.sum(axis=2) - 1.0).max() <= 1e-12
        load = np.einsum("ifj,fi,fj->j", x, a["w"], a["r"])
        assert (load <= a["Kj"] * (1 + 1e-9)).all()
        assert abs(obj - float(np.einsum("ifj,ij,fi->", x, a["d"], a["w"]))) <= 1e-9 * obj
        assert ((x.sum(axis=0) > 0) <= (c2 > 0)).all()                      # flows only to open pods


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(12, 5, 1, 25), (20, 5, 1, 100), (50, 10, 2, 200), (50, 10, 3, 60), (300, 7, 0, None)])
def test_device_two_choice_routing_equals_the_oracle(shape):
    """same nearest / second-nearest pods, same shares (the device accumulates loads in 2^-30 fixed point: 1e-7)"""
    from neptune_mip_b200 import device
    N, F, s, k = shape
    p = synth.random_payload(N, F, s, node_cores=k)
    a = arrays_of(p)
    inst = cuda_batch([p])
    c, _ = device.site_greedy(inst)
    c2, x, n, obj, feas, it = device.route_two_choice(inst, c)
    wc, wx, wobj, wfeas, wit = osite.route_two_choice(a, c[0].cpu().numpy())
    assert np.array_equal(c2[0].cpu().numpy(), wc)
    assert bool(int(feas.cpu()[0])) == wfeas
    if wfeas:
        assert np.abs(x[0].cpu().numpy() - wx).max() <= 1e-6
        assert abs(float(obj.cpu()[0]) - wobj) <= 1e-7 * (1 + abs(wobj))
        load = np.einsum("ifj,fi,fj->j", x[0].cpu().numpy(), a["w"], a["r"])
        assert (load <= a["Kj"] * (1 + 1e-9) + 1e-9).all()
    c3, x3, _, obj3, _, _ = device.route_two_choice(inst, c)
    assert np.array_equal(x3.cpu().numpy(), x.cpu().numpy()) and float(obj3.cpu()[0]) == float(obj.cpu()[0])      # deterministic


@pytest.mark.gpu
def test_c4_feasible_placement():
    """BASELINE config 4: 2000 nodes x 200 functions.  Greedy placement -> two-choice routing -> the reference's
    checkers: every flag set (handle_all_requests, memory, CPU, c<->x, n<->c)."""
    import torch
    from neptune_mip_b200 import device
    from neptune_mip_b200._lib import OK_C_X, OK_CPU, OK_HANDLE, OK_MEMORY, OK_N_C
    from neptune_mip_b200.core.utils import data_to_solver_input
    inst = device.InstanceBatch.from_datas([data_to_solver_input(synth.random_payload(2000, 200, 0, node_cores=None), 1, with_db=False)])
    c, info = device.site_greedy(inst)
    assert int((c[0].sum(dim=1) == 0).sum()) == 0
    c2, x, n, obj, feas, iters = device.route_two_choice(inst, c)
    assert int(feas.cpu()[0]) == 1
    flags, scores = device.check_solution(inst, x, device.u8_to_f64(c2), n)
    fl = int(flags.cpu()[0])
    want = OK_HANDLE | OK_MEMORY | OK_CPU | OK_C_X | OK_N_C
    assert fl & want == want, bin(fl)
    assert abs(float(scores.cpu()[0, 0]) - float(obj.cpu()[0])) <= 1e-9 * float(obj.cpu()[0])
    del x
    torch.cuda.empty_cache()
