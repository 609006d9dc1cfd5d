"""world_size-2 gloo run (CPU) of the function-block-sharded matrix-free PDHG iteration (SURVEY.md section 8(e),
the plan for C4 on 8 GPUs): every rank owns a block of functions, the 2N coupling multipliers are replicated, and ONE
all-reduce of 2N doubles per iteration makes the C4 / C2 activities global.  The sharded ranks must reproduce the
unsharded iteration (tests/mf_reference.MatrixFree, which tests/test_mf_reference.py ties to the CSR iteration on the
oracle's matrix and tests/test_pdhg_mf_gpu.py to the kernels)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ITERS = 96


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.dirname(here))
    from helpers import arrays_of
    from mf_reference import ShardedMatrixFree
    from neptune_mip_b200 import sharding, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        a = arrays_of(synth.random_payload(12, 5, 1, node_cores=25))

        def allreduce(v):
            t = torch.from_numpy(v)                   # shares memory with v: summed in place
            dist.all_reduce(t)

        f0, f1 = sharding.shard_range(a["F"], rank, ws)            # the same block rule as the instance sweeps
        lp = ShardedMatrixFree(a, f0, f1, allreduce)
        for _ in range(ITERS):
            lp.step()
        lp.flush()
        kkt = lp.kkt()
        q.put((rank, f0, f1, lp.x, lp.c, lp.y1, lp.y3, lp.yS, lp.y2, lp.y4, kkt, lp.exchanged_doubles))
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_the_unsharded_iteration():
    from helpers import arrays_of
    from mf_reference import MatrixFree
    from neptune_mip_b200 import synth
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in range(ws)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = arrays_of(synth.random_payload(12, 5, 1, node_cores=25))
    ref = MatrixFree(a)
    for _ in range(ITERS):
        ref.step()
    N = a["N"]
    assert [g[1] for g in got] + [got[-1][2]] == [0, 3, 5]          # 5 functions over 2 ranks: blocks of 3 and 2
    for name, k, want in (("x", 3, ref.x), ("c", 4, ref.c), ("y1", 5, ref.y1), ("y3", 6, ref.y3), ("yS", 7, ref.yS)):
        have = np.concatenate([g[k] for g in got])
        assert np.abs(have - want).max() <= 1e-11 * (1 + np.abs(want).max()), name
    for g in got:                                                    # replicated multipliers: identical everywhere
        assert np.abs(g[8] - ref.y2).max() <= 1e-11 * (1 + np.abs(ref.y2).max())
        assert np.abs(g[9] - ref.y4).max() <= 1e-11 * (1 + np.abs(ref.y4).max())
    assert np.array_equal(got[0][8], got[1][8]) and np.array_equal(got[0][9], got[1][9])
    want = ref.kkt(ref.state())
    for g in got:                                                    # global KKT pieces, same numbers on every rank
        assert g[10] == got[0][10]
        for u, v in zip(g[10], want):
            assert abs(u - v) <= 1e-10 * (1 + abs(v))
        assert g[11] == 2 * N * ITERS + N                            # one exchange of 2N doubles per iteration (+ flush)
