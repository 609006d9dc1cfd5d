"""world_size-2 gloo run of the instance-sharding host logic (the N>1 path of bench.py / sweeps)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neptune_mip_b200 import sharding


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 4096, 4097):
        for ws in (1, 2, 3, 8):
            blocks = [sharding.shard_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[k][1] == blocks[k + 1][0] for k in range(ws - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        def solve_block(lo, hi):          # stand-in for the GPU solve: a per-instance checksum
            idx = torch.arange(lo, hi, dtype=torch.float64)
            return [torch.stack([idx, idx * idx + rank * 0.0], dim=1), (idx % 3).to(torch.int32)]
        scores, flags = sharding.solve_sweep(n_items, solve_block)
        t = sharding.max_over_ranks(10.0 + rank)
        q.put((rank, scores.tolist(), flags.tolist(), t))
    finally:
        dist.destroy_process_group()


def test_sweep_gathers_in_order_on_two_ranks():
    n_items, ws = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n_items, q)) for r in range(ws)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scores, flags, t in got:
        assert scores == [[float(i), float(i * i)] for i in range(n_items)]
        assert flags == [i % 3 for i in range(n_items)]
        assert t == 11.0                   # max over ranks
