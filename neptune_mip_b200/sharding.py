"""Instance sharding across ranks (one process per GPU, `torch.distributed`).

Independent instances shard with no data-path collective: rank r owns a
contiguous block of the sweep; only the per-instance verdicts (scores, flags) are gathered at the end.
Works with NCCL (GPU tensors) and gloo (CPU tensors; used by the world_size-2 tests)."""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; blocks differ in size by at most one item."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """All-gather per-instance rows (first dim = this rank's block) into sweep order on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_items, r, ws)[1] - shard_range(n_items, r, ws)[0] for r in range(ws)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def max_over_ranks(value: float, device="cpu") -> float:
    rank, ws = world()
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def solve_sweep(n_items: int, solve_block: Callable[[int, int], Sequence[torch.Tensor]]):
    """Run `solve_block(lo, hi)` on this rank's block and gather every returned per-instance tensor."""
    rank, ws = world()
    lo, hi = shard_range(n_items, rank, ws)
    outs = solve_block(lo, hi)
    return [gather_rows(t, n_items) for t in outs]
