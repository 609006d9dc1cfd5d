"""Batched solve path: B same-shaped instances through the LP relaxation (PDHG: matrix-free for the min-delay
model, assembly + CSR otherwise) -> EFTTC -> local search -> exact check in one go (what bench.py times, and what a sweep like BASELINE.json's config 5 calls)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import device
from ._lib import FLAG_STRENGTHEN, KINDS, OK_ALL


@dataclass
class BatchParams:
    kind: str = "min_delay"
    alpha: float = 0.5
    lp_iters: int = 2048          # PDHG iterations on the strengthened relaxation (bound + rounding guide)
    lp_check_every: int = 256
    lp_path: str = "auto"         # "auto": matrix-free PDHG for the min-delay model, assembled CSR otherwise; "csr": always CSR
    chains: int = 16              # local-search chains per instance
    sweeps: int = 200
    rng_seed: int = 1


@dataclass
class BatchResult:
    c: torch.Tensor               # uint8 [B,F,N]
    x: torch.Tensor               # float64 [B,N,F,N]
    n: torch.Tensor               # float64 [B,N]
    flags: torch.Tensor           # int32 [B]
    scores: torch.Tensor          # float64 [B,3]
    lp: Optional[np.ndarray]      # PDHG result records (primal/dual objective, residuals, iterations)
    pdhg_ms: float = 0.0          # device time of the PDHG call (CUDA events on the launch stream)
    pdhg_iters: int = 0
    model_dims: tuple = (0, 0, 0)
    pdhg_bytes_per_iter: int = 0  # algorithmic bytes one PDHG iteration of the whole batch moves
    pdhg_path: str = ""           # "matrix-free" | "csr"


def solve_batch(inst: device.InstanceBatch, prm: BatchParams, time_pdhg: bool = False) -> BatchResult:
    kind = prm.kind
    lp_res, guide, pdhg_ms, iters, dims = None, None, 0.0, 0, (0, 0, 0)
    bytes_iter, path = 0, ""
    if prm.lp_iters > 0:
        N, F, B = inst.N, inst.F, inst.B
        X = F * N * N
        if time_pdhg:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if KINDS.get(kind, kind) == 0 and prm.lp_path != "csr":
            # min-delay: the relaxation is solved matrix-free (nothing is assembled; every coefficient of the
            # strengthened model is regenerated from d, w, r, m inside the iteration kernels)
            rows, cols, nnz = device.model_sizes(N, F, 0, FLAG_STRENGTHEN)
            if time_pdhg:
                e0.record()
            xs, ys, lp_res = device.pdhg_mf_solve(inst, max_iters=prm.lp_iters, check_every=prm.lp_check_every,
                                                  eps_rel=1e-6, eps_abs=1e-9)
            bytes_iter, path = B * (64 * X + 112 * F * N + 8 * N * N), "matrix-free"
        else:
            lp = device.assemble(inst, kind, prm.alpha, flags=FLAG_STRENGTHEN)
            rows, cols, nnz = lp.rows, lp.cols, lp.nnz
            if time_pdhg:
                e0.record()
            xs, ys, lp_res = device.pdhg_solve(lp, max_iters=prm.lp_iters, check_every=prm.lp_check_every,
                                               eps_rel=1e-6, eps_abs=1e-9)
            bytes_iter, path = B * (16 * nnz + 88 * cols + 72 * rows) + 8 * nnz + 8 * (rows + cols + 2), "csr"
            del lp
        if time_pdhg:
            e1.record()
            e1.synchronize()
            pdhg_ms = e0.elapsed_time(e1)
        dims = (rows, cols, nnz)
        iters = int(lp_res["iters"].max())
        guide = xs[:, X:X + F * N].contiguous()
        del xs, ys
    seeds = torch.stack([device.efttc(inst, k, prm.alpha)[0] for k in ("min_delay", "min_util", "min_delay_util")],
                        dim=1).contiguous()
    best_c, best_obj, _ = device.local_search(inst, kind, seeds, prm.alpha, prm.chains, prm.sweeps,
                                              prm.rng_seed, guide)
    fallback = seeds[:, KINDS[kind]]
    bad = ~torch.isfinite(best_obj)
    if bool(bad.any()):
        best_c[bad] = fallback[bad]
    # final routing honours the CPU rows (splits flows on binding nodes) and closes pods nobody uses
    best_c, x, n, _, _ = device.route_capacitated(inst, best_c)
    flags, scores = device.check_solution(inst, x, device.u8_to_f64(best_c), n, prm.alpha)
    bad = flags != OK_ALL
    if bool(bad.any()):
        # rare (a pod starved by a split flow that no free source can top up): fall back, per instance, to the
        # first EFTTC seed that passes every check -- a worse objective, never an infeasible answer
        for k in (KINDS[kind], 1, 2, 0):
            cf, xf, nf, _, _ = device.route_capacitated(inst, seeds[:, k].contiguous())
            ff, sf = device.check_solution(inst, xf, device.u8_to_f64(cf), nf, prm.alpha)
            take = bad & (ff == OK_ALL)
            if bool(take.any()):
                best_c[take], x[take], n[take], flags[take], scores[take] = cf[take], xf[take], nf[take], ff[take], sf[take]
                bad = flags != OK_ALL
            if not bool(bad.any()):
                break
    return BatchResult(best_c, x, n, flags, scores, lp_res, pdhg_ms, iters, dims, bytes_iter, path)
