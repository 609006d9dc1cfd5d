"""Function-block-sharded PDHG for one large instance (BASELINE config 4).

One process per GPU.  Rank g owns the functions of its block: the columns x[.,f,.], c[f,.] and the
rows C1a/C1b/C3 of those f are purely local; the coupling rows C2 (memory) and C4 (CPU) are
replicated.  Per iteration every rank computes the partial activity of the 2N coupling rows over its
own columns inside the same fused SpMV pass, ONE all-reduce(sum) of 2N doubles makes them global, and
a tiny kernel applies the dual update identically on every rank.  A^T y needs no exchange (the
coupling multipliers are replicated).  x is never replicated: at C4 that would be 6.4 GB per
iteration instead of 32 KB.
"""
from __future__ import annotations

import copy
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, device
from ._lib import check
from .sharding import shard_range, world


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def slice_functions(data, lo, hi):
    """`Data` restricted to functions [lo, hi) (nodes, delays, node capacities unchanged)."""
    d = copy.copy(data)
    d.functions = list(data.functions[lo:hi])
    for k in ("function_memory_matrix", "max_delay_matrix"):
        setattr(d, k, np.asarray(getattr(data, k))[lo:hi])
    for k in ("workload_matrix", "core_per_req_matrix", "old_allocations_matrix", "cores_matrix"):
        setattr(d, k, np.asarray(getattr(data, k))[lo:hi])
    return d


class ShardedLP:
    def __init__(self, data, flags=0, eta=0.99, omega=1.0):
        self.rank, self.world = world()
        F, N = len(data.functions), len(data.nodes)
        self.f_lo, self.f_hi = shard_range(F, self.rank, self.world)
        self.N, self.Fg = N, self.f_hi - self.f_lo
        self.inst = device.InstanceBatch.from_datas([slice_functions(data, self.f_lo, self.f_hi)])
        m = self.model = device.assemble(self.inst, "min_delay", flags=flags)
        Fg = self.Fg
        # coupling rows of the local (min-delay) layout: C2 = [2*Fg*N, +N), C4 = [3*Fg*N + N, +N)
        self.d = (2 * Fg * N, 2 * Fg * N + N, 3 * Fg * N + N, 3 * Fg * N + 2 * N)
        self.lib = _lib.load()
        dev = m.val.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.plan, self.meta = self._plan(m.rows, m.row_ptr)
        self.planT, self.metaT = self._plan(m.cols, m.rowT_ptr)
        # Pock-Chambolle preconditioner from |A| row / column sums; coupling rows summed over ranks
        ones_r, ones_c = torch.ones((1, m.rows), **f64), torch.ones((1, m.cols), **f64)
        colacc, rowacc = torch.zeros((1, m.cols), **f64), torch.zeros((1, m.rows), **f64)
        check(self.lib.neptune_abs_sums(1, m.rows, m.cols, m.nnz, _p(m.rowT_ptr), _p(m.colT_idx), _p(m.valT),
                                        _p(m.lo), _p(m.hi), _p(ones_r), _p(ones_c), 1, _p(colacc), _p(rowacc),
                                        _stream()), "neptune_abs_sums")
        coup = self._coupling(rowacc)
        self.allreduce(coup)
        self._set_coupling(rowacc, coup)
        self.T = torch.where(colacc > 0, 1.0 / colacc, torch.ones_like(colacc))
        self.S = torch.where(rowacc > 0, 1.0 / rowacc, torch.ones_like(rowacc))
        self.tau, self.sigma = eta / omega, eta * omega
        self.x = torch.zeros((1, m.cols), **f64)
        self.xbar = torch.zeros((1, m.cols), **f64)
        self.xsum = torch.zeros((1, m.cols), **f64)
        self.y = torch.zeros((1, m.rows), **f64)
        self.ysum = torch.zeros((1, m.rows), **f64)
        self.act = torch.zeros((1, 2 * N), **f64)
        self.iters = 0

    # -- plumbing --------------------------------------------------------------------------------------
    def _plan(self, n_rows, ptr):
        nb = C.c_int64()
        check(self.lib.neptune_spmv_plan_bytes(n_rows, C.byref(nb)), "neptune_spmv_plan_bytes")
        plan = torch.empty(nb.value, dtype=torch.uint8, device=ptr.device)
        meta = (C.c_int32 * 3)()
        check(self.lib.neptune_spmv_plan(n_rows, _p(ptr), _p(plan), meta, _stream()), "neptune_spmv_plan")
        return plan, meta

    def _coupling(self, v):
        d0, d1, d2, d3 = self.d
        return torch.cat([v[:, d0:d1], v[:, d2:d3]], dim=1).contiguous()

    def _set_coupling(self, v, c):
        d0, d1, d2, d3 = self.d
        v[:, d0:d1] = c[:, : d1 - d0]
        v[:, d2:d3] = c[:, d1 - d0:]

    def allreduce(self, t):
        if self.world == 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(t)                       # NVLink / NVSwitch, on the current stream
        else:                                        # gloo (single-GPU test boxes): through host memory
            h = t.cpu()
            dist.all_reduce(h)
            t.copy_(h)

    # -- the iteration ---------------------------------------------------------------------------------
    def iterate(self, n=1):
        m, lib = self.model, self.lib
        d0, d1, d2, d3 = self.d
        for _ in range(n):
            check(lib.neptune_pdhg_primal_step(1, m.rows, m.cols, m.nnz, _p(m.rowT_ptr), _p(m.colT_idx), _p(m.valT),
                                               _p(self.planT), self.metaT, _p(m.obj), _p(m.col_lb), _p(m.col_ub),
                                               _p(self.T), C.c_double(self.tau), _p(self.y), _p(self.x),
                                               _p(self.xbar), _p(self.xsum), _stream()), "neptune_pdhg_primal_step")
            check(lib.neptune_pdhg_dual_step(1, m.rows, m.cols, m.nnz, _p(m.row_ptr), _p(m.col_idx), _p(m.val),
                                             _p(self.plan), self.meta, _p(m.lo), _p(m.hi), _p(self.S),
                                             C.c_double(self.sigma), _p(self.xbar), _p(self.y), _p(self.ysum),
                                             d0, d1, d2, d3, _p(self.act), _stream()), "neptune_pdhg_dual_step")
            self.allreduce(self.act)
            check(lib.neptune_pdhg_dual_rows(1, m.rows, d0, d1, d2, d3, C.c_double(self.sigma), _p(self.act),
                                             _p(m.lo), _p(m.hi), _p(self.S), _p(self.y), _p(self.ysum), _stream()),
                  "neptune_pdhg_dual_rows")
        self.iters += n

    # -- restarted solve: KKT pieces are reduced over ranks, so every rank takes the same decisions --------
    def _kkt(self, x, y):
        """(primal residual^2, dual residual^2, primal objective, dual objective) of (x, y), global."""
        m = self.model
        d0, d1, d2, d3 = self.d
        a = device.spmv(m, x)                                 # local activities; coupling rows are partial
        coup = self._coupling(a)
        self.allreduce(coup)
        self._set_coupling(a, coup)
        lo, hi = m.lo, m.hi
        viol = a - torch.minimum(torch.maximum(a, lo), hi)
        own = torch.ones_like(a)
        if self.rank != 0:                                    # replicated rows are counted once (rank 0)
            own[:, d0:d1] = 0.0
            own[:, d2:d3] = 0.0
        pres2 = (viol * viol * own).sum()
        pos, neg = torch.clamp(y, min=0.0), torch.clamp(y, max=0.0)
        fin_h, fin_l = torch.isfinite(hi), torch.isfinite(lo)
        dobj = -(torch.where(fin_h, hi, torch.zeros_like(hi)) * pos * own).sum() \
               - (torch.where(fin_l, lo, torch.zeros_like(lo)) * neg * own).sum()
        dres2 = ((pos * (~fin_h)) ** 2 * own).sum() + ((neg * (~fin_l)) ** 2 * own).sum()
        rc = m.obj + device.spmv_t(m, y)
        rpos, rneg = torch.clamp(rc, min=0.0), torch.clamp(rc, max=0.0)
        fin_u = torch.isfinite(m.col_ub)
        dobj = dobj + (m.col_lb * rpos).sum() + (torch.where(fin_u, m.col_ub, torch.zeros_like(rc)) * rneg).sum()
        dres2 = dres2 + ((rneg * (~fin_u)) ** 2).sum()
        pobj = (m.obj * x).sum()
        v = torch.stack([pres2, dres2, pobj, dobj]).reshape(1, 4).contiguous()
        self.allreduce(v)
        return [float(t) for t in v[0].cpu()]

    def _norm2(self, vx, vy):
        """Squared preconditioned norms of a primal / dual displacement, global."""
        own = torch.ones_like(vy)
        if self.rank != 0:
            d0, d1, d2, d3 = self.d
            own[:, d0:d1] = 0.0
            own[:, d2:d3] = 0.0
        v = torch.stack([(vx * vx / self.T).sum(), (vy * vy / self.S * own).sum()]).reshape(1, 2).contiguous()
        self.allreduce(v)
        return float(v[0, 0]), float(v[0, 1])

    def solve(self, max_iters=20000, check_every=128, eps_rel=1e-6, eps_abs=1e-9, eta=0.99):
        """Restarted, averaged PDHG with the decisions of csrc/pdhg.cu (KKT-error restarts 0.2 / 0.8 / 0.36,
        primal weight clamped to a factor 2 per restart), taken identically on every rank."""
        m = self.model
        fb = torch.maximum(torch.where(torch.isfinite(m.lo), m.lo.abs(), torch.zeros_like(m.lo)),
                           torch.where(torch.isfinite(m.hi), m.hi.abs(), torch.zeros_like(m.hi)))
        zx = torch.zeros_like(self.x)
        nc2, nb2 = self._norm2(m.obj * self.T, fb * self.S)          # ||D_c c||^2, ||D_r b||^2 (T = dc^2, S = dr^2)
        _, nb_plain = self._norm2(zx, fb * torch.sqrt(self.S))
        nc_plain, _ = self._norm2(m.obj * torch.sqrt(self.T), torch.zeros_like(self.y))
        norm_b, norm_c = nb_plain ** 0.5, nc_plain ** 0.5
        omega = (nc2 ** 0.5) / (nb2 ** 0.5) if nb2 > 1e-20 and nc2 > 1e-20 else 1.0
        kkt_restart = kkt_prev = float("inf")
        xr, yr = self.x.clone(), self.y.clone()
        since = restarts = 0
        self.xsum.zero_(); self.ysum.zero_()
        total = 0
        info = {}
        while total < max_iters:
            self.tau, self.sigma = eta / omega, eta * omega
            self.iterate(check_every)
            total += check_every; since += check_every
            cands = []
            for (cx, cy) in ((self.x, self.y), (self.xsum / since, self.ysum / since)):
                p2, d2, po, do = self._kkt(cx, cy)
                gap = abs(po - do)
                k = (omega * omega * p2 + d2 / (omega * omega) + gap * gap) ** 0.5
                ok = p2 ** 0.5 <= eps_abs + eps_rel * norm_b and d2 ** 0.5 <= eps_abs + eps_rel * norm_c and \
                    gap <= eps_abs + eps_rel * (abs(po) + abs(do))
                cands.append((k, ok, po, do, p2 ** 0.5, d2 ** 0.5))
            pick = 1 if (cands[1][1] and not cands[0][1]) else (0 if (cands[0][1] and not cands[1][1])
                                                                 else (1 if cands[1][0] < cands[0][0] else 0))
            k, ok, po, do, pr, dr = cands[pick]
            info = dict(primal_obj=po, dual_obj=do, primal_res=pr, dual_res=dr, iters=total, restarts=restarts,
                        converged=bool(cands[0][1] or cands[1][1]), primal_weight=omega)
            if info["converged"] or total >= max_iters:
                if pick == 1:
                    self.x.copy_(self.xsum / since); self.y.copy_(self.ysum / since)
                break
            act = k <= 0.2 * kkt_restart or (k <= 0.8 * kkt_restart and k > kkt_prev) or \
                (since >= 0.36 * total and restarts > 0)
            kkt_prev = k
            if act:
                kkt_restart = k; restarts += 1
                if pick == 1:
                    self.x.copy_(self.xsum / since); self.y.copy_(self.ysum / since)
                dx2, dy2 = self._norm2(self.x - xr, self.y - yr)
                if dx2 > 1e-20 and dy2 > 1e-20:
                    nw = (dy2 / dx2) ** 0.25 * omega ** 0.5           # exp(0.5 log(dy/dx) + 0.5 log omega)
                    omega = min(max(nw, 0.5 * omega), 2.0 * omega)
                xr.copy_(self.x); yr.copy_(self.y)
                self.xsum.zero_(); self.ysum.zero_(); since = 0
        return info

    def primal_objective(self, average=False):
        x = self.xsum / max(self.iters, 1) if average else self.x
        v = (self.model.obj * x).sum().reshape(1)
        self.allreduce(v)
        return float(v.item())

    def bytes_per_iteration(self):
        """Algorithmic bytes per rank (pattern counted once, B = 1) + the exchange."""
        m = self.model
        return 24 * m.nnz + 96 * m.cols + 80 * m.rows, 2 * self.N * 8
