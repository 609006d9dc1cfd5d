"""GPU experiment: matrix-free PDHG (neptune_pdhg_mf_solve) against the CSR solver (neptune_pdhg_solve with
ruiz_iters = 0, i.e. the same Pock-Chambolle step sizes) -- iterate equality after a fixed number of
iterations, converged objectives, and time per iteration / achieved GB/s at C2 (batch), C3 and the per-GPU
share of C4.  Writes gpurun_out/mf_check.json.   python tools/mf_check.py [--quick] [--profile]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from neptune_mip_b200 import device, synth
from neptune_mip_b200._lib import FLAG_STRENGTHEN
from neptune_mip_b200.core.utils import data_to_solver_input

out = {"equality": [], "converged": [], "timing": []}


def batch_of(N, F, B, cores, seed0=0):
    datas = [data_to_solver_input(synth.random_payload(N, F, seed0 + s, node_cores=cores), 1, with_db=False)
             for s in range(B)]
    return device.InstanceBatch.from_datas(datas)


def synth_batch(N, F, B, seed=0):
    """random instance arrays made on the device (large shapes: no payload round trip)"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    f64 = dict(dtype=torch.float64, device="cuda")
    D = torch.randint(1, 50, (B, N, N), generator=g, device="cuda").to(torch.float64)
    D = torch.floor((D + D.transpose(1, 2)) / 2)
    D.diagonal(dim1=1, dim2=2).zero_()
    W = torch.randint(0, 20, (B, F, N), generator=g, device="cuda").to(torch.float64)
    cores = torch.randint(1, 5, (B, F, N), generator=g, device="cuda").to(torch.float64)
    dest = torch.randint(1, 10, (B, F, N), generator=g, device="cuda").to(torch.float64)
    r = cores / dest
    load = (W.sum(dim=2, keepdim=True) * r).sum(dim=1).mean(dim=1, keepdim=True)     # mean CPU need per node
    Kj = torch.ceil(2.5 * load).expand(B, N).contiguous()
    return device.InstanceBatch(B=B, N=N, F=F, d=D.contiguous(), w=W, r=r, m=torch.full((B, F), 30.0, **f64),
                                Mj=torch.full((B, N), 100.0, **f64), Kj=Kj, old=torch.ones((B, F, N), **f64),
                                maxd=torch.full((B, F), 1000.0, **f64), cost=torch.full((B, N), 5.0, **f64),
                                budget=300.0)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    e1.synchronize()
    return r, e0.elapsed_time(e1)


def equality(N, F, B, cores, K):
    inst = batch_of(N, F, B, cores)
    mdl = device.assemble(inst, "min_delay", flags=FLAG_STRENGTHEN)
    xa, ya, ra = device.pdhg_solve(mdl, max_iters=K, check_every=K, ruiz_iters=0, eps_rel=1e-12, eps_abs=1e-14)
    xb, yb, rb = device.pdhg_mf_solve(inst, max_iters=K, check_every=K, eps_rel=1e-12, eps_abs=1e-14)
    xr, yr, rr = device.pdhg_mf_solve(inst, max_iters=K, check_every=K, eps_rel=1e-12, eps_abs=1e-14, bulk_copy_kernel=True)
    sx = float(xa.abs().max()) + 1e-300
    sy = float(ya.abs().max()) + 1e-300
    rec = dict(N=N, F=F, B=B, K=K, tma_vs_reg_dx=float((xr - xb).abs().max()), tma_vs_reg_dy=float((yr - yb).abs().max()),
               pobj_reg=float(rr[0]["primal_obj"]), dobj_reg=float(rr[0]["dual_obj"]),
               dx=float((xa - xb).abs().max()), dy=float((ya - yb).abs().max()), xmax=sx, ymax=sy,
               pobj=[float(ra[0]["primal_obj"]), float(rb[0]["primal_obj"])],
               dobj=[float(ra[0]["dual_obj"]), float(rb[0]["dual_obj"])],
               pres=[float(ra[0]["primal_res"]), float(rb[0]["primal_res"])],
               omega=[float(ra[0]["primal_weight"]), float(rb[0]["primal_weight"])])
    out["equality"].append(rec)
    print("EQ", json.dumps(rec), flush=True)


def converged(N, F, B, cores):
    inst = batch_of(N, F, B, cores)
    mdl = device.assemble(inst, "min_delay", flags=FLAG_STRENGTHEN)
    (xa, ya, ra), ta = timed(lambda: device.pdhg_solve(mdl, max_iters=40000, check_every=128, ruiz_iters=0, eps_rel=1e-6, eps_abs=1e-9))
    (xb, yb, rb), tb = timed(lambda: device.pdhg_mf_solve(inst, max_iters=40000, check_every=128, eps_rel=1e-6, eps_abs=1e-9))
    (xc, yc, rc), tc = timed(lambda: device.pdhg_solve(mdl, max_iters=40000, check_every=128, ruiz_iters=10, eps_rel=1e-6, eps_abs=1e-9))
    rec = dict(N=N, F=F, B=B, csr_pc=dict(ms=ta, iters=ra["iters"].tolist(), conv=ra["converged"].tolist(), pobj=ra["primal_obj"].tolist(), dobj=ra["dual_obj"].tolist()),
               mf=dict(ms=tb, iters=rb["iters"].tolist(), conv=rb["converged"].tolist(), pobj=rb["primal_obj"].tolist(), dobj=rb["dual_obj"].tolist()),
               csr_ruiz=dict(ms=tc, iters=rc["iters"].tolist(), conv=rc["converged"].tolist(), pobj=rc["primal_obj"].tolist(), dobj=rc["dual_obj"].tolist()))
    out["converged"].append(rec)
    print("CONV", json.dumps(rec), flush=True)


def timing(name, inst, iters, with_csr):
    N, F, B = inst.N, inst.F, inst.B
    X = F * N * N
    device.pdhg_mf_solve(inst, max_iters=64, check_every=64)          # warm-up (graph instantiation, page-in)
    (_, _, rb), tb = timed(lambda: device.pdhg_mf_solve(inst, max_iters=iters, check_every=iters, eps_rel=1e-12, eps_abs=1e-14))
    device.pdhg_mf_solve(inst, max_iters=64, check_every=64, bulk_copy_kernel=True)
    (_, _, rr), tr = timed(lambda: device.pdhg_mf_solve(inst, max_iters=iters, check_every=iters, eps_rel=1e-12, eps_abs=1e-14, bulk_copy_kernel=True))
    bytes_iter = B * (64 * X + 8 * N * N + 8 * (12 * F * N + 6 * N))
    rec = dict(name=name, N=N, F=F, B=B, iters=iters, mf_ms=tb, mf_us_per_iter=1e3 * tb / iters,
               mf_bytes_per_iter=bytes_iter, mf_gbs=bytes_iter * iters / (tb / 1e3) / 1e9,
               reg_us_per_iter=1e3 * tr / iters, reg_gbs=bytes_iter * iters / (tr / 1e3) / 1e9)
    if with_csr:
        mdl = device.assemble(inst, "min_delay", flags=FLAG_STRENGTHEN)
        device.pdhg_solve(mdl, max_iters=64, check_every=64)
        (_, _, ra), ta = timed(lambda: device.pdhg_solve(mdl, max_iters=iters, check_every=iters, eps_rel=1e-12, eps_abs=1e-14))
        csr_bytes = B * (16 * mdl.nnz + 88 * mdl.cols + 72 * mdl.rows) + 8 * mdl.nnz + 8 * (mdl.rows + mdl.cols + 2)
        rec.update(csr_ms=ta, csr_us_per_iter=1e3 * ta / iters, csr_gbs=csr_bytes * iters / (ta / 1e3) / 1e9,
                   speedup_per_iter=ta / tb)
        del mdl
    out["timing"].append(rec)
    print("TIME", json.dumps(rec), flush=True)
    torch.cuda.empty_cache()


def main():
    if "--ceiling" in sys.argv:       # what plain torch kernels reach on this box for copy / read-modify-write streams
        torch.cuda.set_device(0)
        n = 256 * 25000                                            # one x-shaped stream of the C2 batch (51 MB)
        a, b2, c2, d2 = (torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(4))
        big_a, big_b = torch.rand(1 << 27, dtype=torch.float64, device="cuda"), torch.empty(1 << 27, dtype=torch.float64, device="cuda")
        def rate(fn, nbytes, reps=20):
            fn(); torch.cuda.synchronize()
            _, ms = timed(lambda: [fn() for _ in range(reps)])
            return nbytes * reps / ms / 1e6
        print("CEIL copy 1 GiB f64 (read + write) GB/s %.0f" % rate(lambda: big_b.copy_(big_a), 2 * 8 * (1 << 27)), flush=True)
        print("CEIL a.add_(b) on 51 MB streams, 4 arrays round robin (2 reads + 1 write) GB/s %.0f" %
              rate(lambda: (a.add_(b2), c2.add_(d2)), 2 * 24 * n), flush=True)
        print("CEIL torch._foreach_add_ of two 51 MB pairs GB/s %.0f" % rate(lambda: torch._foreach_add_([a, c2], [b2, d2]), 2 * 24 * n), flush=True)
        for name, inst, iters in (("C3 500x50", synth_batch(500, 50, 1), 1024), ("C4 share 2000x25", synth_batch(2000, 25, 1), 256)):
            X = inst.F * inst.N * inst.N
            device.pdhg_mf_solve(inst, max_iters=32, check_every=32)
            _, ms = timed(lambda: device.pdhg_mf_solve(inst, max_iters=iters, check_every=iters, eps_rel=1e-12, eps_abs=1e-14))
            print("LONG", name, "iters", iters, "us/iter %.1f" % (1e3 * ms / iters), "GB/s %.0f" % (inst.B * (64 * X + 112 * inst.F * inst.N + 8 * inst.N ** 2) * iters / ms / 1e6), flush=True)
        return
    if "--variants" in sys.argv:      # register pass: rows in flight per warp; the small-vector kernel alone
        torch.cuda.set_device(0)
        for name, inst, iters in (("C2 batch 256", synth_batch(50, 10, 256), 512), ("C3 500x50", synth_batch(500, 50, 1), 256),
                                  ("C4 share 2000x25", synth_batch(2000, 25, 1), 64)):
            X = inst.F * inst.N * inst.N
            ref = None
            variants = [(True, 1), (True, 2)] + ([(True, 4)] if inst.N <= 32 else [])
            if inst.N > 32 and inst.N % 2 == 0:
                variants += [(False, 2), (False, 4)] if inst.N <= 64 else [(False, 1), (False, 2)]
            if "--experimental" in sys.argv and inst.N > 32 and inst.N % 2 == 0:
                variants.append(("async", 0))
            if "--experimental" in sys.argv:
                variants.append(("fused", 0))
                variants.append(("lean", 0))
            if "--pair" in sys.argv and inst.N > 32 and inst.N % 2 == 0:
                variants = [(True, 1), (True, 2), ("pair", 1), ("pair", 2)]
            for scalar, u in variants:
                kw = (dict(pair_kernel=True, rows_in_flight=u) if scalar == "pair" else dict(async_kernel=True) if scalar == "async" else dict(fused_kernel=True) if scalar == "fused" else dict(lean_kernel=True) if scalar == "lean"
                      else dict(rows_in_flight=u, vector_kernel=not scalar))
                device.pdhg_mf_solve(inst, max_iters=32, check_every=32, **kw)
                (xu, yu, _), ms = timed(lambda: device.pdhg_mf_solve(inst, max_iters=iters, check_every=iters, eps_rel=1e-12, eps_abs=1e-14, **kw))
                if ref is None:
                    ref = (xu, yu)
                print("VAR", name, "pair version, 16-byte" if scalar == "pair" else "cp.async ring" if scalar == "async" else "fused small vectors, 8-byte" if scalar == "fused" else "pointer-bumped, 8-byte" if scalar == "lean" else ("8-byte" if scalar else "16-byte"), "accesses, rows in flight", u, "us/iter %.1f" % (1e3 * ms / iters),
                      "GB/s %.0f" % (inst.B * 64 * X * iters / ms / 1e6), "max |dx| vs first %.1e" % float((xu - ref[0]).abs().max()),
                      "max |dy| %.1e" % float((yu - ref[1]).abs().max()), flush=True)
            device.pdhg_mf_solve(inst, max_iters=32, check_every=32, _diag=4)
            _, ms = timed(lambda: device.pdhg_mf_solve(inst, max_iters=iters, check_every=iters, eps_rel=1e-12, eps_abs=1e-14, _diag=4))
            print("VAR", name, "small-vector kernel alone us/iter %.1f" % (1e3 * ms / iters), flush=True)
        return
    if "--diag" in sys.argv:          # which part of the TMA pass costs what (results are garbage with diag != 0)
        torch.cuda.set_device(0)
        for name, inst, iters in (("C2 batch 256", synth_batch(50, 10, 256), 512), ("C4 share 2000x25", synth_batch(2000, 25, 1), 32)):
            X = inst.F * inst.N * inst.N
            for diag, what in ((0, "full"), (1, "no compute (bulk loads only)"), (2, "no row sums")):
                device.pdhg_mf_solve(inst, max_iters=32, check_every=32, _diag=diag)
                _, ms = timed(lambda: device.pdhg_mf_solve(inst, max_iters=iters, check_every=iters, eps_rel=1e-12, eps_abs=1e-14, _diag=diag))
                print("DIAG", name, what, "us/iter %.1f" % (1e3 * ms / iters), "GB/s (64 B per element) %.0f" % (inst.B * 64 * X * iters / ms / 1e6), flush=True)
        return
    quick = "--quick" in sys.argv
    profile = "--profile" in sys.argv
    torch.cuda.set_device(0)
    if profile:                       # under ncu: one short C2-batch run only
        inst = synth_batch(50, 10, 256)
        device.pdhg_mf_solve(inst, max_iters=64, check_every=64)
        torch.cuda.synchronize()
        return
    for (N, F, B, cores, K) in [(8, 4, 1, 30, 64), (12, 5, 3, 25, 96), (20, 5, 2, 100, 40), (50, 10, 2, 200, 64),
                                (33, 3, 1, 60, 64), (70, 3, 2, 60, 64), (130, 2, 1, 60, 33), (300, 2, 1, 60, 33)]:
        try:
            equality(N, F, B, cores, K)
        except Exception as e:        # keep going: every section reports on its own
            print("EQ-FAIL", N, F, B, repr(e), flush=True)
            out["equality"].append(dict(N=N, F=F, B=B, error=repr(e)))
    for (N, F, B, K) in [(1100, 2, 1, 33), (64, 3, 5, 64), (2, 1, 1, 64)]:      # TMA vs register kernel only (split rows / edge widths)
        try:
            inst = synth_batch(N, F, B, seed=3)
            xb, yb, rb = device.pdhg_mf_solve(inst, max_iters=K, check_every=K, eps_rel=1e-12, eps_abs=1e-14)
            xr, yr, rr = device.pdhg_mf_solve(inst, max_iters=K, check_every=K, eps_rel=1e-12, eps_abs=1e-14, bulk_copy_kernel=True)
            rec = dict(N=N, F=F, B=B, K=K, tma_vs_reg_dx=float((xr - xb).abs().max()), tma_vs_reg_dy=float((yr - yb).abs().max()),
                       ymax=float(yr.abs().max()), pobj=[float(rb[0]["primal_obj"]), float(rr[0]["primal_obj"])])
            out["equality"].append(rec)
            print("EQ2", json.dumps(rec), flush=True)
        except Exception as e:
            print("EQ2-FAIL", N, F, B, repr(e), flush=True)
    if "--converge" in sys.argv:
        try:
            converged(12, 5, 4, 25)
            converged(50, 10, 2, 200)
        except Exception as e:
            print("CONV-FAIL", repr(e), flush=True)
    try:
        timing("C2 batch 256", synth_batch(50, 10, 256), 1024, "--csr" in sys.argv)
        timing("C2 batch 64", synth_batch(50, 10, 64), 1024, False)
        if not quick:
            timing("C3 500x50", synth_batch(500, 50, 1), 256, False)
            timing("C4 share 2000x25", synth_batch(2000, 25, 1), 64, False)
            if "--c5" in sys.argv:
                timing("C5 20x5 x4096", synth_batch(20, 5, 4096), 512, False)
    except Exception as e:
        print("TIME-FAIL", repr(e), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "mf_check.json"), "w"), indent=1)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("mf_check done in %.1f s" % (time.time() - t0))
