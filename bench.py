#!/usr/bin/env python
"""bench.py -- placement instances/sec on BASELINE.json's config 2 (50 nodes x 10 functions, min-delay).

A step = one pass of the whole hot path over one batch of B synthetic C2 instances:
PDHG on the strengthened LP relaxation (b; matrix-free for the min-delay model: the coefficients of (a)'s
matrix are regenerated inside the iteration kernels, nothing is assembled) -> EFTTC seeds (d) -> batched
local search (c2) -> exact routing + the reference's checkers/scorers (c1).  `value` is instances/s with the
inputs already in HBM; `e2e` repeats the measurement from pinned HOST buffers through the batched
plugin call, with the H2D copy of every input and the D2H read of placements, routing, flags and
scores inside the timed region.  `--impl reference` times the reference's CPU path (oracle port:
the reference's model restated in numpy + HiGHS standing in for the un-vendored OR-Tools/SCIP wheel).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_NODES, N_FUNCS = 50, 10
WORKLOAD = "C2: 50 nodes x 10 functions, random symmetric delays, min-delay objective (NeptuneMinDelay step 1)"


def make_hosts(batch, seed0):
    from neptune_mip_b200 import synth
    from neptune_mip_b200.core.utils import data_to_solver_input
    from neptune_mip_b200.device import InstanceBatch
    datas = [data_to_solver_input(synth.config_payload("C2", seed0 + s), 1, with_db=False) for s in range(batch)]
    return InstanceBatch.host_arrays(datas)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([t.strip() for t in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU path (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------
def _cpu_one(args):
    seed, limit = args
    sys.path.insert(0, ROOT)
    from neptune_mip_b200 import synth
    from neptune_mip_b200.core.utils import data_to_solver_input
    from oracle import mip, model
    t0 = time.time()
    a = model.arrays_from_data(data_to_solver_input(synth.config_payload("C2", seed), 1, with_db=False))
    out = mip.solve_step1(a, "min_delay", time_limit=limit)
    return dict(seed=seed, seconds=time.time() - t0, optimal=bool(out["optimal"]), objective=out["objective"],
                gap=out["gap"])


def cpu_reference(n_instances, limit, procs, seed0=0):
    """Wall time of `n_instances` C2 step-1 MIPs, `procs` worker processes (the reference forks up to 10
    workers, main.py:69), each solve capped at `limit` seconds (an unsolved instance counts as done at
    the cap, which can only flatter the CPU)."""
    from concurrent.futures import ProcessPoolExecutor
    t0 = time.time()
    with ProcessPoolExecutor(procs) as ex:
        recs = list(ex.map(_cpu_one, [(seed0 + s, limit) for s in range(n_instances)]))
    wall = time.time() - t0
    return wall, recs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 10))
    n_inst = args.ref_instances if args.ref_instances > 0 else max(procs, 4)
    limit = args.ref_limit
    walls = []
    recs = []
    for it in range(args.warmup + args.steps):
        wall, recs = cpu_reference(n_inst, limit, procs, seed0=1000 * it)
        if it >= args.warmup:
            walls.append(wall)
    ms = 1e3 * sum(walls) / max(len(walls), 1)
    value = n_inst / (ms / 1e3)
    solved = sum(r["optimal"] for r in recs)
    sample = (f"{n_inst} C2 instances per step, {procs} processes, HiGHS (stand-in for SCIP) capped at {limit:.0f} s "
              f"per instance; {solved}/{n_inst} proven optimal inside the cap in the last step "
              "(unsolved ones are counted as finished at the cap)")
    line = {"impl": "reference", "metric": "placement instances/sec", "value": value, "unit": "instances/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_step": n_inst, "time_limit_s": limit},
            "cpu_baseline": {"value": value, "unit": "instances/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from neptune_mip_b200 import _lib, device
    from neptune_mip_b200.batch import BatchParams, solve_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    B = args.batch
    prm = BatchParams(kind="min_delay", lp_iters=args.lp_iters, lp_check_every=args.lp_iters,
                      chains=args.chains, sweeps=args.sweeps, lp_path=args.lp_path)
    # instances are sharded across ranks by seed (weak scaling: B per GPU): rank r owns the contiguous block
    # sharding.shard_range(world*B, r, world) = [r*B, (r+1)*B) -- no data-path collective
    from neptune_mip_b200 import sharding
    lo, hi = sharding.shard_range(world * B, rank, world)
    host = make_hosts(hi - lo, lo)
    pinned = {k: torch.from_numpy(host[k]).pin_memory() for k in device.InstanceBatch.FIELDS}
    inst = device.InstanceBatch.from_host(host, pinned=pinned)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import ctypes
    cnt = ctypes.c_int64()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        lib.neptune_launch_count(ctypes.byref(cnt), 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = [fn() for _ in range(steps)]
        e1.record()
        barrier()
        lib.neptune_launch_count(ctypes.byref(cnt), 0)
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), outs, int(cnt.value)

    # ---- device-resident throughput ("value") + live roofline of the PDHG kernels --------------------
    pd = {"ms": 0.0, "iters": 0, "dims": (0, 0, 0), "bytes": 0, "path": ""}

    def step_resident():
        res = solve_batch(inst, prm, time_pdhg=True)
        pd["ms"] += res.pdhg_ms
        pd["iters"] += res.pdhg_iters
        pd["dims"] = res.model_dims
        pd["bytes"], pd["path"] = res.pdhg_bytes_per_iter, res.pdhg_path
        return res

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        solve_batch(inst, prm)
    pd.update(ms=0.0, iters=0)
    ms_total, outs, launches = timed(step_resident, args.steps, 0)
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    # ---- end to end from pinned host buffers through the batched plugin call -------------------------
    d2h = {}

    def step_e2e_pinned():
        ib = device.InstanceBatch(B=B, N=host["N"], F=host["F"], budget=host["budget"],
                                  **{k: pinned[k].to("cuda", non_blocking=True) for k in device.InstanceBatch.FIELDS})
        res = solve_batch(ib, prm)
        out = (res.c.cpu(), res.x.cpu(), res.flags.cpu(), res.scores.cpu())
        d2h["bytes"] = sum(t.numel() * t.element_size() for t in out)
        return out

    ms_e2e_total, _, _ = timed(step_e2e_pinned, args.steps, max(1, args.warmup // 2))
    ms_e2e = ms_e2e_total / args.steps
    e2e_value = world * B / (ms_e2e / 1e3)

    # ---- quality of what was computed (parity bar: feasible, objective vs the reference optimum) --------
    last = outs[-1]
    flags = last.flags.cpu().numpy()
    delay = last.scores[:, 0].cpu().numpy()
    quality = {"feasible": int((flags == 63).sum()), "instances": int(B)}
    gold_path = os.path.join(ROOT, "tests", "golden", "mip_optima.json")
    if rank == 0 and os.path.exists(gold_path):
        gold = {r["seed"]: r for r in json.load(open(gold_path)) if r["config"] == "C2" and r["optimal"]}
        gaps = [(delay[s] - gold[s]["objective"]) / abs(gold[s]["objective"]) for s in gold if s < B]
        if gaps:
            quality.update(reference_optima_known=len(gaps), max_rel_gap=float(max(gaps)),
                           within_1e4=int(sum(g <= 1e-4 for g in gaps)))
    if last.lp is not None:
        lp = last.lp
        # dual objective of the strengthened relaxation (x <= 1 stated, so it is finite): a lower bound on the
        # MIP optimum up to the remaining dual residual, which is reported next to it
        quality.update(lp_bound_mean=float(np.mean(lp["dual_obj"])), lp_primal_mean=float(np.mean(lp["primal_obj"])),
                       lp_dual_residual_mean=float(np.mean(lp["dual_res"])), lp_converged=int(lp["converged"].sum()),
                       mean_gap_to_lp_bound=float(np.mean((delay - lp["dual_obj"]) / np.maximum(np.abs(delay), 1e-9))))

    # ---- roofline of the dominant kernels (the PDHG iteration pair) -----------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    rows, cols, nnz = pd["dims"]
    X = N_FUNCS * N_NODES * N_NODES
    # algorithmic bytes of one PDHG iteration of the batch: matrix-free = the x-shaped
    # streams x, yS, xsum, ysum read and written once (64 B per x column) + the F*N-sized vectors + d
    bytes_iter = pd["bytes"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01f_pdhg_mf_traffic.json")
    if pd["path"] == "matrix-free" and os.path.exists(tpath):
        # dram__bytes_read+write of one iteration (k_mf_iter + k_mf_small) from the ncu --set full capture (B = 256)
        traffic = json.load(open(tpath))["per_instance_dram_bytes"] * B
    roof = None
    if pd["iters"]:
        ach = bytes_iter * pd["iters"] / (pd["ms"] / 1e3) / 1e9
        kernel = ("PDHG iteration (matrix-free) = k_mf_iter (one streaming pass over x, yS and their running sums) + "
                  "k_mf_small (F*N-sized vectors)" if pd["path"] == "matrix-free" else
                  "PDHG iteration = k_spmv_short/k_spmv_tasks<PrimalUpdate> over A^T + k_spmv_short/k_spmv_tasks<DualUpdate> over A")
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "kernel": kernel, "bytes_per_iteration": bytes_iter, "iterations_timed": pd["iters"], "pdhg_ms": pd["ms"],
                "us_per_iteration": 1e3 * pd["ms"] / pd["iters"], "pdhg_share_of_step": pd["ms"] / ms_total,
                "peak_source": peak_src,
                "note": "one iteration of the CSR solver on the same model moved %d bytes (%.1fx)" %
                        (B * (16 * nnz + 88 * cols + 72 * rows), B * (16 * nnz + 88 * cols + 72 * rows) / max(bytes_iter, 1))}

    line = {"metric": "placement instances/sec", "value": value, "unit": "instances/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "lp_iters": args.lp_iters, "ls_chains": args.chains,
                       "ls_sweeps": args.sweeps, "lp_path": pd["path"],
                       "l2": "PDHG working set of the batch (x, yS and their running sums: %d MB) exceeds the 126 MB L2"
                             % (B * 32 * X // 1000000) if B * 32 * X > 126e6 else
                             "working set smaller than L2: flush not applicable, see DESIGN.md"},
            "e2e": {"value": e2e_value, "unit": "instances/s", "h2d_bytes_per_step": inst.h2d_bytes(),
                    "d2h_bytes_per_step": d2h.get("bytes", 0), "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clocks, "quality": quality}
    if roof:
        line["roofline"] = roof
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = max(1, min(os.cpu_count() or 1, 10))
        n_inst = procs
        wall, recs = cpu_reference(n_inst, 15.0, procs)
        line["cpu_baseline"] = {"value": n_inst / wall, "unit": "instances/s", "cores": procs, "kind": "port",
                                "sample": f"{n_inst} C2 instances (seeds 0..{n_inst - 1}), oracle model + HiGHS capped at 15 s each, "
                                          f"{sum(r['optimal'] for r in recs)}/{n_inst} proven optimal inside the cap"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="instances per GPU per step")
    ap.add_argument("--lp-iters", type=int, default=2048)
    ap.add_argument("--chains", type=int, default=8)
    ap.add_argument("--sweeps", type=int, default=400)
    ap.add_argument("--lp-path", default="auto", choices=["auto", "csr"],
                    help="auto: matrix-free PDHG (default); csr: assemble the model and run the CSR solver (round-1 a-e path)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ref-limit", type=float, default=20.0, help="--impl reference: HiGHS cap per instance (s)")
    ap.add_argument("--ref-instances", type=int, default=0, help="--impl reference: instances per step (0 = one per worker)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
