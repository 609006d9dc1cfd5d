#!/usr/bin/env python
"""BASELINE config 4 (2000 nodes x 200 functions) on ONE GPU: a checked placement and the bracket around it.

  greedy placement (k_site_*, csrc/site.cu) -> two-choice routing (k_tc_*) -> the reference's checkers (k_check)
  -> lower bound of the slot-cut LP relaxation after a bounded number of matrix-free PDHG iterations (the dual
     objective with the box terms is a valid bound at every iterate).

The add/drop/swap search and k_lns keep a chain's state in shared memory and stop at N = 768 / N = 128, so at this
size the placement is the greedy one; the record says so.  Prints one JSON object.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def record(n_nodes=2000, n_funcs=200, lp_iters=512, seed=0, efttc=False):
    import torch
    from neptune_mip_b200 import device, synth
    from neptune_mip_b200._lib import OK_C_X, OK_CPU, OK_HANDLE, OK_MEMORY, OK_N_C
    from neptune_mip_b200.core.utils import data_to_solver_input

    t0 = time.time()
    data = data_to_solver_input(synth.random_payload(n_nodes, n_funcs, seed, node_cores=None), 1, with_db=False)
    inst = device.InstanceBatch.from_datas([data])
    t_build = time.time() - t0

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); out = fn(); e1.record(); e1.synchronize()
        return e0.elapsed_time(e1), out

    ms_e, (c, info) = timed(lambda: device.site_greedy(inst))
    ms_r, (c2, x, n, obj, feas, route_iters) = timed(lambda: device.route_two_choice(inst, c))
    ms_c, (flags, scores) = timed(lambda: device.check_solution(inst, x, device.u8_to_f64(c2), n))
    fl = int(flags.cpu()[0])
    names = {"handle_all_requests": OK_HANDLE, "memory": OK_MEMORY, "cpu": OK_CPU, "c_according_to_x": OK_C_X,
             "n_according_to_c": OK_N_C}
    near_x, _ = device.route_placements(inst, c)
    near_obj = float(device.check_solution(inst, near_x, device.u8_to_f64(c), n)[1].cpu()[0, 0])
    del near_x
    rec = {"workload": f"C4: {n_nodes} nodes x {n_funcs} functions, one instance, one GPU",
           "placement": "round-robin delay-improvement greedy (k_site_*: EFTTC's move, every function proposes per round) + "
                        "two-choice routing (k_tc_*: nearest / second-nearest pod, shares lowered until the CPU rows hold); the searches stop at N = 768 / 128",
           "greedy_rounds": int(info.cpu()[0, 0]), "routing_iterations": route_iters, "pods": int(c2.sum().item()),
           "functions_without_pod": int((c2[0].sum(dim=1) == 0).sum().item()),
           "objective_min_delay": float(obj.cpu()[0]), "objective_if_every_source_took_its_nearest_pod": near_obj,
           "feasible": bool(int(feas.cpu()[0])) and all(bool(fl & v) for v in names.values()),
           "checkers": {k: bool(fl & v) for k, v in names.items()},
           "ms": {"host_instance_build": 1e3 * t_build, "greedy": ms_e, "routing": ms_r, "checkers": ms_c},
           "x_bytes": int(x.numel() * 8)}
    if efttc:
        ms_ef, (ce, _, _) = timed(lambda: device.efttc(inst, "min_delay"))
        rec["efttc_for_comparison"] = {"ms": ms_ef, "pods": int(ce.sum().item()), "functions_without_pod": int((ce[0].sum(dim=1) == 0).sum().item()),
                                       "note": "one block per instance; its trading cycles leave most functions without a pod at this size"}
    del x
    if lp_iters > 0:
        lp = device.slot_relaxation(inst)
        ms_lp, (_, _, sol) = timed(lambda: device.pdhg_mf_solve(lp, max_iters=lp_iters, check_every=min(lp_iters, 256)))
        rec["lp"] = {"iterations": int(sol["iters"][0]), "ms": ms_lp, "us_per_iteration": 1e3 * ms_lp / max(int(sol["iters"][0]), 1),
                     "dual_bound": float(sol["dual_obj"][0]), "primal_obj": float(sol["primal_obj"][0]),
                     "converged": bool(sol["converged"][0]),
                     "note": "slot-cut relaxation; the dual objective (box terms included) is a valid lower bound at every iterate; "
                             "us_per_iteration is the whole solver call divided by its iterations: at this length about a fifth "
                             "of it is the solve's set-up (memsets and copies of the 25.6 GB vectors) and its KKT passes "
                             "(profiles/r02c_small_phases.log)"}
        if rec["lp"]["dual_bound"] > 0:
            rec["gap_to_lp_bound"] = (rec["objective_min_delay"] - rec["lp"]["dual_bound"]) / rec["objective_min_delay"]
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2000)
    ap.add_argument("--funcs", type=int, default=200)
    ap.add_argument("--lp-iters", type=int, default=512)
    ap.add_argument("--efttc", action="store_true", help="also time the EFTTC kernel at this size (minutes)")
    a = ap.parse_args()
    print(json.dumps(record(a.nodes, a.funcs, a.lp_iters, efttc=a.efttc)))
