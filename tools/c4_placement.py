#!/usr/bin/env python
"""BASELINE config 4 (2000 nodes x 200 functions) on ONE GPU: a checked placement and the bracket around it.

  EFTTC placement (k_efttc) -> nearest-pod routing (k_route) -> the reference's checkers (k_check)
  -> lower bound of the slot-cut LP relaxation after a bounded number of matrix-free PDHG iterations (the dual
     objective with the box terms is a valid bound at every iterate).

The add/drop/swap search and k_lns keep a chain's state in shared memory and stop at N = 768 / N = 128, so at this
size the placement is EFTTC's; the record says so.  Prints one JSON object (also used by bench.py's `c4_placement`).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def record(n_nodes=2000, n_funcs=200, lp_iters=512, seed=0):
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

    ms_e, (c, n_e, info) = timed(lambda: device.efttc(inst, "min_delay"))
    # EFTTC's own routing is "every source to its nearest pod" with a global CPU check after every cycle
    # (efttc_step1.py:196-212, utils/constraints_step1.py:70-80), so the nearest routing of its placement is CPU-feasible
    import numpy as np
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"c4_efttc_c_{n_nodes}x{n_funcs}.npy"), c[0].cpu().numpy())
    ms_r, (x, n) = timed(lambda: device.route_placements(inst, c))
    c2 = c
    ms_c, (flags, scores) = timed(lambda: device.check_solution(inst, x, device.u8_to_f64(c2), n))
    fl = int(flags.cpu()[0])
    obj = scores[:, 0]
    feas = torch.tensor([1 if (fl & (OK_HANDLE | OK_MEMORY | OK_CPU | OK_C_X | OK_N_C)) == (OK_HANDLE | OK_MEMORY | OK_CPU | OK_C_X | OK_N_C) else 0])
    names = {"handle_all_requests": OK_HANDLE, "memory": OK_MEMORY, "cpu": OK_CPU, "c_according_to_x": OK_C_X,
             "n_according_to_c": OK_N_C}
    rec = {"workload": f"C4: {n_nodes} nodes x {n_funcs} functions, one instance, one GPU",
           "placement": "EFTTC (k_efttc) + nearest-pod routing (k_route), as EFTTC routes; the searches stop at N = 768 / 128",
           "pods": int(c2.sum().item()), "functions_without_pod": int((c2[0].sum(dim=1) == 0).sum().item()), "objective_min_delay": float(obj.cpu()[0]),
           "feasible": bool(int(feas.cpu()[0])), "checkers": {k: bool(fl & v) for k, v in names.items()},
           "ms": {"host_instance_build": 1e3 * t_build, "efttc": ms_e, "routing": ms_r, "checkers": ms_c},
           "x_bytes": int(x.numel() * 8)}
    del x
    if lp_iters > 0:
        lp = device.slot_relaxation(inst)
        ms_lp, (_, _, sol) = timed(lambda: device.pdhg_mf_solve(lp, max_iters=lp_iters, check_every=min(lp_iters, 256)))
        rec["lp"] = {"iterations": int(sol["iters"][0]), "ms": ms_lp, "us_per_iteration": 1e3 * ms_lp / max(int(sol["iters"][0]), 1),
                     "dual_bound": float(sol["dual_obj"][0]), "primal_obj": float(sol["primal_obj"][0]),
                     "converged": bool(sol["converged"][0]),
                     "note": "slot-cut relaxation; the dual objective (box terms included) is a valid lower bound at every iterate"}
        if rec["lp"]["dual_bound"] > 0:
            rec["gap_to_lp_bound"] = (rec["objective_min_delay"] - rec["lp"]["dual_bound"]) / rec["objective_min_delay"]
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2000)
    ap.add_argument("--funcs", type=int, default=200)
    ap.add_argument("--lp-iters", type=int, default=512)
    a = ap.parse_args()
    print(json.dumps(record(a.nodes, a.funcs, a.lp_iters)))
