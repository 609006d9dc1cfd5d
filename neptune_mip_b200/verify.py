"""Offline verifier: re-score a stored response and re-run the six constraint checks on the device.

The counterpart of the reference's `testing/*/..._score_analysis.py` (`recreate_all_vars_from_json`
+ `efttc/utils/objectives.py` + six checkers of `efttc/utils/constraints_step1.py`,
`simulated_score_analysis.py:25-74,308-320`), without the plotting: the response's rounded routing
(3 decimals, `output.py:30`) and allocations are turned back into dense x / c / n and handed to
`neptune_check_solution`.  Returns plain dicts (the tables of the reference's PDF reports)."""
from __future__ import annotations

import json
import sys

import numpy as np
import torch

from . import device
from ._lib import FLAG_NAMES
from .core import data_to_solver_input


def dense_from_response(payload: dict, response: dict):
    nodes, funcs = payload["node_names"], payload["function_names"]
    ni = {n: i for i, n in enumerate(nodes)}
    fi = {f: i for i, f in enumerate(funcs)}
    N, F = len(nodes), len(funcs)
    x = np.zeros((N, F, N))
    c = np.zeros((F, N))
    for src, per_f in response.get("cpu_routing_rules", {}).items():
        for f, per_d in per_f.items():
            for dst, v in per_d.items():
                x[ni[src], fi[f], ni[dst]] = round(float(v), 6)      # simulated_score_analysis.py:68
    for f, per_n in response.get("cpu_allocations", {}).items():
        for n, on in per_n.items():
            if on:
                c[fi[f], ni[n]] = 1.0
    return x, c, (c.sum(axis=0) > 0).astype(np.float64)


def verify(payload: dict, response: dict, alpha=None) -> dict:
    data = data_to_solver_input(payload, with_db=False, workload_coeff=payload.get("workload_coeff", 1))
    if alpha is None:
        alpha = payload.get("solver", {}).get("args", {}).get("alpha", 0.5)
    inst = device.InstanceBatch.from_datas([data])
    x, c, n = dense_from_response(payload, response)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()[None]  # noqa: E731
    flags, scores = device.check_solution(inst, t(x), t(c), t(n), float(alpha))
    fl = int(flags.cpu()[0])
    sc = scores.cpu().numpy()[0]
    return {"constraints": {name: bool(fl >> k & 1) for k, name in enumerate(FLAG_NAMES)},
            "scores": {"min_delay": float(sc[0]), "min_utilization": int(sc[1]), "min_delay_and_utilization": float(sc[2])},
            "pods": int(c.sum()), "active_nodes": int(n.sum()),
            "reported_score": response.get("score"), "processing_time": response.get("processing_time")}


if __name__ == "__main__":        # python -m neptune_mip_b200.verify stored_output.json  (file embeds "input")
    stored = json.load(open(sys.argv[1]))
    print(json.dumps(verify(stored["input"], stored), indent=1))
