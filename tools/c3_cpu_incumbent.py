#!/usr/bin/env python
"""the CPU incumbent at config 3 (500 nodes x 50 functions), attempted and logged.

Builds the reference's step-1 min-delay MIP (oracle/model.py: 12 525 000 columns, 76 000 rows, 50 M non-zeros) and
hands it to HiGHS (scipy.optimize.milp) with a stated wall-clock limit.  Whatever comes back -- an incumbent, a bound
or neither -- is written to profiles/ as the record of the attempt (DNF is recorded as DNF).  CPU only; run in the build
container:  python tools/c3_cpu_incumbent.py --limit 600
"""
import argparse
import json
import os
import resource
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--limit", type=float, default=600.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_c3_cpu_incumbent.json"))
    a = ap.parse_args()
    from neptune_mip_b200 import synth
    from neptune_mip_b200.core.utils import data_to_solver_input
    from oracle import mip, model
    t0 = time.time()
    arr = model.arrays_from_data(data_to_solver_input(synth.config_payload(a.config, 0), 1, with_db=False))
    rec = {"config": a.config, "N": int(arr["N"]), "F": int(arr["F"]), "solver": "HiGHS via scipy.optimize.milp",
           "time_limit_s": a.limit, "threads": os.cpu_count()}
    try:
        o = mip.solve_step1(arr, "min_delay", 0.5, time_limit=a.limit)
        rec.update({"objective": o["objective"], "optimal": bool(o["optimal"]), "dual_bound": o["dual_bound"], "gap": o["gap"],
                    "solve_seconds": o["seconds"], "incumbent": o["sol"] is not None,
                    "pods": int((o["c"] > 0.5).sum()) if o["sol"] is not None else None})
    except MemoryError as e:
        rec.update({"incumbent": False, "error": f"MemoryError: {e}"})
    except Exception as e:  # recorded, not hidden
        rec.update({"incumbent": False, "error": f"{type(e).__name__}: {e}"})
    rec["wall_seconds_with_model_build"] = time.time() - t0
    rec["peak_rss_gb"] = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6
    rec["verdict"] = ("optimal" if rec.get("optimal") else "incumbent at the limit" if rec.get("incumbent") else "DNF: no incumbent within the limit")
    json.dump(rec, open(a.out, "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
