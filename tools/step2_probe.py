"""GPU tool: step-1 / step-2 scores of the Neptune classes on the 20x5 reference runs (tests/golden/random_small.json)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neptune_mip_b200 import synth
from neptune_mip_b200.server import solve_payload
gold = [r for r in json.load(open(os.path.join(ROOT, "tests/golden/random_small.json"))) if r["N"] == 20]
for rec in gold:
    want = rec["neptune"]["NeptuneMinDelay"]["score"]
    p = synth.random_payload(20, 5, rec["seed"], node_cores=100, solver_type="NeptuneMinDelay", args={"verbose": False, "chains": 64, "sweeps": 300})
    resp = solve_payload(p)
    print("seed", rec["seed"], "step1", resp["score"]["step1"], "want", want["step1"], "| step2", resp["score"]["step2"], "want", want["step2"], "| time %.2f s" % resp["processing_time"])
