import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_solvers_gpu import serve
from neptune_mip_b200 import synth
gold = json.load(open(os.path.join(ROOT, "tests/golden/alibaba_case0.json")))
for name, payload in [("C1", synth.test_py_payload()),
                      ("alibaba MinDelay", dict(gold["input"], solver={"type": "NeptuneMinDelay", "args": {"verbose": False, "chains": 64, "sweeps": 200}})),
                      ("alibaba MinUtil", dict(gold["input"], solver={"type": "NeptuneMinUtilization", "args": {"verbose": False, "chains": 64, "sweeps": 200}})),
                      ("alibaba Combined", dict(gold["input"], solver={"type": "NeptuneMinDelayAndUtilization", "args": {"alpha": 0.5, "verbose": False, "chains": 64, "sweeps": 200}})),
                      ("C2 seed0", synth.config_payload("C2", 0, args={"verbose": False, "chains": 148, "sweeps": 400})),
                      ("r20x5 s2", synth.random_payload(20, 5, 2, node_cores=100, args={"verbose": False, "chains": 64, "sweeps": 300}))]:
    t0 = time.time()
    resp, solver, solved = serve(payload)
    torch.cuda.synchronize()
    print(name, resp["score"], "solved", solved, "pods", sum(len(v) for v in resp["cpu_allocations"].values()), f"{time.time()-t0:.2f}s", flush=True)
