"""GPU probe: the CPU-starved small instances of tests/golden/random_small.json through the drop-in solver class --
gap to the reference optimum and time per request."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neptune_mip_b200 import synth
from neptune_mip_b200.core import check_input, data_to_solver_input
import neptune_mip_b200.core.solvers as S


def serve(payload):
    check_input(payload)
    cfg = payload["solver"]
    solver = getattr(S, cfg["type"])(**cfg.get("args", {}))
    solver.load_data(data_to_solver_input(payload, with_db=payload.get("with_db", True), workload_coeff=payload.get("workload_coeff", 1)))
    solved = solver.solve()
    solver.results()
    return {"score": solver.score()}, solver, solved


gold = json.load(open(os.path.join(ROOT, "tests", "golden", "random_small.json")))
only = os.environ.get('SMALL_ONLY')
for rec in ([gold[int(only)]] if only else gold[:int(os.environ.get('SMALL_MAX', '99'))]):
    want = rec["neptune"]["NeptuneMinDelay"]["score"]["step1"]
    payload = synth.random_payload(rec["N"], rec["F"], rec["seed"], node_cores=rec["node_cores"], solver_type="NeptuneMinDelay",
                                   args={"verbose": False, "chains": 64, "sweeps": 300})
    t0 = time.time()
    resp, solver, _ = serve(payload)
    got = resp["score"]["step1"]
    print(rec["N"], rec["F"], rec["seed"], rec["node_cores"], "want", round(want, 4), "got", round(got, 4), "gap", f"{(got - want) / want:.2e}",
          "feasible", solver.step1.feasible, "t", round(time.time() - t0, 2))
