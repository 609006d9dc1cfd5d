"""GPU exploration: local-search quality vs the HiGHS optima in tests/golden/mip_optima.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import data_of
from neptune_mip_b200 import device, synth

gold = {(r["config"], r["seed"]): r for r in json.load(open(os.path.join(ROOT, "tests/golden/mip_optima.json")))}

def run(config, seeds, chains, sweeps, kind="min_delay"):
    payloads = [synth.config_payload(config, s) for s in seeds]
    inst = device.InstanceBatch.from_datas([data_of(p) for p in payloads])
    torch.cuda.synchronize(); t0 = time.time()
    c0, n0, info = device.efttc(inst, kind)
    seeds_t = c0[:, None].contiguous()
    best_c, best_obj, fl = device.local_search(inst, kind, seeds_t, chains=chains, sweeps=sweeps)
    best_c, x, n, obj, feas = device.route_capacitated(inst, best_c)
    flags, scores = device.check_solution(inst, x, device.u8_to_f64(best_c), n)
    torch.cuda.synchronize(); dt = time.time() - t0
    out = []
    for b, s in enumerate(seeds):
        g = gold.get((config, s))
        ref = g["objective"] if g and g["optimal"] else None
        sc = float(scores[b, 0]); 
        out.append((s, sc, ref, None if ref is None else (sc - ref) / max(1.0, abs(ref)), int(flags[b])))
    return dt, out

if __name__ == "__main__":
    for chains, sweeps in [(64, 200), (296, 400), (296, 1200)]:
        dt, out = run("C2", [0, 1, 2, 3, 4], chains, sweeps)
        print("C2", chains, sweeps, f"{dt:.2f}s", out, flush=True)
    seeds = [s for s in range(64) if ("C5", s) in gold]
    for chains, sweeps in [(8, 100), (32, 300)]:
        dt, out = run("C5", seeds, chains, sweeps)
        gaps = np.array([o[3] for o in out if o[3] is not None])
        print("C5", chains, sweeps, f"{dt:.2f}s n={len(out)} exact={(np.abs(gaps) <= 1e-4).sum()} worst={gaps.max():.3e} infeasible={(np.array([o[4] for o in out]) != 63).sum()}", flush=True)
