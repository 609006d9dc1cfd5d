"""GPU tool: step-1 quality of the batched solve path on the C2 / C5 seeds with known optima
(tests/golden/mip_optima.json).   python tools/c2_quality.py [--config C2] [--chains 32] [--rounds 6000] ..."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neptune_mip_b200 import device, synth  # noqa: E402
from neptune_mip_b200.batch import BatchParams, solve_batch  # noqa: E402
from neptune_mip_b200.core.utils import data_to_solver_input  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--seeds", type=int, default=16)
    ap.add_argument("--chains", type=int, default=32)
    ap.add_argument("--rounds", type=int, default=6000)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--noise", type=float, default=0.06)
    ap.add_argument("--elites", type=int, default=16)
    ap.add_argument("--phases", type=int, default=1)
    ap.add_argument("--cooling", type=float, default=0.6)
    ap.add_argument("--pool", type=int, default=16)
    ap.add_argument("--local", type=int, default=0)
    ap.add_argument("--dump", default="")
    ap.add_argument("--k4", type=int, default=0)
    ap.add_argument("--final-k4", type=int, default=0)
    ap.add_argument("--final-noise", type=float, default=0.15)
    ap.add_argument("--lp-iters", type=int, default=20000)
    ap.add_argument("--no-cut", action="store_true")
    ap.add_argument("--search", default="auto")
    ap.add_argument("--rng", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=1)
    a = ap.parse_args()
    gold = {r["seed"]: r for r in json.load(open(os.path.join(ROOT, "tests", "golden", "mip_optima.json")))
            if r["config"] == a.config and r["optimal"]}
    seeds = [s for s in range(a.seeds) if s in gold]
    datas = [data_to_solver_input(synth.config_payload(a.config, s), 1, with_db=False) for s in seeds]
    inst = device.InstanceBatch.from_datas(datas)
    prm = BatchParams(lp_iters=a.lp_iters, lp_check_every=256, lns_chains=a.chains, lns_rounds=a.rounds, lns_k=a.k,
                      lns_noise=a.noise, elites=a.elites, lns_phases=a.phases, lns_cooling=a.cooling, lns_restart_pool=a.pool, lns_local_chains=a.local, lns_k4_chains=a.k4, lns_final_k4=a.final_k4, lns_final_noise=a.final_noise, lp_cut=not a.no_cut, search=a.search, rng_seed=a.rng,
                      chains=16, sweeps=400)
    for rep in range(a.repeat):
        prm.rng_seed = a.rng + rep
        torch.cuda.synchronize(); t0 = time.time()
        res = solve_batch(inst, prm, time_pdhg=True)
        torch.cuda.synchronize(); dt = time.time() - t0
        sc = res.scores[:, 0].cpu().numpy(); fl = res.flags.cpu().numpy()
        gaps = np.array([(sc[k] - gold[s]["objective"]) / gold[s]["objective"] for k, s in enumerate(seeds)])
        lp = res.lp
        print(json.dumps({"config": a.config, "B": len(seeds), "search": res.search_path, "chains": a.chains, "rounds": a.rounds,
                          "k": a.k, "noise": a.noise, "phases": a.phases, "cooling": a.cooling, "wall_s": round(dt, 3), "pdhg_ms": round(res.pdhg_ms, 1),
                          "lns_ms": round(res.lns_ms, 1), "feasible": int((fl == 63).sum()),
                          "within_1e4": int((gaps <= 1e-4).sum()), "max_gap": float(gaps.max()),
                          "gaps": [float(f"{g:.2e}") for g in gaps],
                          "lp_bound_gap": [float(f"{(gold[s]['objective'] - lp[k]['dual_obj']) / gold[s]['objective']:.2e}") for k, s in enumerate(seeds)] if lp is not None else None,
                          "lp_iters": [int(v) for v in lp["iters"]] if lp is not None else None,
                          "lp_converged": int(lp["converged"].sum()) if lp is not None else None,
                          "elite_g_minus_opt": [[round(float(v) - gold[s]["objective"], 2) for v in res.lns_diag["elite_g"][k].cpu()[[0, 1, res.lns_diag["n_upper"], res.lns_diag["n_upper"] + 1]]] for k, s in enumerate(seeds)] if res.lns_diag else None,
                          "elite_val_minus_opt": [[round(float(v) - gold[s]["objective"], 2) for v in res.lns_diag["elite_val"][k].cpu()] for k, s in enumerate(seeds)] if res.lns_diag else None,
                          "bad": {str(seeds[k]): {"gap": float(f"{gaps[k]:.3g}"), "status": res.lns_diag["status"][k].cpu().tolist()[:8], "fell_back": bool(res.lns_diag["fell_back"][k]),
                                                  "g": [round(float(v), 1) for v in res.lns_diag["elite_g"][k].cpu()[:4]], "opt": gold[seeds[k]]["objective"]}
                                  for k in range(len(seeds)) if gaps[k] > 1e-2} if res.lns_diag else None,
                          "lb_records": [[(round(float(res.lns_diag["elite_g"][k, res.lns_diag["n_upper"] + q]) - gold[s]["objective"], 2),
                                           round(float(res.lns_diag["lb_other"][k, q]) - gold[s]["objective"], 2),
                                           round(float(res.lns_diag["elite_val"][k, res.lns_diag["n_upper"] + q]) - gold[s]["objective"], 2),
                                           int(res.lns_diag["pivots"][k, res.lns_diag["n_upper"] + q])) for q in range(3)] for k, s in enumerate(seeds)] if res.lns_diag and len(seeds) <= 16 else None,
                          "rounds_of_best": res.lns_round.cpu().tolist() if res.lns_round is not None else None}))
        sys.stdout.flush()
        if a.dump and res.lns_diag:
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", a.dump), elite_c=res.lns_diag["elite_c"].cpu().numpy(),
                                elite_g=res.lns_diag["elite_g"].cpu().numpy(), elite_val=res.lns_diag["elite_val"].cpu().numpy(),
                                n_upper=res.lns_diag["n_upper"], seeds=np.array(seeds))


if __name__ == "__main__":
    main()
