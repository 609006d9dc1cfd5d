"""GPU probe (uses the oracle: tools only): a small CPU-starved instance -- HiGHS optimum, then the search from the
optimum and from scratch, records and exact values."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neptune_mip_b200 import device, synth
from neptune_mip_b200.core.utils import data_to_solver_input
from oracle import mip, model
N, F, seed, cores = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
pl = synth.random_payload(N, F, seed, node_cores=cores)
data = data_to_solver_input(pl, 1, with_db=False)
a = model.arrays_from_data(data)
o = mip.solve_step1(a, "min_delay", 0.5)
opt = o["objective"]
inst = device.InstanceBatch.from_datas([data])
st = torch.from_numpy((o["c"] > 0.5).astype(np.uint8)[None, None]).cuda().contiguous()
pr0 = device.route_lp(inst, st)
print("opt", opt, "route_lp(c*) - opt", float(pr0["obj"][0, 0]) - opt, "status", int(pr0["status"][0, 0]), "info", pr0["info"][0, 0].tolist())
for rounds in (0, 1, 200):
    c, g, r = device.lns_search(inst, "min_delay", chains=4, rounds=rounds, k=3, noise_coef=0.0, rng_seed=1, seeds_u8=st)
    ob = device.lns_search.last_other_bound
    print(" from c*, rounds", rounds, "ub rec U-opt", (g[0, :4].cpu().numpy() - opt).round(3), "lb rec g-opt", (g[0, 4:].cpu().numpy() - opt).round(3), "rounds", r[0].cpu().tolist())
c, g, r = device.lns_search(inst, "min_delay", chains=64, rounds=3000, k=3, noise_coef=0.1, rng_seed=1)
pr = device.route_lp(inst, c.contiguous())
val = torch.where(pr["status"] == 1, pr["obj"], torch.full_like(pr["obj"], float("inf")))[0].cpu().numpy()
g = g[0].cpu().numpy()
print(" from scratch: best exact - opt", val.min() - opt, "| best U-opt", g[:64].min() - opt, "| best g-opt", g[64:].min() - opt, "| records without value", int(np.isinf(g).sum()), "of", g.size, "| status0", int((pr["status"] != 1).sum()))
