"""GPU probe: one C5 / C2 instance through neptune_lns_search with and without the LP guide / prices."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neptune_mip_b200 import device, synth
from neptune_mip_b200.core.utils import data_to_solver_input
cfg, seed = sys.argv[1], int(sys.argv[2])
gold = {r["seed"]: r for r in json.load(open(os.path.join(ROOT, "tests/golden/mip_optima.json"))) if r["config"] == cfg}
opt = gold[seed]["objective"]
inst = device.InstanceBatch.from_datas([data_to_solver_input(synth.config_payload(cfg, seed), 1, with_db=False)])
lp = device.slot_relaxation(inst)
xs, ys, res = device.pdhg_mf_solve(lp, max_iters=40000, check_every=256, eps_rel=1e-6, eps_abs=1e-9)
N, F = inst.N, inst.F; X = F * N * N
guide = xs[:, X:X + F * N].contiguous(); lam0 = ys[:, 3 * F * N + N:3 * F * N + 2 * N].contiguous()
print("opt", opt, "lp", float(res[0]["dual_obj"]), "conv", int(res[0]["converged"]), "lam0>1e-6:", [(j, round(float(v), 3)) for j, v in enumerate(lam0[0].cpu()) if v > 1e-6])
for name, g_, l_ in (("guide+lam", guide, lam0), ("guide only", guide, None), ("nothing", None, None)):
    for noise in (0.1, 0.03):
        c, g, r = device.lns_search(inst, "min_delay", chains=32, rounds=3000, k=3, noise_coef=noise, rng_seed=1, guide=g_, lam0=l_)
        other = device.lns_search.last_other_bound
        pr = device.route_lp(inst, c.contiguous())
        val = torch.where(pr["status"] == 1, pr["obj"], torch.full_like(pr["obj"], float("inf")))[0].cpu().numpy()
        g = g[0].cpu().numpy(); o = other[0].cpu().numpy()
        ub, lb = slice(0, 32), slice(32, 64)
        print(f"{name:11s} noise {noise}: best exact {val.min() - opt:+.3f} | ub rec: U-opt min {g[ub].min() - opt:+.2f} (its g {o[ub][g[ub].argmin()] - opt:+.2f}, exact {val[ub][g[ub].argmin()] - opt:+.2f})"
              f" | lb rec: g-opt min {g[lb].min() - opt:+.2f} (its U {o[lb][g[lb].argmin()] - opt:+.2f}, exact {val[lb][g[lb].argmin()] - opt:+.2f}) status0 {int((pr['status'] != 1).sum())}")
if gold[seed].get("placement"):
    start = np.zeros((1, 1, F, N), np.uint8)
    for f, j in gold[seed]["placement"]:
        start[0, 0, f, j] = 1
    st = torch.from_numpy(start).cuda().contiguous()
    pr0 = device.route_lp(inst, st, want_x=True)
    print("golden placement: exact", float(pr0["obj"][0, 0]) - opt, "status", int(pr0["status"][0, 0]), "pods", int(start.sum()), "closed by route_lp", int((pr0["c_out"] != st).sum()), "pivots", pr0["info"][0, 0].tolist())
    for rounds in (0, 1, 50):
        c, g, r = device.lns_search(inst, "min_delay", chains=4, rounds=rounds, k=3, noise_coef=0.0, rng_seed=1, seeds_u8=st)
        o = device.lns_search.last_other_bound
        print("  from the optimum, rounds", rounds, "U-opt", (g[0, :4].cpu().numpy() - opt).round(3), "g-opt", (g[0, 4:].cpu().numpy() - opt).round(3), "other", (o[0].cpu().numpy() - opt).round(2), "hamming", [(int((c[0, q] != st[0, 0]).sum())) for q in range(8)])
    x = pr0["x"][0, 0].cpu().numpy(); cc = start[0, 0]
    load = np.einsum("ifj,fi,fj->j", x, inst.w[0].cpu().numpy(), inst.r[0].cpu().numpy())
    print("  loads at optimum (K = %g):" % float(inst.Kj[0, 0]), np.round(load, 2).tolist())
    share = x.sum(axis=0)
    print("  pods with share < 1:", [(f, j, round(float(share[f, j]), 3)) for f in range(F) for j in range(N) if cc[f, j] and share[f, j] < 1 - 1e-6])
