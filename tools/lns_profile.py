"""GPU probe for profiling k_lns at the bench shape: 32 C2 instances x 96 chains, a few thousand rounds.
usage: python tools/lns_profile.py [rounds] [block_mode] [k]   (run under ncu with -k regex:k_lns -c 1)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neptune_mip_b200 import device, synth
from neptune_mip_b200.core.utils import data_to_solver_input
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
k = int(sys.argv[3]) if len(sys.argv) > 3 else 3
B = int(os.environ.get("LNS_B", "32")); chains = int(os.environ.get("LNS_CHAINS", "96"))
inst = device.InstanceBatch.from_datas([data_to_solver_input(synth.config_payload("C2", s), 1, with_db=False) for s in range(B)])
lp = device.slot_relaxation(inst)
xs, ys, res = device.pdhg_mf_solve(lp, max_iters=2048, check_every=256)
N, F = inst.N, inst.F; X = F * N * N
guide = xs[:, X:X + F * N].contiguous(); lam0 = ys[:, 3 * F * N + N:3 * F * N + 2 * N].contiguous()
device.lns_block_mode(mode)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    c, g, r = device.lns_search(inst, "min_delay", chains=chains, rounds=rounds, k=k, noise_coef=0.1, rng_seed=1, guide=guide, lam0=lam0)
    e1.record(); e1.synchronize()
    print(f"B {B} chains {chains} rounds {rounds} k {k} mode {mode}: {e0.elapsed_time(e1):.1f} ms  "
          f"({1e3 * e0.elapsed_time(e1) / (B * chains * max(rounds, 1)):.3f} us per chain-round)  best {float(g[0].min()):.3f}")
if os.environ.get("LNS_PER_INSTANCE"):
    # the same chains, one instance per launch (12 blocks on 12 SMs): how much do the instances differ?
    ts = []
    for b in range(B):
        one = device.InstanceBatch(B=1, N=N, F=F, budget=inst.budget, **{kk: getattr(inst, kk)[b:b + 1].contiguous() for kk in device.InstanceBatch.FIELDS})
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        device.lns_search(one, "min_delay", chains=chains, rounds=rounds, k=k, noise_coef=0.1, rng_seed=1, guide=guide[b:b + 1].contiguous(), lam0=lam0[b:b + 1].contiguous())
        e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    print("per-instance ms (one launch each):", [round(t, 1) for t in ts], "max/mean", round(max(ts) / (sum(ts) / len(ts)), 2))
