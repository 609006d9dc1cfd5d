"""C3 (500 x 50) through the step-1 pipeline pieces, timed separately."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import data_of
from neptune_mip_b200 import device, synth
from neptune_mip_b200._lib import FLAG_STRENGTHEN
kind = sys.argv[1] if len(sys.argv) > 1 else "min_delay"
chains, sweeps = int(sys.argv[2]), int(sys.argv[3])
p = synth.config_payload("C3")
inst = device.InstanceBatch.from_datas([data_of(p)])
def T(msg, t0):
    torch.cuda.synchronize(); print(f"{msg}: {time.time()-t0:.2f}s", flush=True)
t0 = time.time(); seeds = []
for k in ("min_delay", "min_util", "min_delay_util"):
    t1 = time.time(); c, n, info = device.efttc(inst, k); T(f"efttc {k} iters={int(info[0,0])} pods={int(info[0,1])}", t1); seeds.append(c)
seeds = torch.stack(seeds, dim=1).contiguous()
t0 = time.time()
lp = device.assemble(inst, kind, flags=FLAG_STRENGTHEN)
xs, ys, res = device.pdhg_solve(lp, max_iters=int(sys.argv[4]) if len(sys.argv) > 4 else 2000, check_every=250, eps_rel=1e-4)
T(f"pdhg strengthened rows={lp.rows} cols={lp.cols} nnz={lp.nnz} -> {res[0]}", t0)
X = inst.F * inst.N * inst.N
guide = xs[:, X:X + inst.F * inst.N].contiguous(); del lp, xs, ys
t0 = time.time()
bc, bo, _ = device.local_search(inst, kind, seeds, chains=chains, sweeps=sweeps, guide=guide)
T(f"local search chains={chains} sweeps={sweeps} obj={float(bo[0])}", t0)
t0 = time.time()
c2, x, n, obj, feas = device.route_capacitated(inst, bc)
flags, scores = device.check_solution(inst, x, device.u8_to_f64(c2), n)
T(f"route+check obj={float(obj[0])} feas={int(feas[0])} flags={int(flags[0]):06b} scores={scores.cpu().numpy()[0]} pods={int(c2.sum())}", t0)
for i, k in enumerate(("min_delay", "min_util", "min_delay_util")):
    cs, xs_, ns, o, f = device.route_capacitated(inst, seeds[:, i].contiguous())
    fl, sc = device.check_solution(inst, xs_, device.u8_to_f64(cs), ns)
    print(f"  seed efttc {k}: delay={float(sc[0,0]):.1f} util={float(sc[0,1])} flags={int(fl[0]):06b} pods={int(cs.sum())}")
