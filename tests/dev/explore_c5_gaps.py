import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import data_of, arrays_of
from neptune_mip_b200 import device, synth
from oracle import mip as omip, routing
gold = {(r["config"], r["seed"]): r for r in json.load(open(os.path.join(ROOT, "tests/golden/mip_optima.json")))}
seeds = list(range(64))
payloads = [synth.config_payload("C5", s) for s in seeds]
inst = device.InstanceBatch.from_datas([data_of(p) for p in payloads])
sd = torch.stack([device.efttc(inst, k)[0] for k in ("min_delay", "min_util", "min_delay_util")], dim=1).contiguous()
bc, bo, _ = device.local_search(inst, "min_delay", sd, chains=32, sweeps=300)
c2, x, n, obj, feas = device.route_capacitated(inst, bc)
fl, sc = device.check_solution(inst, x, device.u8_to_f64(c2), n)
for b in seeds:
    ref = gold[("C5", b)]["objective"]; got = float(sc[b, 0])
    if abs(got - ref) <= 1e-4 * abs(ref) and int(fl[b]) == 63: continue
    a = arrays_of(payloads[b])
    lp = routing.lp_routing(a, bc[b].cpu().numpy())
    m = omip.solve_step1(a, "min_delay")
    cm = torch.from_numpy((m["c"] > 0.5).astype(np.uint8)).cuda()[None].contiguous()
    ib = device.InstanceBatch.from_datas([data_of(payloads[b])])
    _, xm, nm, om, fm = device.route_capacitated(ib, cm)
    print(f"seed {b}: ref {ref:.4f} ours {got:.4f} flags {int(fl[b]):06b} feas {int(feas[b])} | LP(our c) {None if lp is None else round(lp[0],4)} | our route of MIP c {float(om[0]):.4f} feas {int(fm[0])}")
