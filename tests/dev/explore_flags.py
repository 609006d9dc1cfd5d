import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import data_of, arrays_of
from neptune_mip_b200 import device, synth
from oracle import checkers
seeds = list(range(64))
payloads = [synth.config_payload("C5", s) for s in seeds]
inst = device.InstanceBatch.from_datas([data_of(p) for p in payloads])
c0, n0, info = device.efttc(inst, "min_delay")
best_c, best_obj, fl = device.local_search(inst, "min_delay", c0[:, None].contiguous(), chains=32, sweeps=300)
c2, x, n, obj, feas = device.route_capacitated(inst, best_c)
flags, scores = device.check_solution(inst, x, device.u8_to_f64(c2), n)
flags = flags.cpu().numpy(); feas = feas.cpu().numpy()
for b in np.flatnonzero(flags != 63):
    a = arrays_of(payloads[b])
    xb = x[b].cpu().numpy(); cb = c2[b].cpu().numpy()
    load = checkers.cpu_load(a, xb)
    recv = xb.sum(axis=0)
    print(b, "flags", bin(flags[b]), "feas", feas[b], "ls_obj", float(best_obj[b]), "route_obj", float(obj[b]), "max over", (load - a["Kj"]).max(),
          "partial pods", [(f, j, round(recv[f, j], 4)) for f, j in zip(*np.nonzero((cb > 0) & (recv < 1 - 1e-6)))][:4], "rowsum dev", np.abs(xb.sum(axis=2) - 1).max())
