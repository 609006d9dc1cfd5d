import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import data_of, arrays_of
from neptune_mip_b200 import device, synth
from oracle import checkers, mip as omip, routing
gold = json.load(open(os.path.join(ROOT, "tests/golden/random_small.json")))
for rec in gold:
    want = rec["neptune"]["NeptuneMinDelay"]["score"]["step1"]
    p = synth.random_payload(rec["N"], rec["F"], rec["seed"], node_cores=rec["node_cores"])
    a = arrays_of(p)
    inst = device.InstanceBatch.from_datas([data_of(p)])
    seeds = torch.stack([device.efttc(inst, k)[0] for k in ("min_delay", "min_util", "min_delay_util")], dim=1).contiguous()
    bc, bo, _ = device.local_search(inst, "min_delay", seeds, chains=64, sweeps=300)
    c2, x, n, obj, feas = device.route_capacitated(inst, bc)
    flags, scores = device.check_solution(inst, x, device.u8_to_f64(c2), n)
    # exact LP routing value of OUR placement, and our routing of the MIP's placement
    lp_ours = routing.lp_routing(a, bc[0].cpu().numpy())
    ref = omip.solve_step1(a, "min_delay")
    cm = torch.from_numpy((ref["c"] > 0.5).astype(np.uint8)).cuda()[None].contiguous()
    _, xm, nm, objm, feasm = device.route_capacitated(inst, cm)
    print(rec["N"], rec["F"], rec["seed"], "want", want, "ls", float(bo[0]), "route", float(obj[0]), "flags", int(flags[0]),
          "| LP(our c)", None if lp_ours is None else round(lp_ours[0], 4), "| our route of MIP c", float(objm[0]), int(feasm[0]))
