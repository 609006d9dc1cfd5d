import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import data_of
from neptune_mip_b200 import device, synth
from neptune_mip_b200.core.solvers.neptune.neptune_step2 import disruption
p = synth.random_payload(20, 5, 1, node_cores=100)
data = data_of(p)
inst = device.InstanceBatch.from_datas([data])
seeds = torch.stack([device.efttc(inst, k)[0] for k in ("min_delay", "min_util", "min_delay_util")], dim=1).contiguous()
bc, bo, _ = device.local_search(inst, "min_delay", seeds, chains=64, sweeps=300)
c1, x1, n1, obj1, feas1 = device.route_capacitated(inst, bc)
print("step1", float(bo[0]), float(obj1[0]), int(feas1[0]), int(c1.sum()))
old = (np.asarray(data.old_allocations_matrix) > 0).astype(np.uint8)
s2 = torch.from_numpy(np.stack([c1[0].cpu().numpy(), old])[None]).cuda().contiguous()
bound = torch.tensor([1.3 * float(obj1[0])], dtype=torch.float64, device="cuda")
for mode in ("delete", "create"):
    b2, o2, _ = device.disruption_search(inst, "min_delay", mode, bound, s2, chains=64, sweeps=300)
    c2, x2, n2, obj2, feas2 = device.route_capacitated(inst, b2)
    fl, sc = device.check_solution(inst, x2, device.u8_to_f64(c2), n2)
    print(mode, "ls obj", float(o2[0]), "pods", int(b2.sum()), "->", int(c2.sum()), "route delay", float(obj2[0]), "bound", float(bound[0]),
          "feas", int(feas2[0]), "flags", bin(int(fl[0])), "disruption", disruption(c2[0].cpu().numpy(), old, mode))
