import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import data_of
from neptune_mip_b200 import device, synth
for cores in (200, 100000):
    payloads = [synth.random_payload(50, 10, s, node_cores=cores) for s in range(64)]
    inst = device.InstanceBatch.from_datas([data_of(p) for p in payloads])
    sd = torch.stack([device.efttc(inst, k)[0] for k in ("min_delay", "min_util", "min_delay_util")], dim=1).contiguous()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        bc, bo, _ = device.local_search(inst, "min_delay", sd, chains=8, sweeps=240)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(f"node_cores={cores}: 64 instances x 8 chains x 240 sweeps: {dt*1e3:.0f} ms  mean obj {float(bo.mean()):.1f}")
