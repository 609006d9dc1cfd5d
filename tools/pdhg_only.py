"""The matrix-free PDHG alone on the slot-cut relaxation (ncu target / timing): python tools/pdhg_only.py [B] [iters] [C2|C3|C5]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from neptune_mip_b200 import device
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 256
shape = sys.argv[3] if len(sys.argv) > 3 else "C2"
inst = device.InstanceBatch.from_host(bench.make_hosts(shape, B, 0))
lp = device.slot_relaxation(inst)
X = inst.F * inst.N * inst.N
by = B * (64 * X + 112 * inst.F * inst.N + 8 * inst.N ** 2)
for rep in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    x, y, res = device.pdhg_mf_solve(lp, max_iters=iters, check_every=iters, eps_rel=1e-12, eps_abs=1e-14)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{shape} B={B} iters={int(res['iters'][0])} {ms:.2f} ms  {ms/iters*1e3:.1f} us/iter  {by*iters/ms/1e6:.1f} GB/s", flush=True)
