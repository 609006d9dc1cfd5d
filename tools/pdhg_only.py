"""PDHG kernels alone on the bench workload (C2 batch, strengthened LP): timing + ncu target."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from neptune_mip_b200 import device
from neptune_mip_b200._lib import FLAG_STRENGTHEN
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 256
flags = FLAG_STRENGTHEN if (len(sys.argv) <= 3 or sys.argv[3] == "s") else 0
shape = sys.argv[4] if len(sys.argv) > 4 else "C2"
if shape == "C2":
    host = bench.make_hosts(B, 0)
    inst = device.InstanceBatch.from_host(host)
else:
    from neptune_mip_b200 import synth
    from neptune_mip_b200.core.utils import data_to_solver_input
    inst = device.InstanceBatch.from_datas([data_to_solver_input(synth.config_payload(shape, s), 1, with_db=False) for s in range(B)])
mdl = device.assemble(inst, "min_delay", flags=flags)
for rep in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    x, y, res = device.pdhg_solve(mdl, max_iters=iters, check_every=iters, eps_rel=0.0, eps_abs=0.0)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    by = B * (16 * mdl.nnz + 88 * mdl.cols + 72 * mdl.rows) + 8 * mdl.nnz + 8 * (mdl.rows + mdl.cols + 2)
    print(f"{shape} B={B} rows={mdl.rows} cols={mdl.cols} nnz={mdl.nnz} iters={int(res['iters'][0])} {ms:.2f} ms  {ms/iters*1e3:.1f} us/iter  {by*iters/ms/1e6:.1f} GB/s", flush=True)
