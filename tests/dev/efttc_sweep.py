"""BASELINE config 5: sweep of 4096 independent 20x5 instances through the EFTTC solvers.
GPU: one launch per objective (device-resident and host-buffer entry points); CPU: the oracle's
restatement of the reference's greedy on a bounded sample.  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from neptune_mip_b200 import device, synth
from neptune_mip_b200.core.utils import data_to_solver_input
from oracle import efttc as oefttc, model as omodel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
datas = [data_to_solver_input(synth.config_payload("C5", s), 1, with_db=False) for s in range(B)]
host = device.InstanceBatch.host_arrays(datas)
inst = device.InstanceBatch.from_host(host)
out = {"workload": f"C5: {B} x (20 nodes x 5 functions), EFTTC", "kinds": {}}
for kind in ("min_delay", "min_util", "min_delay_util"):
    for _ in range(3):
        device.efttc(inst, kind)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c, n, info = device.efttc(inst, kind)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 5
    t0 = time.time()
    for _ in range(3):
        ch, nh, ih, fh, sh = device.efttc_host(host, kind)
    ms_host = (time.time() - t0) / 3 * 1e3
    assert np.array_equal(ch, c.cpu().numpy())
    # CPU: oracle port of the reference greedy, bounded sample
    ns = 32
    t0 = time.time()
    for b in range(ns):
        ref = oefttc.solve(omodel.arrays_from_data(datas[b]), kind, 0.5, strict=False)
        assert np.array_equal(ref.c.astype(np.uint8), ch[b])
    cpu_s = (time.time() - t0) / ns
    out["kinds"][kind] = {"gpu_instances_per_s": B / (ms / 1e3), "gpu_ms_per_sweep": ms,
                          "gpu_host_buffer_instances_per_s": B / (ms_host / 1e3), "host_buffer_ms_per_sweep": ms_host,
                          "cpu_oracle_port_instances_per_s_1core": 1.0 / cpu_s, "cpu_sample": ns,
                          "placements_equal_on_sample": True,
                          "feasible_all_checks": int((fh == 63).sum()), "mean_ttc_iterations": float(ih[:, 0].mean())}
print(json.dumps(out))
