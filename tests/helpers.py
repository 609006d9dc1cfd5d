"""Shared test helpers: payload -> Data -> oracle arrays / device batch."""
import numpy as np

from neptune_mip_b200 import synth
from neptune_mip_b200.core.utils import data_to_solver_input
from oracle import model as omodel

KIND_NAMES = ("min_delay", "min_util", "min_delay_util")


def data_of(payload):
    return data_to_solver_input(payload, payload.get("workload_coeff", 1), with_db=False)


def arrays_of(payload):
    return omodel.arrays_from_data(data_of(payload))


def small_payloads():
    """(name, payload, alpha) triples covering the shapes the reference tests (and ragged ones)."""
    out = [("C1", synth.test_py_payload(), 1.0)]
    for k in (0, 1, 2, 3, 4, 5, 6):
        out.append((f"sim{k}", synth.simulated_case(k), 0.0))
    for s in range(3):
        out.append((f"r8x4s{s}", synth.random_payload(8, 4, s, node_cores=30), 0.5))
        out.append((f"r12x5s{s}", synth.random_payload(12, 5, s, node_cores=25), 0.3))
    out.append(("r7x3float", float_payload(7, 3, 5), 0.4))
    return out


def float_payload(N, F, seed):
    """Non-integer delays / workloads / memories: exercises rounding order."""
    rng = np.random.default_rng(seed)
    p = synth.random_payload(N, F, seed, node_cores=20)
    D = rng.uniform(0.5, 40.0, (N, N)); D = (D + D.T) / 2; np.fill_diagonal(D, 0.0)
    p["node_delay_matrix"] = D.tolist()
    p["workload_on_source_matrix"] = rng.uniform(0.0, 9.0, (F, N)).tolist()
    p["function_memories"] = rng.uniform(0.1, 0.49, F).round(2).tolist()
    p["node_memories"] = [1.0] * N
    return p


def cuda_batch(payloads):
    from neptune_mip_b200.device import InstanceBatch
    return InstanceBatch.from_datas([data_of(p) for p in payloads])
