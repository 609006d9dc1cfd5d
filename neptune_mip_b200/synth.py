"""Synthetic NEPTUNE payloads (frozen generators) -- the workloads of BASELINE.json's configs.

* `test_py_payload()`   -- C1: the literal 3x2 smoke payload the reference posts (`test.py:4-60`).
* `simulated_case(k)`   -- the ten shapes of `testing/simulated/simulated_test.py:25-380`
                           (zero workload, unit delays, memory 100, cores 100).
* `random_payload(...)` -- SURVEY.md section 8(d) generator: random symmetric integer delays,
                           integer workloads, cores/requests matrices; C2 = (50, 10, seed 0,
                           node_cores=200), C5 = 4096 x (20, 5, seeds 0..4095, node_cores=100),
                           C3 = (500, 50), C4 = (2000, 200) with `node_cores=None` (auto ~40 % load).
All generators return plain JSON-able dicts in the REST payload schema (`payload.json`).
"""
from __future__ import annotations

import math

import numpy as np


def _base(node_names, function_names, node_memories, node_cores, function_memories,
          max_delay=100, allocations=None, solver=None):
    p = {
        "with_db": False,
        "workload_coeff": 1,
        "community": "community-test",
        "namespace": "namespace-test",
        "node_names": list(node_names),
        "node_memories": list(node_memories),
        "node_cores": list(node_cores),
        "gpu_node_names": [],
        "gpu_node_memories": [],
        "function_names": list(function_names),
        "function_memories": list(function_memories),
        "function_max_delays": [max_delay for _ in function_names],
        "gpu_function_names": [],
        "gpu_function_memories": [],
        "actual_cpu_allocations": allocations if allocations is not None else {},
        "actual_gpu_allocations": {},
    }
    if solver is not None:
        p["solver"] = solver
    return p


def test_py_payload(solver_type="NeptuneMinDelayAndUtilization", args=None):
    """C1 (reference `test.py:4-60`): 3 nodes x 2 functions, everything pre-allocated."""
    nodes = ["node_a", "node_b", "node_c"]
    funcs = ["ns/fn_1", "ns/fn_2"]
    if args is None:
        args = {"alpha": 1, "verbose": False, "soften_step1_sol": 1.3}
    p = _base(nodes, funcs, [100, 100, 200], [100, 50, 50], [5, 5], max_delay=1000,
              allocations={f: {n: True for n in nodes} for f in funcs},
              solver={"type": solver_type, "args": args})
    p["node_delay_matrix"] = [[0, 3, 2], [3, 0, 4], [2, 4, 0]]
    p["workload_on_source_matrix"] = [[100, 0, 0], [1, 0, 0]]
    p["cores_matrix"] = [[1, 1, 1] for _ in funcs]
    p["workload_on_destination_matrix"] = [[1, 1, 1] for _ in funcs]
    return p


_SIM = [
    # (nodes, functions, function memory, allocation pattern)
    (1, 1, 10, "none_keyed"), (1, 1, 10, "all"), (1, 2, 10, "none"), (1, 2, 10, "first"), (1, 2, 10, "all"),
    (20, 5, 30, "none"), (20, 5, 10, "node1"), (50, 15, 30, "none"), (50, 5, 30, "none"), (25, 15, 30, "none"),
]


def simulated_case(k, solver_type="NeptuneMinDelay", alpha=0.0):
    """Shape k of the reference's simulated suite (`simulated_test.py:25-380`)."""
    n, f, fmem, pattern = _SIM[k]
    if n == 1:
        nodes = ["node_a"]
        funcs = [f"ns/fn_{i + 1}" for i in range(f)]
    else:
        nodes = [f"node_{i}" for i in range(n)]
        funcs = [f"ns/fn_{i}" for i in range(f)]
    if pattern == "none":
        alloc = {}
    elif pattern == "none_keyed":
        alloc = {fn: {} for fn in funcs}
    elif pattern == "all":
        alloc = {fn: {nodes[0]: True} for fn in funcs}
    elif pattern == "first":
        alloc = {funcs[0]: {nodes[0]: True}}
    elif pattern == "node1":
        alloc = {fn: {nodes[1]: True} for fn in funcs}
    else:  # pragma: no cover
        raise ValueError(pattern)
    p = _base(nodes, funcs, [100] * n, [100] * n, [fmem] * f, allocations=alloc,
              solver={"type": solver_type, "args": {"alpha": alpha, "verbose": False}})
    p["case"] = k
    return p


def random_payload(n_nodes, n_functions, seed, node_cores=200, solver_type="NeptuneMinDelay",
                   args=None, function_memory=30, node_memory=100):
    """SURVEY.md section 8(d) generator.  `node_cores=None` -> ceil(2.5 * mean CPU demand per node)."""
    rng = np.random.default_rng(seed)
    N, F = n_nodes, n_functions
    D = rng.integers(1, 50, (N, N))
    D = (D + D.T) // 2
    np.fill_diagonal(D, 0)
    W = rng.integers(0, 20, (F, N))
    cores = rng.integers(1, 5, (F, N))
    wdst = rng.integers(1, 10, (F, N))
    if node_cores is None:
        r = cores / wdst
        node_cores = int(math.ceil(2.5 * float((W.sum(axis=1) * r.mean(axis=1)).sum()) / N))
    nodes = [f"node_{i}" for i in range(N)]
    funcs = [f"ns/fn_{i}" for i in range(F)]
    p = _base(nodes, funcs, [node_memory] * N, [node_cores] * N, [function_memory] * F,
              solver={"type": solver_type, "args": args if args is not None else {"verbose": False}})
    p["node_delay_matrix"] = D.tolist()
    p["workload_on_source_matrix"] = W.tolist()
    p["cores_matrix"] = cores.tolist()
    p["workload_on_destination_matrix"] = wdst.tolist()
    p["seed"] = int(seed)
    return p


def config_payload(name, seed=0, **kw):
    """Named BASELINE.json configs: 'C1'..'C5' (C5 returns ONE instance of the sweep, by seed)."""
    name = name.upper()
    if name == "C1":
        return test_py_payload(**kw)
    if name == "C2":
        return random_payload(50, 10, seed, node_cores=200, **kw)
    if name == "C2-TIGHT":
        return random_payload(50, 10, seed, node_cores=100, **kw)
    if name == "C3":
        return random_payload(500, 50, seed, node_cores=None, **kw)
    if name == "C4":
        return random_payload(2000, 200, seed, node_cores=None, **kw)
    if name == "C5":
        return random_payload(20, 5, seed, node_cores=100, **kw)
    raise ValueError(name)


def payload_json_sample():
    """The shape of the reference's `payload.json` (5 nodes incl. one GPU node, 4 functions incl. one GPU
    function, no matrices at all -> every default of `input_to_data.py:151-183` kicks in), with the `//`
    comments of the shipped file removed and `with_db` set to false.  SURVEY.md section 8(c): the reference
    answers score {'step1': 0.2, 'step2': 39.0} with all four functions on one node."""
    nodes = ["node_a", "node_b", "node_c", "node_d", "gpu_node_e"]
    funcs = ["ns/fn_1", "ns/fn_2", "ns/fn_3", "ns/gpu_fn_4"]
    p = _base(nodes, funcs, [40, 40, 30, 30, 30], [10] * 5, [10] * 4, max_delay=100,
              allocations={"ns/fn_1": {"node_a": True}, "ns/fn_2": {"node_b": True, "node_c": True},
                           "ns/fn_3": {"node_b": True}, "ns/gpu_fn_4": {"node_b": True}},
              solver={"type": "NeptuneMinDelayAndUtilization", "args": {"alpha": 1.0, "verbose": False}})
    p["gpu_node_names"] = ["gpu_node_e"]
    p["gpu_node_memories"] = [100]
    p["gpu_function_names"] = ["ns/gpu_fn_4"]
    p["gpu_function_memories"] = [50]
    p.pop("workload_coeff")
    return p
