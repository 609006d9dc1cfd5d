"""REST payload -> `Data` (the reference's `core/utils/input_to_data.py`, re-implemented).

Behaviour kept from the reference (file:line are `/root/reference/core/utils/input_to_data.py`):
  * required payload keys `:9-25`, consistency asserts `:46-86` (bare `assert`, same messages);
  * defaults `:151-183`: delay 1 off the diagonal / 0 on it, zero workloads, zero cores;
  * `max_delay_matrix` is 1000 for every function, `function_max_delays` is ignored `:136`;
  * `workload_matrix = workload_on_source * workload_coeff` `:102`;
  * `core_per_req = cores / workload_on_destination` with 0/0 -> 0 and x/0 -> DBL_MAX `:272`;
  * old allocations become a 0/1 matrix and **all ones when nothing is allocated** `:274-276`;
  * `node_costs = 5`, `node_budget = 300` `:185-187`.
The Postgres pull (`:206-262`, `with_db=True`) needs the in-cluster metrics database and is out of
scope: it raises instead of silently continuing.
"""
from __future__ import annotations

import numpy as np

from .data import Data

keys = [
    "community", "namespace",
    "function_names", "function_memories",
    "gpu_function_names", "gpu_function_memories",
    "node_names", "node_memories", "node_cores",
    "gpu_node_names", "gpu_node_memories",
    "function_max_delays",
    "actual_cpu_allocations", "actual_gpu_allocations",
]

NODE_COST = 5
NODE_BUDGET = 300
MAX_DELAY = 1000


def check_input(schedule_input, verbose: bool = False):
    """Same assertions as the reference's `check_input` (`:46-86`)."""
    for key in keys:
        assert key in schedule_input.keys(), f"Key `{key}` not in schedule input"
    functions = schedule_input.get("function_names", [])
    gpu_functions = schedule_input.get("gpu_function_names", [])
    assert set(gpu_functions).issubset(set(functions))
    assert len(functions) == len(schedule_input.get("function_memories", []))
    assert len(gpu_functions) == len(schedule_input.get("gpu_function_memories", []))
    nodes = schedule_input.get("node_names", [])
    gpu_nodes = schedule_input.get("gpu_node_names", [])
    assert set(gpu_nodes).issubset(set(nodes))
    assert len(nodes) == len(schedule_input.get("node_memories", []))
    assert len(gpu_nodes) == len(schedule_input.get("gpu_node_memories", []))
    if verbose:
        print(f"{len(nodes)} nodes, {len(functions)} functions: input consistent")


def _matrix_or_default(value, default):
    # the reference tests truthiness of the raw JSON value (`if node_delay_matrix:` `:153,159,166,174`)
    if value:
        return np.array(value)
    return default


def data_to_solver_input(input, workload_coeff, with_db=True) -> Data:
    """Signature and argument meaning of the reference's function (`:88-111`)."""
    functions = input.get("function_names", [])
    nodes = input.get("node_names", [])
    gpu_functions = input.get("gpu_function_names", [])
    gpu_nodes = input.get("gpu_node_names", [])
    assert set(gpu_functions).issubset(set(functions))
    assert set(gpu_nodes).issubset(set(nodes))
    F, N = len(functions), len(nodes)

    if with_db:
        raise RuntimeError(
            "with_db=True pulls metrics from the in-cluster Postgres "
            "(reference input_to_data.py:206-262); that store is out of scope here -- "
            "send the matrices in the payload and set with_db=false")

    delay = _matrix_or_default(input.get("node_delay_matrix"),
                               (1 - np.eye(N, dtype=int)) if N else np.zeros((0, 0), dtype=int))
    w_src = _matrix_or_default(input.get("workload_on_source_matrix"), np.zeros((F, N), dtype=int))
    w_dst = _matrix_or_default(input.get("workload_on_destination_matrix"), np.zeros((F, N), dtype=int))
    cores = _matrix_or_default(input.get("cores_matrix"), np.zeros((F, N), dtype=int))

    node_map = {name: i for i, name in enumerate(nodes)}
    func_map = {name.split("/")[1]: i for i, name in enumerate(functions)}  # `:199`

    old = np.zeros((F, N), dtype=int)
    for function_key, per_node in (input.get("actual_cpu_allocations") or {}).items():
        if not per_node:
            continue
        f = func_map[function_key.split("/")[1]]
        for node, ok in per_node.items():
            old[f, node_map[node]] = ok
    old = old.astype(bool).astype(int)
    if old.sum() == 0:
        old = old + 1

    with np.errstate(divide="ignore", invalid="ignore"):
        core_per_req = np.nan_to_num(np.asarray(cores) / np.asarray(w_dst), nan=0)

    data = Data(nodes, functions)
    data.node_memory_matrix = np.array(input.get("node_memories"))
    data.function_memory_matrix = np.array(input.get("function_memories"))
    data.node_delay_matrix = np.array(delay)
    data.workload_matrix = np.array(w_src) * workload_coeff
    data.max_delay_matrix = np.array([MAX_DELAY for _ in range(F)])
    data.response_time_matrix = np.zeros((F, N), dtype=int)
    data.node_cores_matrix = np.array(input.get("node_cores"))
    data.cores_matrix = np.array(cores)
    data.old_allocations_matrix = old
    data.core_per_req_matrix = np.array(core_per_req)
    data.node_costs = np.array([NODE_COST for _ in nodes])
    data.node_budget = NODE_BUDGET
    return data
