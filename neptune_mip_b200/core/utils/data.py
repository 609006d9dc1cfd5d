"""`Data` -- the bag of numpy arrays every solver consumes.

Mirrors the field set of the reference's `core/utils/data.py:5-26` (same attribute names, so code
written against the reference's `Data` keeps working).  Solvers mutate it exactly like the reference
does: `alpha` (`neptune_step1.py:73`), `prev_x/prev_c/prev_n` (`neptune_step1.py:25-26,58`),
`max_score` (`neptune.py:22`).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


class Data:
    def __init__(self, nodes: Optional[List[str]] = None, functions: Optional[List[str]] = None):
        self.nodes = list(nodes) if nodes else []
        self.functions = list(functions) if functions else []
        empty = lambda: np.array([])  # noqa: E731
        self.node_memory_matrix = empty()
        self.function_memory_matrix = empty()
        self.node_delay_matrix = empty()
        self.workload_matrix = empty()
        self.max_delay_matrix = empty()
        self.response_time_matrix = empty()
        self.node_cores_matrix = empty()
        self.cores_matrix = empty()
        self.old_allocations_matrix = empty()
        self.core_per_req_matrix = empty()
        self.gpu_function_memory_matrix = empty()
        self.gpu_node_memory_matrix = empty()
        self.prev_x = empty()
        self.node_costs = empty()
        self.node_budget = 0

    # -- convenience used by the B200 host layer (not part of the reference surface) ------------
    @property
    def num_nodes(self) -> int:
        return len(self.nodes)

    @property
    def num_functions(self) -> int:
        return len(self.functions)
