"""Response shaping (reference `neptune/utils/output.py:23-40`, duplicated in `efttc/utils/output.py`):
routing fractions > 0.001 rounded to 3 decimals, placements > 0.001 as `True`, absent keys for zeros."""
from __future__ import annotations

import numpy as np


def convert_x_matrix(matrix, nodes, functions):
    matrix = np.asarray(matrix)
    assert matrix.shape == (len(nodes), len(functions), len(nodes)), \
        f"X matrix shape malformed. matrix shape is {matrix.shape} but it should be {(len(nodes), len(functions), len(nodes))}"
    routings = {}
    for i, f, j in zip(*np.nonzero(matrix > 0.001)):
        routings.setdefault(nodes[i], {}).setdefault(functions[f], {})[nodes[j]] = float(np.round(matrix[i, f, j], 3))
    return routings


def convert_c_matrix(matrix, functions, nodes):
    matrix = np.asarray(matrix)
    assert matrix.shape == (len(functions), len(nodes)), \
        f"X matrix shape malformed. matrix shape is {matrix.shape} but it should be {(len(functions), len(nodes))}"
    allocations = {}
    for f, j in zip(*np.nonzero(matrix > 0.001)):
        allocations.setdefault(functions[f], {})[nodes[j]] = True
    return allocations
