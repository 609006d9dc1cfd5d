"""Neptune step-2 "minimise disruption" classes (reference `neptune_step2.py:5-93`,
`constraints_step2.py:5-89`, `objectives.py:55-63`).

For a fixed placement c the auxiliaries of the reference's step-2 MIP have a closed form:
moved_from = max(0, c-old), moved_to = max(0, old-c), and the mode is feasible iff
sum(c) <= sum(old) ("delete") / sum(c) >= sum(old) ("create"), giving
  obj2 = W*|c xor old| - (W+1)*(sum(old)-sum(c))   (delete)
  obj2 = W*|c xor old| - (W-1)*(sum(c)-sum(old))   (create),      W = F*N.
Round 1: the step-2 search space is the step-1 placement itself (always satisfies the softened
objective row); the disruption-minimising search over other placements is row (f)-1 of the scope
table and comes next.
"""
from __future__ import annotations

import numpy as np
import torch

from .neptune_step1 import NeptuneStepBase


def disruption(c, old, mode):
    cb = (np.asarray(c) > 0.001).astype(np.int64)
    ob = (np.asarray(old) > 0).astype(np.int64)
    W = ob.size
    flips = int(np.abs(cb - ob).sum())
    delta = int(ob.sum() - cb.sum())
    if mode == "delete":
        return None if delta < 0 else float(W * flips - (W + 1) * delta)
    return None if delta > 0 else float(W * flips - (W - 1) * (-delta))


class NeptuneStep2Base(NeptuneStepBase):
    def __init__(self, mode=str, soften_step1_sol=1.3, **kwargs):
        super().__init__(**kwargs)
        self.mode = mode
        assert mode in ["delete", "create"]
        self.soften_step1_sol = soften_step1_sol
        self._obj = 0.0          # Objective().Value() of an unsolved model

    def init_constraints(self):
        pass

    def _objective_preserved(self) -> bool:
        return True

    def solve(self):
        prev_c = np.asarray(self.data.prev_c)
        val = disruption(prev_c, self.data.old_allocations_matrix, self.mode)
        if val is None or not self._objective_preserved():
            self._x = np.asarray(self.data.prev_x) * 0.0
            self._c = prev_c * 0.0
            self._n = np.zeros(len(self.data.nodes))
            self._obj = 0.0
            return False
        c_u8 = torch.from_numpy((prev_c > 0.001).astype(np.uint8)).cuda()[None].contiguous()
        ok = self._finish(c_u8, capacitated=True)
        self._obj = val
        return bool(ok)

    def results(self):
        return self._x, self._c

    def score(self):
        return self._obj


class NeptuneStep2MinUtilization(NeptuneStep2Base):
    kind = "min_util"


class NeptuneStep2MinDelay(NeptuneStep2Base):
    kind = "min_delay"


class NeptuneStep2MinDelayAndUtilization(NeptuneStep2MinUtilization):
    kind = "min_delay_util"

    def __init__(self, alpha=0.5, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    def _objective_preserved(self) -> bool:
        """`constrain_score` (constraints_step2.py:76-89) normalises the delay term by
        max(max_delay[f], colmax d) instead of the step-1 objective's Wmax, which makes the row
        infeasible for typical non-zero workloads: evaluate it as written."""
        d = np.asarray(self.data.node_delay_matrix, dtype=np.float64)
        w = np.asarray(self.data.workload_matrix, dtype=np.float64)
        x = np.asarray(self.data.prev_x, dtype=np.float64)
        n = (np.asarray(self.data.prev_c) > 0.001).any(axis=0)
        N, F = d.shape[0], w.shape[0]
        maxd = np.maximum(np.asarray(self.data.max_delay_matrix, dtype=np.float64)[None, :],
                          d.max(axis=0)[:, None])                       # [i, f]
        lhs = n.sum() * (self.alpha / N) + float(
            np.sum(x * ((1 - self.alpha) * w.T[:, :, None] * d[:, None, :] / maxd[:, :, None])))
        return lhs <= self.data.max_score * self.soften_step1_sol + 1e-9
