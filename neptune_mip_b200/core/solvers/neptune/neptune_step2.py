"""Neptune step-2 "minimise disruption" classes (reference `neptune_step2.py:5-93`,
`constraints_step2.py:5-89`, `objectives.py:55-63`).

For a fixed placement c the auxiliaries of the reference's step-2 MIP have a closed form:
moved_from = max(0, c-old), moved_to = max(0, old-c), and the mode is feasible iff
sum(c) <= sum(old) ("delete") / sum(c) >= sum(old) ("create"), giving
  obj2 = W*|c xor old| - (W+1)*(sum(old)-sum(c))   (delete)
  obj2 = W*|c xor old| - (W-1)*(sum(c)-sum(old))   (create),      W = F*N.
`solve()` searches placements for the smallest obj2 with the kernel `neptune_disruption_search`
(csrc/search.cu: add / drop / swap / replace / exchange moves, exact routing with the CPU rows), seeded
with the step-1 placement and the old allocation, under the step-1 rows and the softened objective row
(`constrain_network_delay` :57-69, `constrain_node_utilization` :71-73, `constrain_score` :76-89).
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
        return self._kind_score() <= self._bound() + 1e-9 * max(1.0, abs(self._bound()))

    def _bound(self) -> float:
        """Right-hand side of the softened step-1 objective row."""
        return float(self.data.max_score) * float(self.soften_step1_sol)

    def _unsolved(self):
        prev_c = np.asarray(self.data.prev_c)
        self._x = np.asarray(self.data.prev_x) * 0.0
        self._c = prev_c * 0.0
        self._n = np.zeros(len(self.data.nodes))
        self._obj = 0.0
        return False

    def solve(self):
        from .... import device
        prev_c = (np.asarray(self.data.prev_c) > 0.001).astype(np.uint8)
        old = (np.asarray(self.data.old_allocations_matrix) > 0).astype(np.uint8)
        seeds = torch.from_numpy(np.stack([prev_c, old])[None]).cuda().contiguous()          # [1, 2, F, N]
        bound = torch.tensor([self._bound()], dtype=torch.float64, device="cuda")
        best_c, best_obj, _ = device.disruption_search(self.inst, self.kind, self.mode, bound, seeds, self._alpha(),
                                                       self.chains, self.sweeps, self.rng_seed)
        # the search's winner first; the step-1 placement as a fallback (it satisfies the softened row by
        # construction and is feasible, so the mode is "solved" whenever its pod count allows it)
        candidates = ([best_c] if bool(torch.isfinite(best_obj[0])) else []) + [seeds[:, 0].contiguous()]
        for cand in candidates:
            ok = self._finish(cand, capacitated=True)
            val = disruption(self._c, self.data.old_allocations_matrix, self.mode)
            # re-check the softened objective row on the exact routing, as written in the reference
            if ok and val is not None and self._objective_preserved():
                self._obj = val
                return True
        return self._unsolved()

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
        x = np.asarray(self._x, dtype=np.float64)
        n = (np.asarray(self._c) > 0.001).any(axis=0)
        N, F = d.shape[0], w.shape[0]
        maxd = np.maximum(np.asarray(self.data.max_delay_matrix, dtype=np.float64)[None, :],
                          d.max(axis=0)[:, None])                       # [i, f]
        lhs = n.sum() * (self.alpha / N) + float(
            np.sum(x * ((1 - self.alpha) * w.T[:, :, None] * d[:, None, :] / maxd[:, :, None])))
        return lhs <= self.data.max_score * self.soften_step1_sol + 1e-9
