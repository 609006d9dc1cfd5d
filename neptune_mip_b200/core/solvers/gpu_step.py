"""Shared device plumbing of the step classes: one instance on the GPU, exact evaluation of a placement."""
from __future__ import annotations

import numpy as np
import torch

from ... import device
from ..._lib import OK_ALL


class GpuStepMixin:
    """Holds the `InstanceBatch` of the request and the last placement (device + host copies)."""

    kind = "min_delay"

    def _upload(self):
        self.inst = device.InstanceBatch.from_datas([self.data])

    def _alpha(self):
        return float(getattr(self, "alpha", 0.5))

    def _finish(self, c_u8: torch.Tensor, capacitated: bool = False):
        """Exact routing + the reference's checkers/scorers for placement c (uint8 [1,F,N]).
        `capacitated`: route with the CPU rows honoured (MIP semantics) instead of the plain
        nearest-open-pod rule (EFTTC semantics)."""
        inst = self.inst
        if capacitated:
            c_u8, x, n, _, _ = device.route_capacitated(inst, c_u8)
        else:
            x, n = device.route_placements(inst, c_u8)
        flags, scores = device.check_solution(inst, x, device.u8_to_f64(c_u8), n, self._alpha())
        return self._take(c_u8, x, n, flags, scores)

    def _take(self, c_u8, x, n, flags, scores):
        """Host copies of an evaluated placement (first instance of the batch)."""
        self._x = x[0].cpu().numpy()
        self._c = c_u8[0].cpu().numpy().astype(np.float64)
        self._n = n[0].cpu().numpy()
        self.flags = int(flags.cpu()[0])
        self.scores = scores[0].cpu().numpy()
        self.feasible = self.flags == OK_ALL
        return self.feasible

    def _kind_score(self):
        k = {"min_delay": 0, "min_util": 1, "min_delay_util": 2}[self.kind]
        v = float(self.scores[k])
        return int(v) if self.kind == "min_util" else v
