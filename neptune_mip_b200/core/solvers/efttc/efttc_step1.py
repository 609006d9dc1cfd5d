"""EFTTC step classes (reference `core/solvers/efttc/efttc_step1.py:6-441`) on the GPU kernel
`neptune_efttc` (csrc/efttc.cu): same greedy, same tie-breaks, same placement."""
from __future__ import annotations

import numpy as np

from .... import device
from ..gpu_step import GpuStepMixin
from ..solver import Solver


class EfttcStepBase(GpuStepMixin, Solver):
    objective = "min_delay_min_utilization"

    def __init__(self, strict_reference_errors: bool = False, **kwargs):
        # the reference raises KeyError (efttc_step1.py:118) whenever a function is met twice in an
        # accepted cycle; by default this build keeps going with `discard` semantics and records it
        self.strict_reference_errors = strict_reference_errors
        super().__init__(**kwargs)
        self.reference_would_raise = False
        self.ttc_iterations = 0

    def init_vars(self):
        pass

    def init_constraints(self):
        pass

    def solve(self):
        self._upload()
        c, n, info = device.efttc(self.inst, self.kind, self._alpha())
        info = info.cpu().numpy()[0]
        self.ttc_iterations = int(info[0])
        self.reference_would_raise = bool(info[2])
        if self.reference_would_raise and self.strict_reference_errors:
            raise KeyError("remaining_functions.remove(f): f already removed (reference efttc_step1.py:118)")
        self._finish(c)

    def results(self):
        # the reference stores prev_x / prev_c only in the utilisation classes (:380-387); the step-2 search
        # seeds itself from the step-1 placement, so every EFTTC step publishes it
        self.data.prev_x = self._x
        self.data.prev_c = self._c
        return self._x, self._c

    def score(self):
        return self._kind_score()


class EfttcStep1CPUBase(EfttcStepBase):
    pass


class EfttcStep1CPUMinUtilization(EfttcStep1CPUBase):
    kind = "min_util"
    objective = "min_utilization"

    def results(self):
        x, c = super().results()
        self.data.prev_n = self._n
        self.data.prev_x = x
        self.data.prev_c = c
        return x, c


class EfttcStep1CPUMinDelay(EfttcStep1CPUBase):
    kind = "min_delay"
    objective = "min_delay"


class EfttcStep1CPUMinDelayAndUtilization(EfttcStep1CPUMinUtilization):
    kind = "min_delay_util"
    objective = "min_delay_min_utilization"

    def __init__(self, alpha=0.5, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    def load_data(self, data):
        data.alpha = self.alpha
        super().load_data(data)
