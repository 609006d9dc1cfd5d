"""Neptune step-1 classes (reference `core/solvers/neptune/neptune_step1.py:5-77`) on the GPU path.

`load_data` assembles the reference's MIP matrix on device (kernel group (a)); `solve` runs
LP relaxation by PDHG (b) for the bound and the rounding guide, the EFTTC greedy (d) for seeds, the
batched local search (c2), and the exact checkers/scorers (c1) on the winner.
"""
from __future__ import annotations

import numpy as np
import torch

from .... import device
from ...._lib import FLAG_STRENGTHEN
from ..gpu_step import GpuStepMixin
from ..solver import Solver


class NeptuneStepBase(GpuStepMixin, Solver):
    def __init__(self, chains: int = 296, sweeps: int = 300, lp_iters: int = 0, rng_seed: int = 1,
                 keep_model: bool = True, **kwargs):
        super().__init__(**kwargs)
        self.chains, self.sweeps, self.lp_iters, self.rng_seed = chains, sweeps, lp_iters, rng_seed
        self.keep_model = keep_model
        self.model = None
        self.lp_bound = None
        self.lp_result = None
        self._x = self._c = self._n = None

    def init_vars(self):
        self._upload()

    def init_constraints(self):
        # the reference builds rows here; so do we (on device), flags=0 is the as-written matrix
        if self.keep_model:
            self.model = device.assemble(self.inst, self.kind, self._alpha())

    def init_objective(self):
        pass                        # objective coefficients are part of `assemble`

    def results(self):
        self.data.prev_x = self._x
        self.data.prev_c = self._c
        return self._x, self._c


class NeptuneStep1CPUBase(NeptuneStepBase):
    def solve(self):
        self.init_objective()
        inst, alpha = self.inst, self._alpha()
        guide = None
        if self.lp_iters > 0:
            if self.kind == "min_delay":
                # matrix-free relaxation of the strengthened min-delay model (nothing is assembled)
                xs, ys, res = device.pdhg_mf_solve(inst, max_iters=self.lp_iters, eps_rel=1e-5)
            else:
                lp = device.assemble(inst, self.kind, alpha, flags=FLAG_STRENGTHEN)
                xs, ys, res = device.pdhg_solve(lp, max_iters=self.lp_iters, eps_rel=1e-5)
                del lp
            X = inst.F * inst.N * inst.N
            guide = xs[:, X:X + inst.F * inst.N].contiguous()
            self.lp_result = res[0]
            self.lp_bound = float(res[0]["dual_obj"])
            del xs, ys
        seeds = [device.efttc(inst, k, alpha)[0] for k in ("min_delay", "min_util", "min_delay_util")]
        seeds = torch.stack(seeds, dim=1).contiguous()                       # [1, 3, F, N]
        best_c, best_obj, _ = device.local_search(inst, self.kind, seeds, alpha, self.chains, self.sweeps,
                                                  self.rng_seed, guide)
        if not torch.isfinite(best_obj[0]):
            best_c = seeds[:, {"min_delay": 0, "min_util": 1, "min_delay_util": 2}[self.kind]].contiguous()
        ok = self._finish(best_c, capacitated=True)
        if not ok:      # rare: fall back to the first EFTTC seed that passes every check
            for k in ({"min_delay": 0, "min_util": 1, "min_delay_util": 2}[self.kind], 1, 2, 0):
                if self._finish(seeds[:, k].contiguous(), capacitated=True):
                    ok = True
                    break
        self.log(f"step 1: score {self._kind_score()} flags {self.flags:06b} lp bound {self.lp_bound}")
        return ok

    def score(self):
        return self._kind_score()


class NeptuneStep1CPUMinUtilization(NeptuneStep1CPUBase):
    kind = "min_util"

    def results(self):
        x, c = super().results()
        self.data.prev_n = self._n
        return x, c


class NeptuneStep1CPUMinDelay(NeptuneStep1CPUBase):
    kind = "min_delay"


class NeptuneStep1CPUMinDelayAndUtilization(NeptuneStep1CPUMinUtilization):
    kind = "min_delay_util"

    def __init__(self, alpha=0.5, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    def load_data(self, data):
        data.alpha = self.alpha
        super().load_data(data)
