"""Solver base class -- the drop-in boundary (reference `core/solvers/solver.py:5-46`).

Same template as the reference: `load_data` -> `init_vars` + `init_constraints`, `solve` ->
`init_objective` + the engine, `results`, `score`.  The engine is not OR-Tools/SCIP but the
CUDA kernels behind `libneptune_b200.so`; a step object keeps its device tensors for the lifetime of
the request and nothing outlives it.
"""
from __future__ import annotations

import datetime

from ..utils.data import Data


class Solver:
    def __init__(self, verbose: bool = True, **kwargs):
        self.verbose = verbose
        self.data = None
        self.args = kwargs          # unknown keyword arguments are swallowed, like the reference (:13)

    def load_data(self, data: Data):
        self.data = data
        self.log("Initializing variables...")
        self.init_vars()
        self.log("Initializing constraints...")
        self.init_constraints()

    def init_vars(self):
        raise NotImplementedError("Solvers must implement init_vars()")

    def init_constraints(self):
        raise NotImplementedError("Solvers must implement init_constraints()")

    def init_objective(self):
        raise NotImplementedError("Solvers must implement init_objective()")

    def log(self, msg: str):
        if self.verbose:
            print(f"{datetime.datetime.now()}: {msg}")

    def solve(self):
        raise NotImplementedError("Solvers must implement solve()")

    def results(self):
        raise NotImplementedError("Solvers must implement results()")

    def score(self) -> float:
        raise NotImplementedError("Solvers must implement score()")
