"""TEST INFRASTRUCTURE ONLY -- the reference's `Solver.solve()` (`core/solvers/solver.py:35-40`) on CPU.

The reference hands its model to OR-Tools' bundled SCIP (`ortools==9.6.2534`, `requirements.txt:8`,
not vendored, not installable offline).  Here the model restated by `oracle.model` goes to HiGHS
through `scipy.optimize.milp` with `mip_rel_gap=0` -- a different branch-and-bound engine cannot
change the optimal objective value, which is what parity is anchored on (placements at degenerate
optima legitimately differ, SURVEY.md section 7 "hard parts" 4).
"""
from __future__ import annotations

import time

import numpy as np
from scipy.optimize import Bounds, LinearConstraint, linprog, milp

from . import model


def solve_model(mdl, time_limit=None, relax=False):
    """Solve a model dict from `oracle.model.build_*`.  Returns dict(status, objective, sol, ...)."""
    opts = {"mip_rel_gap": 0.0, "disp": False}
    if time_limit is not None:
        opts["time_limit"] = float(time_limit)
    integ = np.zeros_like(mdl["integ"]) if relax else mdl["integ"]
    t0 = time.time()
    res = milp(mdl["obj"], constraints=LinearConstraint(mdl["A"], mdl["lo"], mdl["hi"]),
               integrality=integ, bounds=Bounds(mdl["lb"], mdl["ub"]), options=opts)
    if res.status == 2:
        # HiGHS' presolve wrongly reports some big-M (1e6) models infeasible (e.g. the 1x1 case of the
        # simulated suite, feasible point x=c=1); the verdict only counts with presolve off.
        res = milp(mdl["obj"], constraints=LinearConstraint(mdl["A"], mdl["lo"], mdl["hi"]),
                   integrality=integ, bounds=Bounds(mdl["lb"], mdl["ub"]), options=dict(opts, presolve=False))
    dt = time.time() - t0
    out = dict(status=int(res.status), optimal=(res.status == 0 and res.x is not None),
               seconds=dt, sol=res.x, objective=(float(res.fun) if res.x is not None else None),
               dual_bound=getattr(res, "mip_dual_bound", None), gap=getattr(res, "mip_gap", None),
               node_count=getattr(res, "mip_node_count", None))
    return out


def split_solution(mdl, sol):
    """x[N,F,N] (index [i][f][j]), c[F,N], n[N] from a column vector (`neptune/utils/output.py:5-21`)."""
    N, F = mdl["N"], mdl["F"]
    X, C = F * N * N, F * N
    x = sol[:X].reshape(F, N, N).transpose(1, 0, 2).copy()
    c = sol[X:X + C].reshape(F, N).copy()
    n = sol[X + C:X + C + N].copy() if mdl["with_n"] else (c > 0.5).any(axis=0).astype(np.float64)
    return x, c, n


def solve_step1(a, kind, alpha=0.5, time_limit=None):
    mdl = model.build_step1(a, kind, alpha)
    out = solve_model(mdl, time_limit=time_limit)
    if out["sol"] is not None:
        out["x"], out["c"], out["n"] = split_solution(mdl, out["sol"])
    out["model"] = mdl
    return out


def lp_relaxation(mdl):
    """LP relaxation optimum by HiGHS (reference value for the PDHG kernels)."""
    res = linprog(mdl["obj"], A_ub=None, bounds=list(zip(mdl["lb"], mdl["ub"])), method="highs",
                  A_eq=None) if mdl["A"].shape[0] == 0 else None
    if res is None:
        out = solve_model(mdl, relax=True)
        return out
    return dict(objective=float(res.fun), sol=res.x, optimal=res.status == 0)
