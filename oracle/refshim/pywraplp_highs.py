"""TEST INFRASTRUCTURE ONLY -- a `pywraplp`-compatible shim over SciPy/HiGHS.

The reference builds its MIPs through OR-Tools' `pywraplp` and solves them with the
bundled SCIP (`/root/reference/core/solvers/solver.py:1,7`, `requirements.txt:8`).
OR-Tools is not installable offline, so this shim exposes exactly the `pywraplp`
surface the reference touches and hands the recorded model to
`scipy.optimize.milp` (HiGHS, `mip_rel_gap=0`).  With it the reference's *own* model
builders run unmodified; only the branch-and-bound engine differs, which cannot change
optimal objective values.

It is used in exactly two places:
  * `oracle/make_golden.py` (run in the build container, where /root/reference exists)
    to generate the fixtures under `tests/golden/`;
  * `tests/test_oracle_vs_reference.py`, skipped when /root/reference is absent.
Nothing under `neptune_mip_b200/` imports it.
"""
from __future__ import annotations

import numbers
import time

import numpy as np
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, milp

_INF = float("inf")


class LinExpr:
    """Affine expression sum(coef_k * var_k) + const."""

    __array_ufunc__ = None  # make numpy scalars defer to our reflected operators
    __slots__ = ("terms", "const")

    def __init__(self, terms=None, const=0.0):
        self.terms = terms if terms is not None else {}
        self.const = const

    # -- helpers ---------------------------------------------------------
    @staticmethod
    def lift(v):
        if isinstance(v, LinExpr):
            return v
        if isinstance(v, Var):
            return LinExpr({v.index: 1.0}, 0.0)
        if isinstance(v, (numbers.Number, np.generic)):
            return LinExpr({}, float(v))
        raise TypeError(f"cannot lift {type(v)} into a linear expression")

    def copy(self):
        return LinExpr(dict(self.terms), self.const)

    # -- arithmetic ------------------------------------------------------
    def __add__(self, other):
        o = LinExpr.lift(other)
        out = self.copy()
        for k, v in o.terms.items():
            out.terms[k] = out.terms.get(k, 0.0) + v
        out.const += o.const
        return out

    __radd__ = __add__

    def __neg__(self):
        return LinExpr({k: -v for k, v in self.terms.items()}, -self.const)

    def __sub__(self, other):
        return self + (-LinExpr.lift(other))

    def __rsub__(self, other):
        return LinExpr.lift(other) + (-self)

    def __mul__(self, k):
        if not isinstance(k, (numbers.Number, np.generic)):
            raise TypeError("only multiplication by constants is linear")
        k = float(k)
        return LinExpr({i: v * k for i, v in self.terms.items()}, self.const * k)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self * (1.0 / float(k))

    # -- comparisons -> constraints ---------------------------------------
    def __le__(self, other):
        d = self - LinExpr.lift(other)
        return Constraint(d.terms, -_INF, -d.const)

    def __ge__(self, other):
        d = self - LinExpr.lift(other)
        return Constraint(d.terms, -d.const, _INF)

    def __eq__(self, other):  # noqa: D105 - constraint builder, like pywraplp
        d = self - LinExpr.lift(other)
        return Constraint(d.terms, -d.const, -d.const)

    __hash__ = None


class Var:
    __array_ufunc__ = None
    __slots__ = ("solver", "index", "_name", "lb", "ub", "integer")

    def __init__(self, solver, index, name, lb, ub, integer):
        self.solver, self.index, self._name = solver, index, name
        self.lb, self.ub, self.integer = lb, ub, integer

    def name(self):
        return self._name

    def solution_value(self):
        sol = self.solver._solution
        return 0.0 if sol is None else float(sol[self.index])

    def _e(self):
        return LinExpr({self.index: 1.0}, 0.0)

    def __add__(self, o): return self._e() + o
    def __radd__(self, o): return self._e() + o
    def __sub__(self, o): return self._e() - o
    def __rsub__(self, o): return LinExpr.lift(o) - self._e()
    def __mul__(self, k): return self._e() * k
    def __rmul__(self, k): return self._e() * k
    def __neg__(self): return -self._e()
    def __le__(self, o): return self._e() <= o
    def __ge__(self, o): return self._e() >= o
    def __eq__(self, o): return self._e() == o
    def __hash__(self): return hash((id(self.solver), self.index))


class Constraint:
    __slots__ = ("terms", "lo", "hi")

    def __init__(self, terms, lo, hi):
        self.terms, self.lo, self.hi = terms, lo, hi


class Objective:
    def __init__(self, solver):
        self.solver = solver
        self.coefs = {}
        self.maximize = False

    def SetCoefficient(self, var, coef):
        self.coefs[var.index] = float(coef)

    def SetMinimization(self):
        self.maximize = False

    def SetMaximization(self):
        self.maximize = True

    def Value(self):
        sol = self.solver._solution
        if sol is None:
            return 0.0
        return float(sum(c * sol[i] for i, c in self.coefs.items()))


class Solver:
    OPTIMAL, FEASIBLE, INFEASIBLE, UNBOUNDED, ABNORMAL, NOT_SOLVED = 0, 1, 2, 3, 4, 6

    #: wall-clock cap handed to HiGHS (None = unlimited, like the reference)
    time_limit = None
    #: filled after every Solve(): list of dicts (sizes, status, seconds)
    solve_log = []

    def __init__(self):
        self.vars = []
        self.cons = []
        self._objective = Objective(self)
        self._solution = None
        self.last_status = None
        self.last_mip_gap = None
        self.last_dual_bound = None

    @staticmethod
    def CreateSolver(name):
        return Solver()

    def EnableOutput(self):
        pass

    @staticmethod
    def infinity():
        return _INF

    def _new(self, lb, ub, name, integer):
        v = Var(self, len(self.vars), name, float(lb), float(ub), integer)
        self.vars.append(v)
        return v

    def NumVar(self, lb, ub, name): return self._new(lb, ub, name, False)
    def IntVar(self, lb, ub, name): return self._new(lb, ub, name, True)
    def BoolVar(self, name): return self._new(0, 1, name, True)

    def Sum(self, items):
        out = LinExpr()
        for it in items:
            if isinstance(it, Var):
                out.terms[it.index] = out.terms.get(it.index, 0.0) + 1.0
            elif isinstance(it, LinExpr):
                for k, v in it.terms.items():
                    out.terms[k] = out.terms.get(k, 0.0) + v
                out.const += it.const
            else:
                out.const += float(it)
        return out

    def Add(self, constraint):
        if isinstance(constraint, bool):  # e.g. a tautology evaluated by python
            return None
        self.cons.append(constraint)
        return constraint

    def Objective(self):
        return self._objective

    def NumVariables(self): return len(self.vars)
    def NumConstraints(self): return len(self.cons)

    # -- matrix export (used by the oracle cross-checks) --------------------
    def export(self):
        """Return (A csr with sorted indices, lo, hi, obj, lb, ub, integrality)."""
        n = len(self.vars)
        indptr = [0]
        indices, data, lo, hi = [], [], [], []
        for con in self.cons:
            cols = sorted(con.terms)
            indices.extend(cols)
            data.extend(con.terms[c] for c in cols)
            indptr.append(len(indices))
            lo.append(con.lo)
            hi.append(con.hi)
        A = sp.csr_matrix((np.asarray(data, dtype=np.float64),
                           np.asarray(indices, dtype=np.int64),
                           np.asarray(indptr, dtype=np.int64)), shape=(len(self.cons), n))
        obj = np.zeros(n)
        for i, c in self._objective.coefs.items():
            obj[i] = c
        lb = np.array([v.lb for v in self.vars])
        ub = np.array([v.ub for v in self.vars])
        integ = np.array([1 if v.integer else 0 for v in self.vars], dtype=np.uint8)
        return A, np.asarray(lo), np.asarray(hi), obj, lb, ub, integ

    def Solve(self):
        A, lo, hi, obj, lb, ub, integ = self.export()
        sign = -1.0 if self._objective.maximize else 1.0
        opts = {"mip_rel_gap": 0.0, "disp": False}
        if Solver.time_limit is not None:
            opts["time_limit"] = float(Solver.time_limit)
        t0 = time.time()
        res = milp(sign * obj, constraints=LinearConstraint(A, lo, hi), integrality=integ,
                   bounds=Bounds(lb, ub), options=opts)
        if res.status == 2:
            # HiGHS' presolve wrongly declares some big-M models infeasible (1x1 simulated case:
            # x=c=1 is feasible); an "infeasible" verdict only counts with presolve off.
            res = milp(sign * obj, constraints=LinearConstraint(A, lo, hi), integrality=integ,
                       bounds=Bounds(lb, ub), options=dict(opts, presolve=False))
        dt = time.time() - t0
        self.last_mip_gap = getattr(res, "mip_gap", None)
        self.last_dual_bound = getattr(res, "mip_dual_bound", None)
        if res.status == 0 and res.x is not None:
            self._solution = res.x
            status = Solver.OPTIMAL
        elif res.status == 2:
            self._solution = None
            status = Solver.INFEASIBLE
        elif res.status == 3:
            self._solution = None
            status = Solver.UNBOUNDED
        elif res.x is not None:  # time limit with incumbent
            self._solution = res.x
            status = Solver.FEASIBLE
        else:
            self._solution = None
            status = Solver.NOT_SOLVED
        self.last_status = status
        Solver.solve_log.append({"cols": A.shape[1], "rows": A.shape[0], "nnz": int(A.nnz),
                                 "status": status, "seconds": dt,
                                 "mip_gap": self.last_mip_gap,
                                 "dual_bound": self.last_dual_bound})
        return status
