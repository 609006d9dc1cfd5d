"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's EFTTC greedy on dense numpy state.

Follows `/root/reference/core/solvers/efttc/efttc_step1.py` (line numbers below) with the dict-of-
dicts state replaced by arrays `c[F,N]`, `n[N]`, `x[N,F,N]`; the control flow, iteration orders,
tie-breaks and tolerances are the reference's:

  solve loop `:39-90`  preference graph `:123-146`  find_cycle `:148-188`
  can_assign_cycle / can_assign `:290-312`  change_x_one `:196-212`  change_n_one `:190-194`
  handle_cycle `:92-121`  find_best_node_by_delay_improvement `:214-288`
  score_local x3 `:356-378, 397-410, 425-439`
  global checks: CPU `efttc/utils/constraints_step1.py:70-80`, budget `:126-133`.

`strict=True` reproduces the reference's `KeyError` (`remaining_functions.remove(f)` executed once
per cycle pair, `:118`); `strict=False` is the documented `discard` fallback.
"""
from __future__ import annotations

import numpy as np

from . import checkers

OBJECTIVES = {"min_delay": "min_delay", "min_util": "min_utilization",
              "min_delay_util": "min_delay_min_utilization"}


class EfttcResult:
    def __init__(self, x, c, n, iterations, would_raise):
        self.x, self.c, self.n = x, c, n
        self.iterations = iterations
        self.would_raise = would_raise


def _score_matrix(a, kind, alpha, c, wd):
    """score_local(f, j) for every pair, as an [F, N] array."""
    warm = np.where(a["old"] == 1, 0.5, 1.0)
    planned = c.sum(axis=0).astype(np.float64)                      # per node
    if kind == "min_util":
        actual = np.floor(a["old"].sum(axis=0))                     # int(np.sum(old[:, j]))
        return (a["cost"] / (1 + (planned + actual)))[None, :] * warm
    if kind == "min_delay":
        return wd * warm
    if kind == "min_delay_util":
        return (alpha * (a["cost"] / (1 + planned))[None, :] + (1 - alpha) * wd) * warm
    raise ValueError(kind)


def _find_cycle(graph):
    """`:148-188` verbatim in behaviour: functional-graph walk, both edge directions become pairs."""
    visited = set()
    for start in graph:
        if start in visited:
            continue
        path, current, local = [], start, set()
        while current not in local:
            local.add(current)
            path.append(current)
            if current not in graph:
                break
            nxt = graph[current]
            path.append(nxt)
            if nxt in local:
                k = path.index(nxt)
                cleaned, seen = [], set()
                for p in range(k, len(path) - 1):
                    u, v = path[p], path[p + 1]
                    if u >= 0 and v < 0:
                        pair = (u, ~v)
                    elif u < 0 and v >= 0:
                        pair = (v, ~u)
                    else:
                        continue
                    if pair not in seen:
                        seen.add(pair)
                        cleaned.append(pair)
                return cleaned
            current = nxt
        visited |= local
    return []


def _route_function(a, c, x, f):
    """change_x_one `:196-212`: every source splits equally among its (1e-6-)nearest active nodes."""
    active = np.flatnonzero(c[f])
    if active.size == 0:
        return
    dd = a["d"][:, active]                                           # [N, |A|]
    best = np.abs(dd - dd.min(axis=1, keepdims=True)) < 1e-6
    x[:, f, active] = np.where(best, 1.0 / best.sum(axis=1, keepdims=True), 0.0)


def _mem_used(a, c, j):
    total = 0                                                        # Python sum: sequential from int 0
    for f2 in range(a["F"]):
        total = total + (a["m"][f2] if c[f2, j] else 0)
    return total


def _has_improving_node(a, kind, alpha, c, n, invalid, f, remaining_nodes):
    """`find_best_node_by_delay_improvement(...) is not None` (`:214-288`)."""
    cand = [j for j in remaining_nodes if not c[f, j] and not invalid[f, j]]
    if not cand:
        return False
    active = np.flatnonzero(c[f])
    cur = a["d"][:, active].min(axis=1) if active.size else np.full(a["N"], np.inf)
    with np.errstate(invalid="ignore"):
        cur_score = np.sum(a["w"][f] * cur)
        best = 0.0
        found = False
        for j in cand:
            new_score = np.sum(a["w"][f] * np.minimum(cur, a["d"][:, j]))
            delta = cur_score - new_score
            if kind == "min_delay":
                if delta > best + 1e-6:
                    best, found = delta, True
            else:
                du = (1 / a["N"]) if not n[j] else 0
                ds = (1 - alpha) * delta - alpha * du
                if ds > best + 1e-6:
                    best, found = ds, True
    return found


def solve(a, kind, alpha=0.5, strict=True, max_iterations=None):
    N, F = a["N"], a["F"]
    c = np.zeros((F, N), dtype=bool)
    n = np.zeros(N, dtype=bool)
    x = np.zeros((N, F, N))
    invalid = np.zeros((F, N), dtype=bool)
    rem_f = set(range(F))
    rem_n = set(range(N))
    tried = set()
    wd = a["w"] @ a["d"] if kind != "min_util" else None             # wd[f,j] = sum_i d[i,j] w[f,i]
    uses_budget = kind != "min_delay"
    iterations = 0
    would_raise = False

    while rem_f:
        if max_iterations is not None and iterations >= max_iterations:
            break
        iterations += 1
        S = _score_matrix(a, kind, alpha, c, wd)
        graph = {}
        nodes_sorted = sorted(rem_n)
        funcs_sorted = sorted(rem_f)
        for f in funcs_sorted:
            valid = [j for j in nodes_sorted if not invalid[f, j]]
            if valid:
                graph[f] = ~min(valid, key=lambda j: (S[f, j], j))
        for j in nodes_sorted:
            graph[~j] = min(funcs_sorted, key=lambda f: (S[f, j], f))
        cycle = _find_cycle(graph)
        if not cycle:
            break
        key = tuple(sorted(cycle))
        if key in tried:
            break
        snap = (c.copy(), n.copy(), x.copy())
        success = False
        for f, j in cycle:
            if not (_mem_used(a, c, j) + a["m"][f] <= a["Mj"][j]):
                invalid[f, j] = True
                continue
            c[f, j] = True
            _route_function(a, c, x, f)
            n[j] = c[:, j].any()
            success = True
        if not success:
            tried.add(key)
            continue
        ok = checkers.check_cpu(a, x)
        if ok and uses_budget:
            ok = checkers.check_budget(a, n)
        if not ok:
            tried.add(key)
            c, n, x = snap
            for f, j in cycle:
                invalid[f, j] = True
            continue
        # handle_cycle `:92-121` -- the body runs once per pair of the cycle
        for _, j in cycle:
            if _mem_used(a, c, j) == a["Mj"][j]:
                rem_n.discard(j)
            for f2, j2 in cycle:
                invalid[f2, j2] = True
            if kind in ("min_delay", "min_delay_util"):
                for f2, _ in cycle:
                    if not _has_improving_node(a, kind, alpha, c, n, invalid, f2, sorted(rem_n)):
                        if f2 not in rem_f:
                            would_raise = True
                            if strict:
                                raise KeyError(f2)
                        rem_f.discard(f2)
            else:
                for f2, _ in cycle:
                    rem_f.discard(f2)
    return EfttcResult(x, c.astype(np.float64), n.astype(np.float64), iterations, would_raise)


def score(a, kind, alpha, res):
    if kind == "min_delay":
        return checkers.score_delay(a, res.x)
    if kind == "min_util":
        return checkers.score_util(a, res.n)
    return checkers.score_delay_util(a, res.n, res.x, alpha)
