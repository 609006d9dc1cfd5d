"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's feasibility checkers and scorers.

Follows `/root/reference/core/solvers/efttc/utils/constraints_step1.py:5-133` (the only explicit
definition of "feasible" in the reference; also what `testing/*/..._score_analysis.py` re-runs on MIP
answers) and `efttc/utils/objectives.py:23-36,48-49,53-98`.  Inputs are dense arrays
`x[N,F,N]` (index [i][f][j]), `c[F,N]`, `n[N]` instead of the reference's dicts of
{"name","val"}; every sum is taken in the reference's order (Python's left-to-right `sum`, here
`np.add.accumulate(...)[-1]`, which is sequential).
"""
from __future__ import annotations

import numpy as np

BIG_M = 1e6
EPSILON = 1e-6

FLAG_NAMES = ("c_x", "memory", "handle_all", "cpu", "n_c", "budget")


def _seq_sum(t, axis):
    """Left-to-right sum along `axis` (bit-identical to Python's sum over that axis)."""
    t = np.asarray(t, dtype=np.float64)
    if t.shape[axis] == 0:
        return np.zeros(np.delete(t.shape, axis))
    return np.take(np.add.accumulate(t, axis=axis), -1, axis=axis)


def check_c_x(a, x, c):                       # constraints_step1.py:5-18
    cb = np.asarray(c) > 0
    sum_x = _seq_sum(x, 0)                    # [F, N]: over sources i, ascending
    bad_hi = sum_x > np.where(cb, BIG_M, 0.0)
    bad_lo = sum_x + EPSILON < np.where(cb, 1.0, 0.0)
    return not bool((bad_hi | bad_lo).any())


def memory_used(a, c):
    cb = np.asarray(c) > 0
    return _seq_sum(np.where(cb, a["m"][:, None], 0.0), 0)      # [N], f ascending


def check_memory(a, c):                       # constraints_step1.py:22-33
    return not bool((memory_used(a, c) > a["Mj"]).any())


def check_handle_all(a, x, tol=1e-1):         # constraints_step1.py:37-47
    total = _seq_sum(x, 2)                    # [N(i), F]
    return bool((np.abs(total - 1) < tol).all())


def cpu_load(a, x):
    """total_j accumulated as `total += val * w[f,i] * r[f,j]`, f outer, i inner (`:70-77`)."""
    N, F = a["N"], a["F"]
    xf = np.transpose(np.asarray(x, dtype=np.float64), (1, 0, 2))          # [f, i, j]
    with np.errstate(over="ignore", invalid="ignore"):
        terms = (xf * a["w"][:, :, None]) * a["r"][:, None, :]             # (val*w)*r
    return _seq_sum(terms.reshape(F * N, N), 0)


def check_cpu(a, x):                          # constraints_step1.py:70-80
    return not bool((cpu_load(a, x) > a["Kj"] + 1e-6).any())


def check_n_c(a, n, c):                       # constraints_step1.py:85-95
    sum_c = (np.asarray(c) > 0).sum(axis=0)
    n_val = (np.asarray(n) > 0).astype(np.int64)
    bad = (sum_c > n_val * BIG_M) | (sum_c + EPSILON < n_val)
    return not bool(bad.any())


def check_budget(a, n):                       # constraints_step1.py:126-133
    nb = (np.asarray(n) > 0).astype(np.float64)
    total = _seq_sum(nb * a["cost"], 0)
    return not bool(total > a["budget"] + 1e-6)


def check_all(a, x, c, n):
    return {
        "c_x": check_c_x(a, x, c),
        "memory": check_memory(a, c),
        "handle_all": check_handle_all(a, x),
        "cpu": check_cpu(a, x),
        "n_c": check_n_c(a, n, c),
        "budget": check_budget(a, n),
    }


def flags_to_mask(flags):
    """bit k set <=> check FLAG_NAMES[k] passed."""
    return sum((1 << k) for k, name in enumerate(FLAG_NAMES) if flags[name])


# ---- scorers -----------------------------------------------------------------------------------
def score_delay(a, x):                        # objectives.py:23-36
    x = np.asarray(x, dtype=np.float64)
    return float(np.sum(x * a["d"][:, None, :] * a["w"].T[:, :, None]))


def score_util(a, n):                         # objectives.py:48-49
    return int((np.asarray(n) > 0).sum())


def score_delay_util(a, n, x, alpha):         # objectives.py:53-98
    N = a["N"]
    node_util = score_util(a, n) * (alpha / N) if N else 0.0
    w, d, maxd = a["w"], a["d"], a["maxd"]
    if np.sum(w) == 0:
        return node_util
    masked = np.where(d[None, :, :] <= maxd[:, None, None], d[None, :, :], 0)
    wmax = np.sum(w * masked.max(axis=2))
    if wmax == 0:
        return node_util
    x32 = np.asarray(x, dtype=np.float32)                  # the reference stores x as float32 (`:86`)
    contrib = x32 * w.T[:, :, None] * d[:, None, :]
    return float(node_util + np.sum(contrib) * (1 - alpha) / wmax)


def disruption_closed_form(a, c, mode):
    """Step-2 objective for a fixed placement (`objectives.py:55-63` + `constraints_step2.py:5-52`).

    With c fixed the optimal auxiliaries are moved_from = max(0, c-old), moved_to = max(0, old-c),
    and in the feasible direction (allocated, deallocated) = (0, -(sum_old-sum_c)) for "delete" and
    (-(sum_c-sum_old), 0) for "create"; the other direction is infeasible.  Returns None then.
    """
    cb = (np.asarray(c) > 0).astype(np.int64)
    old = (a["old"] > 0).astype(np.int64)
    W = old.size
    flips = int(np.abs(cb - old).sum())
    delta = int(old.sum() - cb.sum())
    if mode == "delete":
        if delta < 0:
            return None
        return float(W * flips - (W + 1) * delta)
    if mode == "create":
        if delta > 0:
            return None
        return float(W * flips - (W - 1) * (-delta))
    raise ValueError(mode)
