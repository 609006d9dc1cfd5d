"""TEST INFRASTRUCTURE ONLY -- vectorised numpy/scipy restatement of the reference's MIP models.

The reference builds its models variable-by-variable through `pywraplp`; here the same matrix is
written down in closed form.  Canonical layout (proven equal to what the
reference's code emits -- `tests/test_oracle_vs_reference.py`):

  columns  x[i,f,j] -> f*N*N + i*N + j      `neptune/utils/variables.py:4-8`   continuous [0, +inf)
           c[f,j]   -> F*N*N + f*N + j      `variables.py:10-13`               binary
           n[j]     -> F*N*N + F*N + j      `variables.py:15-17`               binary (util models)
  rows     C1a/C1b interleaved over (f, j)  `constraints_step1.py:5-15`
           C2 (j)                           `constraints_step1.py:18-23`
           C3 (f, i)                        `constraints_step1.py:47-53`
           C4 (j)                           `constraints_step1.py:57-65`   (explicit zeros kept)
           C5a/C5b interleaved (j), C6 (j)  `constraints_step1.py:69-78,101-103`   (util models)
  objective  `neptune/utils/objectives.py:4-11` (min_delay), `:24-27` (min_util), `:30-52` (combined)
Step-2 rows/columns (`constraints_step2.py`, `variables.py:19-33`, `objectives.py:55-63`) are
appended by `build_step2`.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

BIG_M = 10 ** 6          # constraints_step1.py:1
EPSILON = 10 ** -6       # constraints_step1.py:2
INF = float("inf")

KINDS = ("min_delay", "min_util", "min_delay_util")


def arrays_from_data(data):
    """float64 views of the `Data` fields the model reads."""
    return dict(
        N=len(data.nodes), F=len(data.functions),
        d=np.asarray(data.node_delay_matrix, dtype=np.float64),
        w=np.asarray(data.workload_matrix, dtype=np.float64),
        r=np.asarray(data.core_per_req_matrix, dtype=np.float64),
        m=np.asarray(data.function_memory_matrix, dtype=np.float64),
        Mj=np.asarray(data.node_memory_matrix, dtype=np.float64),
        Kj=np.asarray(data.node_cores_matrix, dtype=np.float64),
        old=np.asarray(data.old_allocations_matrix, dtype=np.float64),
        maxd=np.asarray(data.max_delay_matrix, dtype=np.float64),
        cost=np.asarray(data.node_costs, dtype=np.float64),
        budget=float(data.node_budget),
    )


def model_sizes(N, F, with_n):
    cols = F * N * N + F * N + (N if with_n else 0)
    rows = 3 * F * N + 2 * N + (3 * N if with_n else 0)
    nnz = 4 * F * N * N + 3 * F * N + ((2 * N * (F + 1) + N) if with_n else 0)
    return rows, cols, nnz


def max_workload_delay(a):
    """`objectives.py:36-44`: sequential sum over (f, i) of w[f,i] * max{d[i,:] <= max_delay[f]}."""
    N, F, d, w, maxd = a["N"], a["F"], a["d"], a["w"], a["maxd"]
    total = 0.0
    for f in range(F):
        masked = np.where(d <= maxd[f], d, -INF).max(axis=1)     # per source i
        for i in range(N):
            total += w[f, i] * masked[i]
    return total


def objective_step1(a, kind, alpha):
    N, F, d, w = a["N"], a["F"], a["d"], a["w"]
    X, C = F * N * N, F * N
    with_n = kind != "min_delay"
    obj = np.zeros(X + C + (N if with_n else 0))
    if kind == "min_delay":
        # float(d[i,j] * w[f,i]) at column f*N*N + i*N + j
        obj[:X] = (d[None, :, :] * w[:, :, None]).reshape(-1)
    elif kind == "min_util":
        obj[X + C:] = 1.0
    elif kind == "min_delay_util":
        obj[X + C:] = float(alpha / N)
        if w.sum():
            wmax = max_workload_delay(a)
            with np.errstate(divide="ignore", invalid="ignore"):
                # float((1 - alpha) * workload * delay / max_workload_delay): left-to-right
                obj[:X] = ((((1 - alpha) * w)[:, :, None] * d[None, :, :]) / wmax).reshape(-1)
    else:
        raise ValueError(kind)
    return obj


def build_step1(a, kind, alpha=0.5):
    """Return dict(A (csr, sorted indices), lo, hi, obj, lb, ub, integ) for a step-1 model."""
    N, F = a["N"], a["F"]
    w, r, m = a["w"], a["r"], a["m"]
    with_n = kind != "min_delay"
    X, C = F * N * N, F * N
    rows, cols, nnz = model_sizes(N, F, with_n)

    ar_n = np.arange(N)
    indptr_parts, idx_parts, val_parts, lo_parts, hi_parts = [], [], [], [], []

    # ---- C1a / C1b, interleaved over (f, j) --------------------------------------------
    f_ = np.repeat(np.arange(F), N)                       # (f, j) flattened
    j_ = np.tile(ar_n, F)
    xcols = f_[:, None] * N * N + ar_n[None, :] * N + j_[:, None]          # [F*N, N] over i
    ccol = X + f_ * N + j_
    one_row_idx = np.concatenate([xcols, ccol[:, None]], axis=1)           # [F*N, N+1]
    idx_parts.append(np.repeat(one_row_idx, 2, axis=0).reshape(-1))
    va = np.ones((F * N, N + 1)); va[:, -1] = -float(BIG_M)
    vb = np.ones((F * N, N + 1)); vb[:, -1] = -1.0
    val_parts.append(np.stack([va, vb], axis=1).reshape(-1))
    indptr_parts.append(np.full(2 * F * N, N + 1))
    lo_parts.append(np.tile([-INF, -EPSILON], F * N))
    hi_parts.append(np.tile([0.0, INF], F * N))

    # ---- C2 (j): sum_f m[f] c[f,j] <= Mj ---------------------------------------------------
    idx_parts.append((X + np.arange(F)[None, :] * N + ar_n[:, None]).reshape(-1))
    val_parts.append(np.tile(m, N))
    indptr_parts.append(np.full(N, F))
    lo_parts.append(np.full(N, -INF)); hi_parts.append(a["Mj"].astype(np.float64))

    # ---- C3 (f, i): sum_j x[i,f,j] == 1 ----------------------------------------------------
    idx_parts.append(np.arange(X))
    val_parts.append(np.ones(X))
    indptr_parts.append(np.full(F * N, N))
    lo_parts.append(np.ones(F * N)); hi_parts.append(np.ones(F * N))

    # ---- C4 (j): sum_{f,i} w[f,i] r[f,j] x[i,f,j] <= Kj  (f-major, i-minor; zeros explicit) --
    fi = (np.arange(F)[:, None] * N * N + ar_n[None, :] * N).reshape(-1)   # [F*N]
    idx_parts.append((fi[None, :] + ar_n[:, None]).reshape(-1))
    coef = w[None, :, :] * r.T[:, :, None]                # [j, f, i] = w[f,i] * r[f,j]
    val_parts.append(coef.reshape(-1))
    indptr_parts.append(np.full(N, F * N))
    lo_parts.append(np.full(N, -INF)); hi_parts.append(a["Kj"].astype(np.float64))

    if with_n:
        # ---- C5a / C5b interleaved over j ---------------------------------------------------
        ccols = X + np.arange(F)[None, :] * N + ar_n[:, None]               # [N, F]
        ncol = X + C + ar_n
        one = np.concatenate([ccols, ncol[:, None]], axis=1)                # [N, F+1]
        idx_parts.append(np.repeat(one, 2, axis=0).reshape(-1))
        va = np.ones((N, F + 1)); va[:, -1] = -float(BIG_M)
        vb = np.ones((N, F + 1)); vb[:, -1] = -1.0
        val_parts.append(np.stack([va, vb], axis=1).reshape(-1))
        indptr_parts.append(np.full(2 * N, F + 1))
        lo_parts.append(np.tile([-INF, -EPSILON], N)); hi_parts.append(np.tile([0.0, INF], N))
        # ---- C6 (j): cost_j n[j] <= budget --------------------------------------------------
        idx_parts.append(ncol)
        val_parts.append(a["cost"].astype(np.float64))
        indptr_parts.append(np.full(N, 1))
        lo_parts.append(np.full(N, -INF)); hi_parts.append(np.full(N, a["budget"]))

    lens = np.concatenate(indptr_parts)
    indptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate(idx_parts).astype(np.int64)
    values = np.concatenate(val_parts).astype(np.float64)
    assert indptr[-1] == nnz == indices.size == values.size, (indptr[-1], nnz, indices.size)
    A = sp.csr_matrix((values, indices, indptr), shape=(rows, cols))

    lb = np.zeros(cols)
    ub = np.full(cols, INF); ub[X:] = 1.0
    integ = np.zeros(cols, dtype=np.uint8); integ[X:] = 1
    return dict(A=A, lo=np.concatenate(lo_parts), hi=np.concatenate(hi_parts),
                obj=objective_step1(a, kind, alpha), lb=lb, ub=ub, integ=integ,
                N=N, F=F, kind=kind, with_n=with_n)


def transpose_csr(A):
    """The stored transpose the GPU path also assembles: CSR of A^T with ascending row ids."""
    At = A.T.tocsr()
    At.sort_indices()
    return At
