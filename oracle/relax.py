"""TEST INFRASTRUCTURE ONLY -- LP relaxations of the step-1 min-delay model by HiGHS (`scipy.optimize.linprog`).

`strengthened_lp(a, slot_cut)` is the relaxation the device solves matrix-free (csrc/pdhg_mf.cu): the reference
rows C2 (memory, `constraints_step1.py:18-23`), C3 (`:47-53`), C4 (`:57-65`) plus the valid rows x[i,f,j] <= c[f,j];
with `slot_cut` the memory rows are replaced by their Chvatal-Gomory rounding  sum_f c[f,j] <= floor(Mj / m)
(one memory size m per instance).  Returns (optimum, c-bar[F,N], CPU-row duals[N])."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


def strengthened_lp(a, slot_cut=True):
    N, F = a["N"], a["F"]
    d, w, r, m, Mj, Kj = a["d"], a["w"], a["r"], a["m"], a["Mj"], a["Kj"]
    X, C = F * N * N, F * N
    obj = np.concatenate([(w[:, :, None] * d[None, :, :]).ravel(), np.zeros(C)])
    f, i, j = np.meshgrid(np.arange(F), np.arange(N), np.arange(N), indexing="ij")
    xi, ci = (f * N * N + i * N + j).ravel(), (X + f * N + j).ravel()
    A1 = sp.coo_matrix((np.r_[np.ones(X), -np.ones(X)], (np.r_[np.arange(X), np.arange(X)], np.r_[xi, ci])),
                       shape=(X, X + C))
    ff, jj = np.meshgrid(np.arange(F), np.arange(N), indexing="ij")
    if slot_cut:
        assert np.all(m == m[0])
        A2 = sp.coo_matrix((np.ones(C), (jj.ravel(), (X + ff * N + jj).ravel())), shape=(N, X + C))
        b2 = np.minimum(np.floor(Mj / m[0] + 1e-9), F)
    else:
        A2 = sp.coo_matrix((np.repeat(m, N), (jj.ravel(), (X + ff * N + jj).ravel())), shape=(N, X + C))
        b2 = Mj
    A3 = sp.coo_matrix(((w[:, :, None] * r[:, None, :]).ravel(), (j.ravel(), xi)), shape=(N, X + C))
    A_ub = sp.vstack([A1, A2, A3]).tocsr()
    b_ub = np.r_[np.zeros(X), b2, Kj]
    A_eq = sp.coo_matrix((np.ones(X), ((f * N + i).ravel(), xi)), shape=(F * N, X + C)).tocsr()
    res = linprog(obj, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=np.ones(F * N), bounds=(0, 1), method="highs")
    assert res.status == 0
    return float(res.fun), res.x[X:].reshape(F, N), -res.ineqlin.marginals[X + N:]
