"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the round-robin delay-improvement greedy of csrc/site.cu.

The move is EFTTC's `find_best_node_by_delay_improvement` (reference `core/solvers/efttc/efttc_step1.py:214-288`):
    gain(f, j) = sum_i w[f,i] * max(0, cur[f,i] - d[i,j]),  cur[f,i] = delay from i to its nearest pod of f,
with the memory test of `can_assign` (`:290-312`).  Every function proposes its best node per round (lowest index among
equal gains); proposals are accepted in order of gain (lowest function among equals) while the node's memory lasts; a
function without workload still gets one pod (every source must be routed, `constraints_step1.py:27-33`).  Sums run
over i in ascending order with separate multiply and add, as the kernel does, so placements are compared bit for bit.
Only tests/ and tools/ import this."""
from __future__ import annotations

import numpy as np


def solve(a, max_rounds=None):
    N, F = a["N"], a["F"]
    d, w, m = a["d"], a["w"], a["m"]
    memfree = a["Mj"].astype(np.float64).copy()
    unserved = 2.0 * float(d.max()) + 1.0
    cur = np.full((F, N), unserved)
    c = np.zeros((F, N), dtype=np.uint8)
    npods = np.zeros(F, dtype=np.int64)
    rounds = pods = 0
    for _ in range(max_rounds or N * F):
        gain = np.zeros((F, N))
        for i in range(N):                                        # ascending i, multiply then add: the kernel's order
            gain = gain + w[:, i, None] * np.maximum(cur[:, i, None] - d[i][None, :], 0.0)
        fits = m[:, None] <= memfree[None, :] + 1e-9
        gain = np.where((c > 0) | ~fits, -1.0, gain)
        best_j = gain.argmax(axis=1)                              # first maximum = lowest node
        best_g = gain[np.arange(F), best_j]
        best_j = np.where(best_g < 0, -1, best_j)
        best_g = np.where((npods == 0) & (best_g == 0.0), 1e-300, best_g)
        placed = 0
        for _k in range(F):
            bf = -1; bv = 0.0
            for f in range(F):
                if best_g[f] > bv:
                    bv, bf = best_g[f], f
            if bf < 0:
                break
            best_g[bf] = -1.0
            j = int(best_j[bf])
            if j < 0 or m[bf] > memfree[j] + 1e-9:
                continue
            memfree[j] -= m[bf]
            c[bf, j] = 1
            cur[bf] = np.minimum(cur[bf], d[:, j])
            npods[bf] += 1
            placed += 1
        if not placed:
            break
        rounds += 1
        pods += placed
    return c, rounds, pods


def route_two_choice(a, c, max_iters=500):
    """numpy statement of `neptune_route_two_choice` (csrc/site.cu): nearest / second-nearest pod per source, the share
    theta[j] of the flows whose nearest pod is on node j lowered from 1 until `constrain_CPU_usage`
    (`constraints_step1.py:57-65`) holds; pods that are nobody's nearest are closed first.  Float loads here, fixed
    point (2^-30) on the device: compare objectives to ~1e-7.  Returns (c_out, x[N,F,N], objective, feasible, iterations)."""
    N, F = a["N"], a["F"]
    d, w, r, K = a["d"], a["w"], a["r"], a["Kj"]
    c = (np.asarray(c) > 0).astype(np.uint8).copy()

    def two_nearest():
        j1 = np.full((F, N), -1); j2 = np.full((F, N), -1)
        for f in range(F):
            pods = np.flatnonzero(c[f])
            if pods.size == 0:
                continue
            dd = d[:, pods]                                         # [i, pod]; stable argsort = lowest node among equals
            order = np.argsort(dd, axis=1, kind="stable")
            j1[f] = pods[order[:, 0]]
            if pods.size > 1:
                j2[f] = pods[order[:, 1]]
        return j1, j2

    j1, j2 = two_nearest()
    prim = np.zeros((F, N), dtype=np.int64)
    for f in range(F):
        np.add.at(prim[f], j1[f][j1[f] >= 0], 1)
    c[(c > 0) & (prim == 0)] = 0
    j1, j2 = two_nearest()
    ff = np.repeat(np.arange(F), N).reshape(F, N)
    act = (w > 0) & (j1 >= 0)
    a1 = np.where(act, w * r[ff, np.maximum(j1, 0)], 0.0)
    split = act & (j2 >= 0)
    P = np.zeros(N); Pfix = np.zeros(N)
    np.add.at(P, j1[split], a1[split]); np.add.at(Pfix, j1[act & ~split], a1[act & ~split])
    theta = np.ones(N)
    it = 0
    over = 0
    changed = True
    while it < max_iters and changed:
        S = np.zeros(N)
        sp = np.where(split, w * r[ff, np.maximum(j2, 0)] * (1.0 - theta[np.maximum(j1, 0)]), 0.0)
        np.add.at(S, j2[split], sp[split])
        cap = K * (1.0 - 1e-9)
        want = np.where(P > 0, (cap - S - Pfix) / np.where(P > 0, P, 1.0), 1.0)
        want = np.minimum(np.maximum(want, 0.0), theta)
        changed = bool((want < theta - 1e-12).any())
        over = int((P * want + Pfix + S > K * (1.0 + 1e-9) + 1e-9).sum())
        theta = np.where(want < theta - 1e-12, want, theta)
        it += 1
    x = np.zeros((N, F, N))
    obj = 0.0
    for f in range(F):
        for i in range(N):
            if j1[f, i] < 0:
                continue
            th = theta[j1[f, i]] if j2[f, i] >= 0 else 1.0
            x[i, f, j1[f, i]] = th
            if j2[f, i] >= 0 and th < 1.0:
                x[i, f, j2[f, i]] = 1.0 - th
            obj += w[f, i] * (th * d[i, j1[f, i]] + ((1.0 - th) * d[i, j2[f, i]] if j2[f, i] >= 0 else 0.0))
    return c, x, obj, over == 0, it
