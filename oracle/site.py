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
