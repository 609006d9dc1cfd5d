// route_lp.cu -- EXACT routing of a fixed placement: the LP the reference's MIP reduces to once c is fixed.
//
// With c fixed, step 1 (reference constraints_step1.py:47-65, objective objectives.py:4-11) is
//     min sum d[i,j] w[f,i] x[i,f,j]   s.t.  sum_j x[i,f,j] = 1 over the open pods of f,
//                                            sum_{f,i} w[f,i] r[f,j] x[i,f,j] <= K_j,   x >= 0.
// Without binding CPU rows the optimum is the nearest-open-pod rule (efttc_step1.py:196-212).  When rows
// bind, this kernel solves the LP exactly with a dense-tableau DUAL simplex restricted to the "active"
// sources: the nearest-pod routing is dual feasible (all reduced costs d[i,j] - d[i,near] >= 0) and primal
// infeasible exactly on the overloaded rows, so no phase 1 is needed.  Sources whose nearest pod sits on a
// node outside the active set A keep that pod (its price is 0, so by complementary slackness nothing else
// can be cheaper); A starts as the overloaded nodes and grows by every node the optimum makes tight, or to
// all nodes when the restricted LP is infeasible.  At termination primal feasibility, dual feasibility and
// complementary slackness hold for the whole LP, i.e. the value is the LP optimum (checked against HiGHS in
// tests/test_route_lp_gpu.py to 1e-9).  Pods that end up with a total share below 1 - eps (C1b,
// constraints_step1.py:12-15) are topped up by zero-workload sources (free columns of the MIP) or reported.
//
// One block per placement (persistent: a block loops over placements), tableau in a per-block slab of the
// caller's workspace.  Pivot rule: most negative right-hand side, lowest index among equals; entering
// column by minimum ratio, lowest index among equals (fixed order => bit-reproducible).
#include "common.cuh"

namespace neptune {

struct RlpArgs {
  int B, P, N, F;
  const double *d, *w, *r, *Kj;
  const uint8_t* c;      // [B][P][F][N]
  uint8_t* c_out;        // [B][P][F][N] or null
  double* x_out;         // [B][P][N][F][N] or null
  double* n_out;         // [B][P][N] or null
  double* obj;           // [B][P]
  int32_t* status;       // [B][P]: 1 exact & feasible, 0 infeasible, 2 tableau does not fit the workspace slab
  int32_t* info;         // [B][P][2] or null: pivots, final |A|
  char* ws;
  int64_t slab_bytes;    // per block
  int64_t tab_doubles;   // capacity of the tableau region of a slab
};

constexpr double kRlpFeasTol = 1e-9;    // primal feasibility of a basic value
constexpr double kRlpPivTol = 1e-11;    // smallest admissible pivot magnitude

__host__ __device__ inline int64_t rlp_fixed_bytes(int N, int F) {
  const int64_t fn = (int64_t)F * N;
  // near, src, coloff (fn+1), podlist: 4 * fn ints (+1); colsrc / basis live after the tableau sizing below
  int64_t b = (5 * fn + 8 + 2 * F + N) * 4 + ((fn + 7) & ~(int64_t)7);   // + C1b pod list, fully-active functions, working copy of the placement
  b = (b + 7) & ~(int64_t)7;
  b += ((int64_t)2 * N + 2 * fn + N) * 8;      // load, colq (rows <= fn sources + N nodes + fn C1b rows), ...
  b += (int64_t)N * fn * 8;                    // dense x scratch
  return (b + 255) & ~(int64_t)255;
}

// argmin with index over the block (lowest index among equals).  All threads call; result valid everywhere.
__device__ inline void block_argmin(double v, int idx, double* s_val, int* s_idx, double& out_v, int& out_i) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v2 = __shfl_xor_sync(0xffffffffu, v, o); const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    if (v2 < v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
  }
  __syncthreads();
  if (lane == 0) { s_val[wid] = v; s_idx[wid] = idx; }
  __syncthreads();
  double bv = s_val[0]; int bi = s_idx[0];
  for (int k = 1; k < nw; ++k) if (s_val[k] < bv || (s_val[k] == bv && s_idx[k] < bi)) { bv = s_val[k]; bi = s_idx[k]; }
  out_v = bv; out_i = bi;
}

__global__ void __launch_bounds__(256) k_route_lp(RlpArgs a) {
  const int N = a.N, F = a.F, tid = threadIdx.x, nt = blockDim.x;
  const int64_t fn = (int64_t)F * N;
  __shared__ double s_val[8];
  __shared__ int s_idx[8];
  __shared__ double red[32];
  __shared__ int s_nS, s_nC, s_flag, s_grow, s_nA, s_nP;
  // ---- carve the slab --------------------------------------------------------------------------------
  char* p = a.ws + (int64_t)blockIdx.x * a.slab_bytes;
  int* near = (int*)p; p += fn * 4;
  int* src = (int*)p; p += fn * 4;
  int* coloff = (int*)p; p += (fn + 1) * 4;
  int* podlist = (int*)p; p += fn * 4;
  int* npods = (int*)p; p += (int64_t)F * 4;
  int* inA = (int*)p; p += (int64_t)N * 4;
  int* c1b = (int*)p; p += fn * 4;          // pods (f*N + j) that get a C1b row: sum_i x[i,f,j] >= 1 - eps
  int* fullF = (int*)p; p += (int64_t)F * 4; // functions whose every source is in the LP (they own a C1b row)
  uint8_t* cw = (uint8_t*)p; p += fn;       // working copy of the placement: starved pods are closed and the LP re-solved
  p = (char*)(((uintptr_t)p + 7) & ~(uintptr_t)7);
  double* load = (double*)p; p += (int64_t)N * 8;
  p += (int64_t)N * 8;
  double* colq = (double*)p; p += (2 * fn + N) * 8;       // one entry per tableau row: sources + nodes + C1b rows
  double* xs = (double*)p; p += (int64_t)N * fn * 8;
  p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  // tableau region: T[R][C] | red[C] | colsrc[C] (int) | basis[R] (int)
  double* Treg = (double*)p;

  int attempt = 0;
  for (int64_t job = blockIdx.x; job < (int64_t)a.B * a.P;) {
    const int b = (int)(job / a.P);
    const double* d = a.d + (int64_t)b * N * N;
    const double* w = a.w + b * fn;
    const double* r = a.r + b * fn;
    const double* K = a.Kj + (int64_t)b * N;
    __syncthreads();
    if (attempt == 0) {
      for (int q = tid; q < (int)fn; q += nt) cw[q] = a.c[job * fn + q];
      for (int f = tid; f < F; f += nt) fullF[f] = 0;
      if (tid == 0) s_nP = 0;
    }
    const uint8_t* c = cw;
    __syncthreads();
    if (tid == 0) s_flag = 1;
    __syncthreads();
    // pod lists
    for (int f = tid; f < F; f += nt) {
      int n = 0;
      for (int j = 0; j < N; ++j) if (c[(int64_t)f * N + j]) podlist[(int64_t)f * N + n++] = j;
      npods[f] = n;
      if (n == 0) s_flag = 0;
    }
    __syncthreads();
    int status = 1, pivots = 0;
    double cost = INFINITY;
    if (!s_flag) status = 0;
    if (status) {
      // nearest open pod (lowest index among equals)
      for (int fi = tid; fi < (int)fn; fi += nt) {
        const int f = fi / N, i = fi - f * N;
        const double* di = d + (int64_t)i * N;
        const int* pl = podlist + (int64_t)f * N;
        double best = INFINITY; int bj = -1;
        for (int q = 0; q < npods[f]; ++q) { const double v = di[pl[q]]; if (v < best) { best = v; bj = pl[q]; } }
        near[fi] = bj;
      }
      __syncthreads();
      for (int j = tid; j < N; j += nt) {
        double s = 0.0;
        for (int f = 0; f < F; ++f) {
          if (!c[(int64_t)f * N + j]) continue;
          const double rfj = r[(int64_t)f * N + j];
          double t = 0.0;
          for (int i = 0; i < N; ++i) if (near[(int64_t)f * N + i] == j) t += w[(int64_t)f * N + i];
          s += t * rfj;
        }
        load[j] = s;
        inA[j] = s > K[j] + kRlpFeasTol ? 1 : 0;
      }
      __syncthreads();
      if (tid == 0) { int n = 0; for (int j = 0; j < N; ++j) n += inA[j]; s_nA = n; }
      __syncthreads();
    }
    int nS = 0, nC = 0, R = 0, C = 0;
    double* T = Treg; double* redc = nullptr; int* colsrc = nullptr; int* basis = nullptr;
    bool solved_lp = false, c1b_infeasible = false;
    while (status == 1 && (s_nA > 0 || s_nP > 0)) {
      // ---- active sources and their columns (ordered: thread 0) ---------------------------------------
      if (tid == 0) {
        int ns = 0, nc = 0;
        for (int fi = 0; fi < (int)fn; ++fi) {
          if ((w[fi] > 0.0 && inA[near[fi]]) || fullF[fi / N]) { src[ns] = fi; coloff[ns] = nc; nc += npods[fi / N]; ++ns; }
        }
        coloff[ns] = nc;
        s_nS = ns; s_nC = nc;
      }
      __syncthreads();
      const int nP = s_nP;
      nS = s_nS; nC = s_nC; R = nS + N + nP; C = nC + N + nP + 1;
      const int64_t need = (int64_t)R * C + C + (C + R + 2) / 2 + 2;
      if (need > a.tab_doubles) { status = 2; break; }
      redc = T + (int64_t)R * C;
      colsrc = (int*)(redc + C);
      basis = colsrc + C;
      for (int64_t k = tid; k < (int64_t)R * C; k += nt) T[k] = 0.0;
      for (int k = tid; k < C; k += nt) redc[k] = 0.0;
      __syncthreads();
      // remaining capacity once the inactive sources sit on their nearest pods = K - load (the active ones are
      // moved out of the row by the canonical form below)
      for (int s = tid; s < nS; s += nt) {
        const int fi = src[s], f = fi / N, i = fi - f * N, nj = near[fi];
        const double wv = w[fi];
        const double anj = wv * r[(int64_t)f * N + nj], dnj = d[(int64_t)i * N + nj];
        const int* pl = podlist + (int64_t)f * N;
        for (int q = 0; q < npods[f]; ++q) {
          const int col = coloff[s] + q, j = pl[q];
          colsrc[col] = s;
          T[(int64_t)s * C + col] = 1.0;
          if (j != nj) {
            T[(int64_t)(nS + j) * C + col] = wv * r[(int64_t)f * N + j];
            T[(int64_t)(nS + nj) * C + col] = -anj;
            redc[col] = wv * (d[(int64_t)i * N + j] - dnj);
            // C1b rows of this function's starved pods, in canonical form: -sum_s x[s, jp] + surplus = -(1 - eps) with
            // the basic x[s, near(s)] = 1 - (the other columns of s) substituted
            for (int pp = 0; pp < nP; ++pp) {
              const int fj = c1b[pp];
              if (fj / N != f) continue;
              const int jp = fj - f * N;
              if (nj == jp) T[(int64_t)(nS + N + pp) * C + col] = 1.0;
              else if (j == jp) T[(int64_t)(nS + N + pp) * C + col] = -1.0;
            }
          } else {
            basis[s] = col;
          }
        }
        T[(int64_t)s * C + (C - 1)] = 1.0;
      }
      for (int pp = tid; pp < nP; pp += nt) {
        const int fj = c1b[pp], f = fj / N, jp = fj - f * N;
        int cnt = 0;
        for (int i = 0; i < N; ++i) cnt += near[(int64_t)f * N + i] == jp;          // every source of f is active (fullF)
        T[(int64_t)(nS + N + pp) * C + nC + N + pp] = 1.0;
        T[(int64_t)(nS + N + pp) * C + (C - 1)] = (double)cnt - (1.0 - kEps);
        basis[nS + N + pp] = nC + N + pp;
        colsrc[nC + N + pp] = -1;
      }
      for (int j = tid; j < N; j += nt) {
        T[(int64_t)(nS + j) * C + nC + j] = 1.0;
        T[(int64_t)(nS + j) * C + (C - 1)] = K[j] - load[j];
        basis[nS + j] = nC + j;
        colsrc[nC + j] = -1;
      }
      __syncthreads();
      // ---- dual simplex ----------------------------------------------------------------------------
      bool infeasible = false;
      const int max_piv = 50 * R + 100;
      int it = 0;
      for (; it < max_piv; ++it) {
        double mv = INFINITY; int mi = 0x7fffffff;
        for (int rr = tid; rr < R; rr += nt) { const double v = T[(int64_t)rr * C + (C - 1)]; if (v < mv) { mv = v; mi = rr; } }
        double bv; int rr;
        block_argmin(mv, mi, s_val, s_idx, bv, rr);
        if (!(bv < -kRlpFeasTol)) break;                       // primal feasible: optimal
        const double* Tr = T + (int64_t)rr * C;
        double qv = INFINITY; int qi = 0x7fffffff;
        for (int col = tid; col < C - 1; col += nt) {
          const double e = Tr[col];
          if (e < -kRlpPivTol) { const double ratio = redc[col] / (-e); if (ratio < qv || (ratio == qv && col < qi)) { qv = ratio; qi = col; } }
        }
        double rv; int q;
        block_argmin(qv, qi, s_val, s_idx, rv, q);
        if (q == 0x7fffffff) { infeasible = true; break; }
        // pivot on (rr, q)
        const double piv = Tr[q];
        for (int row = tid; row < R; row += nt) colq[row] = row == rr ? 0.0 : T[(int64_t)row * C + q];
        const double rq = redc[q];
        __syncthreads();
        double* Trw = T + (int64_t)rr * C;
        for (int col = tid; col < C; col += nt) Trw[col] = Trw[col] / piv;
        __syncthreads();
        for (int row = 0; row < R; ++row) {
          const double fct = colq[row];
          if (fct == 0.0) continue;
          double* To = T + (int64_t)row * C;
          for (int col = tid; col < C; col += nt) To[col] -= fct * Trw[col];
        }
        if (rq != 0.0) for (int col = tid; col < C - 1; col += nt) redc[col] -= rq * Trw[col];
        if (tid == 0) basis[rr] = q;
        __syncthreads();
      }
      pivots += it;
      if (it >= max_piv) { status = 2; break; }                 // did not converge (never seen): let the caller fall back
      if (infeasible) {
        if (s_nA < N) {
          for (int j = tid; j < N; j += nt) inA[j] = 1;
          if (tid == 0) s_nA = N;
          __syncthreads();
          continue;
        }
        if (nP > 0) { c1b_infeasible = true; break; }          // the C1b rows cannot be met: those pods are closed below
        status = 0; break;
      }
      // ---- nodes the optimum makes tight although their own sources are still pinned: widen A ----------
      if (tid == 0) s_grow = 0;
      __syncthreads();
      for (int j = tid; j < N; j += nt) {
        if (inA[j] || !(load[j] > 0.0)) continue;
        double val = 0.0;                                     // slack nonbasic => tight
        for (int row = 0; row < R; ++row) if (basis[row] == nC + j) { val = T[(int64_t)row * C + (C - 1)]; break; }
        if (val <= kRlpFeasTol) { inA[j] = 1; atomicAdd(&s_grow, 1); }
      }
      __syncthreads();
      if (s_grow > 0) { if (tid == 0) s_nA += s_grow; __syncthreads(); continue; }
      solved_lp = true;
      break;
    }
    if (c1b_infeasible) {
      // no routing gives the starved pods their share: close them and price the smaller placement
      __syncthreads();
      for (int pp = tid; pp < s_nP; pp += nt) cw[c1b[pp]] = 0;
      for (int f = tid; f < F; f += nt) fullF[f] = 0;
      __syncthreads();
      if (tid == 0) s_nP = 0;
      attempt = attempt < 2 ? 2 : attempt + 1;
      __syncthreads();
      if (attempt <= 5) continue;
      status = 0;
    }
    // ---- solution: dense x in the slab, cost, C1b repair, outputs ---------------------------------------
    int bad = 0;
    if (status == 1) {
      for (int64_t k = tid; k < (int64_t)N * fn; k += nt) xs[k] = 0.0;
      __syncthreads();
      for (int fi = tid; fi < (int)fn; fi += nt) {
        const int f = fi / N, i = fi - f * N;
        const bool active = solved_lp && ((w[fi] > 0.0 && inA[near[fi]]) || fullF[f]);
        if (!active) xs[((int64_t)i * F + f) * N + near[fi]] = 1.0;
      }
      if (solved_lp) {
        for (int row = tid; row < R; row += nt) {
          const int col = basis[row];
          if (col < 0 || col >= nC) continue;
          const double v = T[(int64_t)row * C + (C - 1)];
          if (!(v > 0.0)) continue;
          const int s = colsrc[col];
          if (s < 0 || s >= nS) continue;                        // (defensive: indices read back from the slab)
          const int fi = src[s], f = fi / N, i = fi - f * N, qq = col - coloff[s];
          if (fi < 0 || fi >= (int)fn || qq < 0 || qq >= npods[f]) continue;
          const int j = podlist[(int64_t)f * N + qq];
          xs[((int64_t)i * F + f) * N + j] = v > 1.0 ? 1.0 : v;
        }
      }
      __syncthreads();
      // cost in a fixed order (per source, pods ascending; block tree on top)
      double part = 0.0;
      for (int fi = tid; fi < (int)fn; fi += nt) {
        const double wv = w[fi];
        if (wv == 0.0) continue;
        const int f = fi / N, i = fi - f * N;
        const double* xr = xs + ((int64_t)i * F + f) * N;
        const int* pl = podlist + (int64_t)f * N;
        double t = 0.0;
        for (int q = 0; q < npods[f]; ++q) { const double xv = xr[pl[q]]; if (xv != 0.0) t += xv * d[(int64_t)i * N + pl[q]]; }
        part += wv * t;
      }
      part = block_sum(part, red);
      __syncthreads();
      if (tid == 0) red[0] = part;
      __syncthreads();
      cost = red[0];
      __syncthreads();
      // C1b: a pod with a total share in (0, 1 - eps) is topped up by zero-workload sources of its function
      // (their columns are free in the MIP: no delay cost, no CPU load); the donor pod must keep >= 1.
      for (int f = tid; f < F; f += nt) {
        for (int jq = 0; jq < npods[f]; ++jq) {
          const int j = podlist[(int64_t)f * N + jq];
          double rc = 0.0;
          for (int i = 0; i < N; ++i) rc += xs[((int64_t)i * F + f) * N + j];
          if (rc == 0.0 || rc + kEps >= 1.0) continue;
          double deficit = 1.0 - rc;
          for (int i = 0; i < N && deficit > 0.0; ++i) {
            if (w[(int64_t)f * N + i] != 0.0) continue;
            double* xr = xs + ((int64_t)i * F + f) * N;
            for (int q2 = 0; q2 < npods[f] && deficit > 0.0; ++q2) {
              const int p2 = podlist[(int64_t)f * N + q2];
              if (p2 == j || xr[p2] <= 0.0) continue;
              double rp = 0.0;
              for (int i2 = 0; i2 < N; ++i2) rp += xs[((int64_t)i2 * F + f) * N + p2];
              const double amt = fmin(fmin(xr[p2], rp - 1.0), deficit);
              if (amt <= 0.0) continue;
              xr[p2] -= amt; xr[j] += amt; deficit -= amt;
            }
          }
        }
      }
      __syncthreads();
      if (tid == 0) s_flag = 0;
      __syncthreads();
      uint8_t* cout = a.c_out ? a.c_out + job * fn : nullptr;
      for (int fj = tid; fj < (int)fn; fj += nt) {
        uint8_t on = c[fj];
        if (on) {
          const int f = fj / N, j = fj - f * N;
          double rc = 0.0;
          for (int i = 0; i < N; ++i) rc += xs[((int64_t)i * F + f) * N + j];
          if (rc == 0.0) on = 0;                               // serves nobody: closed (free for the objective)
          else if (rc + kEps < 1.0) s_flag = 1;                // still starved: reported
        }
        if (cout) cout[fj] = on;
      }
      __syncthreads();
      bad = s_flag;
      if (bad && attempt < 2) {
        // C1b is violated at the LP optimum of the capacity rows alone (constraints_step1.py:12-15): give every such
        // pod its row  sum_i x[i,f,j] >= 1 - eps  (all sources of its function enter the LP, the zero-workload ones as
        // free columns) and solve again -- the MIP's own treatment
        if (tid == 0) {
          int np_ = s_nP;
          for (int fj = 0; fj < (int)fn; ++fj) {
            if (!cw[fj]) continue;
            const int f = fj / N, j = fj - f * N;
            double rc = 0.0;
            for (int i = 0; i < N; ++i) rc += xs[((int64_t)i * F + f) * N + j];
            if (rc != 0.0 && rc + kEps < 1.0) {
              bool have = false;
              for (int pp = 0; pp < np_; ++pp) have = have || c1b[pp] == fj;
              if (!have) { c1b[np_++] = fj; fullF[f] = 1; }
            }
          }
          s_nP = np_;
        }
        ++attempt;
        __syncthreads();
        continue;
      }
      if (bad && attempt < 5) {
        if (tid == 0) s_nP = 0;
        for (int f = tid; f < F; f += nt) fullF[f] = 0;
        // C1b cannot be met for some pod at this routing (constraints_step1.py:12-15): close the starved pods and
        // solve the LP of the smaller placement -- still a feasible point of the MIP, at the LP value of that placement
        for (int fj = tid; fj < (int)fn; fj += nt) {
          if (!cw[fj]) continue;
          const int f = fj / N, j = fj - f * N;
          double rc = 0.0;
          for (int i = 0; i < N; ++i) rc += xs[((int64_t)i * F + f) * N + j];
          if (rc == 0.0 || rc + kEps < 1.0) cw[fj] = 0;
        }
        ++attempt;
        __syncthreads();
        continue;
      }
      if (a.x_out) {
        double* xo = a.x_out + job * (int64_t)N * fn;
        for (int64_t k = tid; k < (int64_t)N * fn; k += nt) xo[k] = xs[k];
      }
      if (a.n_out) {
        for (int j = tid; j < N; j += nt) {
          int any = 0;
          for (int f = 0; f < F; ++f) {
            if (!c[(int64_t)f * N + j]) continue;
            double rc = 0.0;
            for (int i = 0; i < N; ++i) rc += xs[((int64_t)i * F + f) * N + j];
            any |= rc != 0.0;
          }
          a.n_out[job * N + j] = any ? 1.0 : 0.0;
        }
      }
    } else {
      if (a.c_out) for (int fj = tid; fj < (int)fn; fj += nt) a.c_out[job * fn + fj] = c[fj];
      if (a.x_out) { double* xo = a.x_out + job * (int64_t)N * fn; for (int64_t k = tid; k < (int64_t)N * fn; k += nt) xo[k] = 0.0; }
      if (a.n_out) for (int j = tid; j < N; j += nt) a.n_out[job * N + j] = 0.0;
    }
    if (tid == 0) {
      a.obj[job] = status == 1 ? cost : INFINITY;
      a.status[job] = status == 1 ? (bad ? 0 : 1) : status;
      if (a.info) { a.info[job * 2] = pivots; a.info[job * 2 + 1] = status == 1 ? s_nA : -1; }
    }
    __syncthreads();
    job += gridDim.x; attempt = 0;
  }
}

}  // namespace neptune

using namespace neptune;

static int rlp_blocks(int64_t jobs) { return (int)(jobs < 2 * kNumSMs ? jobs : 2 * kNumSMs); }

extern "C" int neptune_route_lp_workspace_bytes(int B, int P, int N, int F, int64_t tableau_doubles, int64_t* bytes) {
  if (B <= 0 || P <= 0 || N <= 0 || F <= 0 || tableau_doubles <= 0 || !bytes) return NEPTUNE_E_ARG;
  const int64_t slab = rlp_fixed_bytes(N, F) + 256 + ((tableau_doubles * 8 + 255) & ~(int64_t)255);
  *bytes = slab * rlp_blocks((int64_t)B * P);
  return 0;
}

extern "C" int neptune_route_lp(int B, int P, int N, int F, const double* d, const double* w, const double* r,
                                const double* Kj, const uint8_t* c, uint8_t* c_out, double* x, double* n,
                                double* obj_out, int32_t* status_out, int32_t* info_out, int64_t tableau_doubles,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  if (B <= 0 || P <= 0 || N <= 0 || F <= 0 || !d || !w || !r || !Kj || !c || !obj_out || !status_out || !workspace ||
      tableau_doubles <= 0)
    return NEPTUNE_E_ARG;
  if ((int64_t)F * N > (1 << 24)) return NEPTUNE_E_SIZE;
  int64_t need = 0;
  neptune_route_lp_workspace_bytes(B, P, N, F, tableau_doubles, &need);
  if (workspace_bytes < need) return NEPTUNE_E_NOMEM;
  RlpArgs a{};
  a.B = B; a.P = P; a.N = N; a.F = F; a.d = d; a.w = w; a.r = r; a.Kj = Kj; a.c = c; a.c_out = c_out; a.x_out = x;
  a.n_out = n; a.obj = obj_out; a.status = status_out; a.info = info_out; a.ws = (char*)workspace;
  const int blocks = rlp_blocks((int64_t)B * P);
  a.slab_bytes = need / blocks;
  a.tab_doubles = tableau_doubles;
  { k_route_lp<<<blocks, 256, 0, (cudaStream_t)stream>>>(a); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  return 0;
}
