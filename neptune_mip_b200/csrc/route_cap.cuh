// route_cap.cuh -- capacity-aware routing of a fixed placement (block-cooperative device code).
//
// For a fixed c the reference's step-1 MIP reduces to an LP in x: every source i of function f
// spreads its requests over the open pods of f at minimum delay subject to the per-node CPU rows
// (constraints_step1.py:57-65).  Without binding CPU rows the optimum is the nearest-open-pod rule
// (efttc_step1.py:196-212).  When rows bind, the LP optimum moves the flows with the smallest
// delay increase per freed core  rho = (alt - d[i,j]) / r[f,j]  off the overloaded node, the last one
// fractionally -- a fractional knapsack per node, repeated with node prices so that flows moved onto
// a node that overloads in turn are re-priced (dual coordinate ascent on the CPU rows).  The HiGHS
// optima of the 20x5 / 50x10 instances have exactly this shape (1-2 binding nodes, one split flow per
// binding node); the result is re-verified by the exact checkers in any case.
#pragma once
#include "common.cuh"

namespace neptune {

struct CapRoute {
  int N, F;
  const double *d, *w, *r, *Kj;
  const uint8_t* c;       // [F][N]
  // per-placement work arrays (global memory)
  int* ch;                // [F*N] pod carrying the flow (f,i)            (-1: function has no pod)
  int* sec;               // [F*N] second pod of a split flow             (-1: not split)
  double* th;             // [F*N] fraction on `ch` (1 - th goes to `sec`)
  double* rho;            // [F*N] scratch
  int* alt;               // [F*N] scratch
  double* load;           // [N]
  double* lam;            // [N] node prices
};

struct CapResult { double cost; int feasible; int rounds; };

__device__ inline void cap_loads(const CapRoute& q) {
  const int N = q.N, F = q.F;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    double s = 0.0;
    for (int f = 0; f < F; ++f) {
      if (!q.c[(int64_t)f * N + j]) continue;
      const double rfj = q.r[(int64_t)f * N + j];
      const double* wf = q.w + (int64_t)f * N;
      const int* ch = q.ch + (int64_t)f * N; const int* sc = q.sec + (int64_t)f * N;
      const double* th = q.th + (int64_t)f * N;
      double t = 0.0;
      for (int i = 0; i < N; ++i) {
        if (ch[i] == j) t += th[i] * wf[i];
        else if (sc[i] == j) t += (1.0 - th[i]) * wf[i];
      }
      s += t * rfj;
    }
    q.load[j] = s;
  }
}

// `red` needs 32 doubles of shared memory, `sh` 32 doubles (first 16: one int per warp; rest: scalars).
// All threads of the block (up to 32 warps) must call.
__device__ inline CapResult cap_route(const CapRoute& q, int max_rounds, double* red, double* sh) {
  const int N = q.N, F = q.F, tid = threadIdx.x;
  __shared__ int s_j, s_ok;
  __syncthreads();
  if (tid == 0) s_ok = 1;
  for (int j = tid; j < N; j += blockDim.x) q.lam[j] = 0.0;
  __syncthreads();
  // nearest open pod (lowest index among equals)
  for (int fi = tid; fi < F * N; fi += blockDim.x) {
    const int f = fi / N, i = fi - f * N;
    const double* di = q.d + (int64_t)i * N;
    double best = INFINITY; int bj = -1;
    for (int j = 0; j < N; ++j) if (q.c[(int64_t)f * N + j] && di[j] < best) { best = di[j]; bj = j; }
    q.ch[fi] = bj; q.sec[fi] = -1; q.th[fi] = 1.0;
    if (bj < 0) s_ok = 0;
  }
  __syncthreads();
  CapResult res{INFINITY, 0, 0};
  if (!s_ok) return res;
  int round = 0;
  for (; round < max_rounds; ++round) {
    cap_loads(q);
    __syncthreads();
    // most overloaded node
    double worst = 0.0; int wj = -1;
    for (int j = tid; j < N; j += blockDim.x) {
      const double ex = q.load[j] - q.Kj[j];
      if (ex > worst) { worst = ex; wj = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double w2 = __shfl_xor_sync(0xffffffffu, worst, o); const int j2 = __shfl_xor_sync(0xffffffffu, wj, o);
      if (w2 > worst || (w2 == worst && j2 >= 0 && (wj < 0 || j2 < wj))) { worst = w2; wj = j2; }
    }
    if ((tid & 31) == 0) { red[tid >> 5] = worst; reinterpret_cast<int*>(sh)[tid >> 5] = wj; }
    __syncthreads();
    if (tid == 0) {
      double bw = 0.0; int bj = -1;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
        const double v = red[k]; const int jj = reinterpret_cast<int*>(sh)[k];
        if (v > bw || (v == bw && jj >= 0 && (bj < 0 || jj < bj))) { bw = v; bj = jj; }
      }
      s_j = (bw > 1e-9) ? bj : -1;
      sh[20] = bw;
    }
    __syncthreads();
    const int js = s_j;
    const double need = sh[20];
    if (js < 0) break;
    // candidates: unsplit flows on js; rho = delay increase per freed core at current prices
    double my_max = 0.0, my_mov = 0.0;
    for (int fi = tid; fi < F * N; fi += blockDim.x) {
      q.rho[fi] = -1.0;                                   // not a candidate
      if (q.ch[fi] != js || q.sec[fi] != -1) continue;       // split (>= 0) and frozen (-2) flows stay put
      const int f = fi / N, i = fi - f * N;
      const double rfj = q.r[(int64_t)f * N + js], wv = q.w[fi];
      if (!(wv * rfj > 0.0)) continue;
      const double* di = q.d + (int64_t)i * N;
      double best = INFINITY; int bj = -1;
      for (int j = 0; j < N; ++j) {
        if (j == js || !q.c[(int64_t)f * N + j]) continue;
        const double v = di[j] + q.lam[j] * q.r[(int64_t)f * N + j];
        if (v < best) { best = v; bj = j; }
      }
      if (bj < 0) continue;                               // the only pod of f: cannot leave
      const double rh = fmax((best - di[js]) / rfj, 0.0);
      q.rho[fi] = rh; q.alt[fi] = bj;
      my_max = fmax(my_max, rh); my_mov += wv * rfj;
    }
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) my_max = fmax(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
    if ((tid & 31) == 0) red[tid >> 5] = my_max;
    __syncthreads();
    if (tid == 0) { double m = 0.0; for (int k = 0; k < (int)(blockDim.x >> 5); ++k) m = fmax(m, red[k]); sh[21] = m; }
    __syncthreads();
    const double rho_max = sh[21];
    const double movable = block_sum(my_mov, red);
    __syncthreads();
    if (tid == 0) sh[22] = movable;
    __syncthreads();
    if (sh[22] < need * (1.0 - 1e-12)) { if (tid == 0) s_ok = 0; __syncthreads(); break; }
    // threshold t*: smallest t with freed(t) = sum_{rho <= t} w*r >= need   (bisection on a step function)
    double lo = -1e-300, hi = rho_max;                    // freed(lo) = 0 < need <= freed(hi)
    for (int it = 0; it < 64; ++it) {
      const double mid = 0.5 * (lo + hi);
      double fr = 0.0;
      for (int fi = tid; fi < F * N; fi += blockDim.x) {
        const double rh = q.rho[fi];
        if (rh >= 0.0 && rh <= mid) fr += q.w[fi] * q.r[(int64_t)(fi / N) * N + js];
      }
      fr = block_sum(fr, red);
      __syncthreads();
      if (tid == 0) sh[23] = fr;
      __syncthreads();
      if (sh[23] >= need) hi = mid; else lo = mid;
      if (hi - lo <= 1e-14 * fmax(1.0, hi)) break;
    }
    // flows with rho <= lo move completely, flows in (lo, hi] share the remainder
    double fr_lo = 0.0, fr_mg = 0.0;
    for (int fi = tid; fi < F * N; fi += blockDim.x) {
      const double rh = q.rho[fi];
      if (rh < 0.0) continue;
      const double amt = q.w[fi] * q.r[(int64_t)(fi / N) * N + js];
      if (rh <= lo) fr_lo += amt; else if (rh <= hi) fr_mg += amt;
    }
    fr_lo = block_sum(fr_lo, red); __syncthreads();
    if (tid == 0) sh[18] = fr_lo;
    __syncthreads();
    fr_mg = block_sum(fr_mg, red); __syncthreads();
    if (tid == 0) sh[19] = fr_mg;
    __syncthreads();
    const double rest = need - sh[18];
    const double mv = sh[19] > 0.0 ? fmin(fmax(rest / sh[19], 0.0), 1.0) : 0.0;   // moved share of marginal flows
    for (int fi = tid; fi < F * N; fi += blockDim.x) {
      const double rh = q.rho[fi];
      if (rh >= 0.0 && (rh <= lo || (rh <= hi && mv >= 1.0))) q.ch[fi] = q.alt[fi];
    }
    __syncthreads();
    if (mv > 0.0 && mv < 1.0 && tid == 0) {
      // the marginal flows (usually one) are handled one after the other by a single thread, in ascending index
      // order: each decision reads the shares its predecessors left behind, so tied thresholds cannot race
      for (int fi = 0; fi < F * N; ++fi) {
        const double rh = q.rho[fi];
        if (rh < 0.0 || rh <= lo || rh > hi) continue;
        // C1b (constraints_step1.py:12-15): the pod (f, js) must keep a total share >= 1 - eps.  If this
        // flow is (nearly) its only customer, splitting it would starve the pod: move the flow completely
        // instead (the emptied pod is closed afterwards) -- feasible, marginally dearer than the LP vertex.
        const int f = fi / N;
        double others = 0.0;
        for (int i2 = 0; i2 < N; ++i2) {
          const int g = f * N + i2;
          if (g == fi) continue;
          if (q.ch[g] == js) others += q.th[g];
          else if (q.sec[g] == js) others += 1.0 - q.th[g];
        }
        bool starve = others + (1.0 - mv) + kEps < 1.0;
        if (starve) {
          // ... unless a source without workload for f can top the pod up for free (its routing costs
          // nothing and loads nothing): some i2 with w[f,i2] == 0 served by a pod that keeps >= 1 without it
          for (int i2 = 0; i2 < N && starve; ++i2) {
            const int g = f * N + i2;
            if (q.w[g] != 0.0 || q.ch[g] == js || q.sec[g] >= 0) continue;
            const int p2 = q.ch[g];
            double inflow = 0.0;
            for (int i3 = 0; i3 < N; ++i3) {
              const int h = f * N + i3;
              if (q.ch[h] == p2) inflow += q.th[h]; else if (q.sec[h] == p2) inflow += 1.0 - q.th[h];
            }
            if (inflow >= 2.0) starve = false;
          }
        }
        if (starve) {
          // most of the flow has to leave anyway: move all of it (the emptied pod is closed afterwards);
          // otherwise keep it whole on js and let the next-cheapest flows make room in the next round
          if (mv > 0.5) q.ch[fi] = q.alt[fi];
          else q.sec[fi] = -2;
        } else { q.sec[fi] = q.alt[fi]; q.th[fi] = 1.0 - mv; }
      }
    }
    if (tid == 0) q.lam[js] = hi;
    __syncthreads();
  }
  __syncthreads();
  if (!s_ok || round >= max_rounds) return res;
  double cost = 0.0;
  for (int fi = tid; fi < F * N; fi += blockDim.x) {
    const int i = fi % N;
    const double wv = q.w[fi];
    if (wv == 0.0) continue;
    const double* di = q.d + (int64_t)i * N;
    double t = q.th[fi] * di[q.ch[fi]];
    if (q.sec[fi] >= 0) t += (1.0 - q.th[fi]) * di[q.sec[fi]];
    cost += wv * t;
  }
  cost = block_sum(cost, red);
  __syncthreads();
  if (tid == 0) sh[16] = cost;
  __syncthreads();
  res.cost = sh[16]; res.feasible = 1; res.rounds = round;
  return res;
}

}  // namespace neptune
