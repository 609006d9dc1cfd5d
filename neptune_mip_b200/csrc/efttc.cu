// efttc.cu -- (d) the EFTTC ("top trading cycles") greedy on device, one thread block per instance.
//
// Replaces EfttcStepBase.solve() and helpers (reference core/solvers/efttc/efttc_step1.py:39-312,
// score_local :356-439; global checks efttc/utils/constraints_step1.py:70-80,126-133).  The
// reference spends its time copying F*N^2 dict entries per iteration (snapshot_vars) and in an
// F*N^2 Python loop (constrain_CPU_usage); here the routing x is never materialised: per (f, i)
// only the nearest-open-pod delay `mind` and the tie share `inv` are kept, which is all that
// change_x_one (:196-212) encodes.  Control flow, iteration orders, tie-breaks and tolerances are
// the reference's; sums whose rounding can matter (memory, CPU, budget) are
// taken in the reference's order with explicit _rn intrinsics.
//
// `tried_cycles` (:43,64-68) is not stored: every cycle that is found contains an edge f->j with
// (f,j) not in invalid_pairs (:141-143), and every cycle that was ever processed has all of its
// pairs in invalid_pairs afterwards (:77-80, :86-90, :108-109), so a repeat is impossible and the
// `cycle_key in tried_cycles` branch is dead code.
#include "common.cuh"

namespace neptune {

struct EfttcArgs {
  int N, F, kind;
  double alpha, budget;
  const double *d, *w, *r, *m, *Mj, *Kj, *old, *cost;
  uint8_t *c, *n;
  int32_t* info;
  // workspace (per instance)
  double *dT, *wd, *mind, *inv;
  uint8_t* invalid;
  int32_t *act, *nact;
};

struct Sh {            // shared-memory views
  uint8_t *rem_f, *rem_n;
  int *planned, *actual, *out_f, *out_n, *stamp, *path, *cyc_f, *cyc_j, *prev_c;
};

__device__ __forceinline__ double efttc_score(const EfttcArgs& a, const Sh& s, int f, int j) {
  const double warm = (a.old[(int64_t)f * a.N + j] == 1.0) ? 0.5 : 1.0;
  if (a.kind == NEPTUNE_KIND_MIN_DELAY) return __dmul_rn(a.wd[(int64_t)f * a.N + j], warm);
  if (a.kind == NEPTUNE_KIND_MIN_UTIL)
    return __dmul_rn(__ddiv_rn(a.cost[j], (double)(1 + s.planned[j] + s.actual[j])), warm);
  const double q = __ddiv_rn(a.cost[j], (double)(1 + s.planned[j]));
  return __dmul_rn(__dadd_rn(__dmul_rn(a.alpha, q),
                             __dmul_rn(1.0 - a.alpha, a.wd[(int64_t)f * a.N + j])), warm);
}

// lexicographic (score, index) minimum across a warp
__device__ __forceinline__ void warp_argmin(double& s, int& k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const int k2 = __shfl_xor_sync(0xffffffffu, k, o);
    if (k2 >= 0 && (k < 0 || s2 < s || (s2 == s && k2 < k))) { s = s2; k = k2; }
  }
}

// sum_{f2 ascending} (c[f2][j] ? m[f2] : 0), strictly left to right (Python's sum), one warp.
__device__ double warp_mem_used(const EfttcArgs& a, int j) {
  const int lane = threadIdx.x & 31;
  double tot = 0.0;
  for (int base = 0; base < a.F; base += 32) {
    const int f2 = base + lane;
    const double t = (f2 < a.F && a.c[(int64_t)f2 * a.N + j]) ? a.m[f2] : 0.0;
    for (int k = 0; k < 32 && base + k < a.F; ++k) tot = __dadd_rn(tot, __shfl_sync(0xffffffffu, t, k));
  }
  return tot;
}

// routing of function f: per source i, nearest open pod and tie share (change_x_one :196-212)
__device__ void route_function(const EfttcArgs& a, int f) {
  const int na = a.nact[f];
  const int32_t* act = a.act + (int64_t)f * a.N;
  for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
    const double* di = a.d + (int64_t)i * a.N;
    double mn = INFINITY;
    for (int k = 0; k < na; ++k) mn = fmin(mn, di[act[k]]);
    int cnt = 0;
    for (int k = 0; k < na; ++k) cnt += fabs(di[act[k]] - mn) < 1e-6;
    a.mind[(int64_t)f * a.N + i] = mn;
    a.inv[(int64_t)f * a.N + i] = cnt ? __ddiv_rn(1.0, (double)cnt) : 0.0;
  }
}

__device__ void rebuild_active(const EfttcArgs& a, int f) {   // thread 0 only: ascending list
  int k = 0;
  for (int j = 0; j < a.N; ++j) if (a.c[(int64_t)f * a.N + j]) a.act[(int64_t)f * a.N + k++] = j;
  a.nact[f] = k;
}

// find_best_node_by_delay_improvement(...) is not None   (:214-288); whole block cooperates.
__device__ bool has_improving_node(const EfttcArgs& a, const Sh& s, int f, int* flag_sm) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int N = a.N;
  __syncthreads();
  if (threadIdx.x == 0) *flag_sm = 0;
  __syncthreads();
  const double* wf = a.w + (int64_t)f * N;
  const double* cur = a.mind + (int64_t)f * N;
  const bool none_active = a.nact[f] == 0;
  // current score (all warps compute it redundantly; order: lanes strided, then tree)
  double cs = 0.0;
  for (int i = lane; i < N; i += 32) cs += __dmul_rn(wf[i], none_active ? INFINITY : cur[i]);
  cs = warp_sum(cs);
  for (int j = wid; j < N; j += nw) {
    if (!s.rem_n[j] || a.c[(int64_t)f * N + j] || a.invalid[(int64_t)f * N + j]) continue;
    const double* dj = a.dT + (int64_t)j * N;
    double ns = 0.0;
    for (int i = lane; i < N; i += 32)
      ns += __dmul_rn(wf[i], fmin(none_active ? INFINITY : cur[i], dj[i]));
    ns = warp_sum(ns);
    const double delta = cs - ns;
    bool ok;
    if (a.kind == NEPTUNE_KIND_MIN_DELAY) ok = delta > 0.0 + 1e-6;
    else {
      const double du = a.n[j] ? 0.0 : __ddiv_rn(1.0, (double)N);
      ok = __dsub_rn(__dmul_rn(1.0 - a.alpha, delta), __dmul_rn(a.alpha, du)) > 0.0 + 1e-6;
    }
    if (ok && lane == 0) *flag_sm = 1;
  }
  __syncthreads();
  return *flag_sm != 0;
}

__global__ void __launch_bounds__(256) k_efttc(EfttcArgs a0, int64_t ws_stride_bytes, char* ws_base) {
  const int b = blockIdx.x;
  EfttcArgs a = a0;
  const int N = a.N, F = a.F;
  {
    const int64_t n = N, f = F;
    a.d += b * n * n; a.w += b * f * n; a.r += b * f * n; a.m += b * f; a.Mj += b * n; a.Kj += b * n;
    a.old += b * f * n; if (a.cost) a.cost += b * n;
    a.c += b * f * n; a.n += b * n; a.info += b * 4;
    char* p = ws_base + b * ws_stride_bytes;
    a.dT = (double*)p; p += n * n * 8;
    a.wd = (double*)p; p += f * n * 8;
    a.mind = (double*)p; p += f * n * 8;
    a.inv = (double*)p; p += f * n * 8;
    a.act = (int32_t*)p; p += f * n * 4;
    a.nact = (int32_t*)p; p += ((f * 4 + 7) / 8) * 8;
    a.invalid = (uint8_t*)p;
  }
  extern __shared__ int smem_i[];
  Sh s;
  {
    int* p = smem_i;
    s.planned = p; p += N; s.actual = p; p += N; s.out_f = p; p += F; s.out_n = p; p += N;
    s.stamp = p; p += F + N; s.path = p; p += F + N + 1;
    s.cyc_f = p; p += F + N; s.cyc_j = p; p += F + N; s.prev_c = p; p += F + N;
    s.rem_f = (uint8_t*)p; s.rem_n = s.rem_f + F;
  }
  __shared__ int sh_flag, sh_ncyc, sh_stop, sh_ok, sh_nremf, sh_raise;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;

  // ---- init ----------------------------------------------------------------------------------------
  for (int k = tid; k < F * N; k += blockDim.x) { a.c[k] = 0; a.invalid[k] = 0; a.mind[k] = INFINITY; a.inv[k] = 0.0; }
  for (int k = tid; k < N * N; k += blockDim.x) { const int i = k / N, j = k - i * N; a.dT[(int64_t)j * N + i] = a.d[k]; }
  for (int j = tid; j < N; j += blockDim.x) {
    a.n[j] = 0; s.rem_n[j] = 1; s.planned[j] = 0;
    double so = 0.0;
    for (int f = 0; f < F; ++f) so += a.old[(int64_t)f * N + j];
    s.actual[j] = (int)so;
  }
  for (int f = tid; f < F; f += blockDim.x) { s.rem_f[f] = 1; a.nact[f] = 0; }
  __syncthreads();
  if (a.kind != NEPTUNE_KIND_MIN_UTIL) {
    // wd[f][j] = sum_i d[i][j] * w[f][i]   (score_local's dot product, :400,437)
    for (int fj = wid; fj < F * N; fj += nw) {
      const int f = fj / N, j = fj - f * N;
      const double* dj = a.dT + (int64_t)j * N;
      const double* wf = a.w + (int64_t)f * N;
      double acc = 0.0;
      for (int i = lane; i < N; i += 32) acc += dj[i] * wf[i];
      acc = warp_sum(acc);
      if (lane == 0) a.wd[fj] = acc;
    }
  }
  if (tid == 0) { sh_stop = 0; sh_nremf = F; sh_raise = 0; }
  __syncthreads();

  int iterations = 0;
  const bool min_delay_kind = a.kind != NEPTUNE_KIND_MIN_UTIL;   // 'min_delay' in objective (:111)
  while (true) {
    __syncthreads();
    if (sh_nremf == 0 || sh_stop) break;
    ++iterations;
    // ---- preference graph (:123-146) ---------------------------------------------------------------
    for (int f = wid; f < F; f += nw) {
      double bs = INFINITY; int bj = -1;
      if (s.rem_f[f])
        for (int j = lane; j < N; j += 32) {
          if (!s.rem_n[j] || a.invalid[(int64_t)f * N + j]) continue;
          const double sc = efttc_score(a, s, f, j);
          if (bj < 0 || sc < bs) { bs = sc; bj = j; }
        }
      warp_argmin(bs, bj);
      if (lane == 0) s.out_f[f] = bj;
    }
    for (int j = wid; j < N; j += nw) {
      double bs = INFINITY; int bf = -1;
      if (s.rem_n[j])
        for (int f = lane; f < F; f += 32) {
          if (!s.rem_f[f]) continue;
          const double sc = efttc_score(a, s, f, j);
          if (bf < 0 || sc < bs) { bs = sc; bf = f; }
        }
      warp_argmin(bs, bf);
      if (lane == 0) s.out_n[j] = bf;
    }
    for (int k = tid; k < F + N; k += blockDim.x) s.stamp[k] = -1;
    __syncthreads();
    // ---- find_cycle (:148-188): graph vertex v in [0,F) = function, F + j = node ---------------------
    if (tid == 0) {
      int ncyc = 0;
      bool found = false;
      // `visited` only filters the *start* vertices; stamp >= 0 means "in some earlier local_visited"
      for (int start = 0; start < F + N && !found; ++start) {
        const bool in_graph = start < F ? (s.rem_f[start] && s.out_f[start] >= 0)
                                        : (s.rem_n[start - F] != 0);
        if (!in_graph || s.stamp[start] >= 0) continue;
        int len = 0, cur = start;
        while (true) {
          if (s.stamp[cur] == start + 1 + (F + N)) break;      // cannot happen (checked below)
          s.stamp[cur] = start + 1 + (F + N);                   // local mark for this walk
          s.path[len++] = cur;
          int nxt;
          if (cur < F) nxt = (s.rem_f[cur] && s.out_f[cur] >= 0) ? F + s.out_f[cur] : -1;
          else nxt = s.rem_n[cur - F] ? s.out_n[cur - F] : -1;
          if (nxt < 0) break;                                   // dead end: current not in graph
          if (s.stamp[nxt] == start + 1 + (F + N)) {
            int k0 = 0;
            while (s.path[k0] != nxt) ++k0;
            for (int k = k0; k < len; ++k) {
              const int u = s.path[k], v = (k + 1 < len) ? s.path[k + 1] : nxt;
              int pf, pj;
              if (u < F && v >= F) { pf = u; pj = v - F; }
              else if (u >= F && v < F) { pf = v; pj = u - F; }
              else continue;
              bool dup = false;
              for (int q = 0; q < ncyc; ++q) if (s.cyc_f[q] == pf && s.cyc_j[q] == pj) { dup = true; break; }
              if (!dup) { s.cyc_f[ncyc] = pf; s.cyc_j[ncyc] = pj; ++ncyc; }
            }
            found = true;
            break;
          }
          cur = nxt;
        }
        if (!found)   // visited |= local_visited
          for (int k = 0; k < len; ++k) s.stamp[s.path[k]] = 0;
      }
      sh_ncyc = ncyc;
      if (ncyc == 0) sh_stop = 1;                               // no cycle: allocation finished (:61-63)
    }
    __syncthreads();
    const int ncyc = sh_ncyc;
    if (ncyc == 0) break;

    // ---- can_assign_cycle (:290-312), pairs in order ---------------------------------------------------
    if (tid == 0) sh_ok = 0;
    __syncthreads();
    for (int q = 0; q < ncyc; ++q) {
      const int f = s.cyc_f[q], j = s.cyc_j[q];
      if (wid == 0) {
        const double used = warp_mem_used(a, j);
        if (lane == 0) {
          s.prev_c[q] = a.c[(int64_t)f * N + j];
          const bool fits = __dadd_rn(used, a.m[f]) <= a.Mj[j];
          sh_flag = fits;
          if (!fits) { a.invalid[(int64_t)f * N + j] = 1; s.prev_c[q] = -1; }
          else {
            if (!a.c[(int64_t)f * N + j]) {
              a.c[(int64_t)f * N + j] = 1;
              a.act[(int64_t)f * N + a.nact[f]] = j; a.nact[f] += 1;
              s.planned[j] += 1;
            }
            a.n[j] = 1;
            sh_ok = 1;
          }
        }
      }
      __syncthreads();
      if (sh_flag) route_function(a, f);
      __syncthreads();
    }
    if (!sh_ok) continue;                                       // no pair fitted: next iteration (:76-80)

    // ---- global constraints (:328-351): CPU, then budget for the utilisation objectives ---------------
    if (tid == 0) sh_flag = 1;
    __syncthreads();
    for (int j = tid; j < N; j += blockDim.x) {
      if (s.planned[j] == 0) continue;
      double tot = 0.0;
      for (int f = 0; f < F; ++f) {
        if (!a.c[(int64_t)f * N + j]) continue;
        const double rfj = a.r[(int64_t)f * N + j];
        const double* mn = a.mind + (int64_t)f * N;
        const double* iv = a.inv + (int64_t)f * N;
        const double* wf = a.w + (int64_t)f * N;
        for (int i = 0; i < N; ++i) {
          const double dij = a.d[(int64_t)i * N + j];
          if (fabs(dij - mn[i]) < 1e-6)
            tot = __dadd_rn(tot, __dmul_rn(__dmul_rn(iv[i], wf[i]), rfj));
        }
      }
      if (tot > __dadd_rn(a.Kj[j], 1e-6)) sh_flag = 0;
    }
    __syncthreads();
    if (tid == 0 && sh_flag && a.kind != NEPTUNE_KIND_MIN_DELAY) {
      double tot = 0.0;
      for (int j = 0; j < N; ++j) if (a.n[j]) tot = __dadd_rn(tot, a.cost[j]);
      if (tot > __dadd_rn(a.budget, 1e-6)) sh_flag = 0;
    }
    __syncthreads();
    if (!sh_flag) {
      // restore the snapshot (:86-90) and invalidate the whole cycle
      if (tid == 0) {
        for (int q = ncyc - 1; q >= 0; --q) {
          const int f = s.cyc_f[q], j = s.cyc_j[q];
          if (s.prev_c[q] == 0 && a.c[(int64_t)f * N + j]) { a.c[(int64_t)f * N + j] = 0; s.planned[j] -= 1; }
        }
        for (int q = 0; q < ncyc; ++q) {
          const int f = s.cyc_f[q], j = s.cyc_j[q];
          a.invalid[(int64_t)f * N + j] = 1;
          rebuild_active(a, f);
          a.n[j] = s.planned[j] > 0;
        }
      }
      __syncthreads();
      for (int q = 0; q < ncyc; ++q) {
        const int f = s.cyc_f[q];
        if (a.nact[f] == 0) {
          for (int i = tid; i < N; i += blockDim.x) { a.mind[(int64_t)f * N + i] = INFINITY; a.inv[(int64_t)f * N + i] = 0.0; }
        } else route_function(a, f);
        __syncthreads();
      }
      continue;
    }

    // ---- handle_cycle (:92-121): the body runs once per pair -------------------------------------------
    for (int q = 0; q < ncyc; ++q) {
      const int j = s.cyc_j[q];
      if (wid == 0) {
        const double used = warp_mem_used(a, j);
        if (lane == 0) {
          if (used == a.Mj[j]) s.rem_n[j] = 0;
          for (int t = 0; t < ncyc; ++t) a.invalid[(int64_t)s.cyc_f[t] * N + s.cyc_j[t]] = 1;
        }
      }
      __syncthreads();
      for (int t = 0; t < ncyc; ++t) {
        const int f = s.cyc_f[t];
        bool keep = false;
        if (min_delay_kind) keep = has_improving_node(a, s, f, &sh_flag);
        __syncthreads();
        if (tid == 0 && !keep) {
          if (s.rem_f[f]) { s.rem_f[f] = 0; sh_nremf -= 1; }
          else if (min_delay_kind) sh_raise = 1;       // the reference's set.remove raises KeyError (:118)
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int pods = 0;
    for (int k = 0; k < F * N; ++k) pods += a.c[k];
    a.info[0] = iterations; a.info[1] = pods; a.info[2] = sh_raise; a.info[3] = sh_stop;
  }
}

static inline int64_t efttc_ws_per_instance(int N, int F) {
  int64_t n = N, f = F;
  int64_t b = n * n * 8 + 3 * f * n * 8 + f * n * 4 + ((f * 4 + 7) / 8) * 8 + f * n;
  return (b + 255) & ~(int64_t)255;
}

}  // namespace neptune

using namespace neptune;

extern "C" int neptune_efttc_workspace_bytes(int B, int N, int F, int64_t* bytes) {
  if (B <= 0 || N <= 0 || F <= 0 || !bytes) return NEPTUNE_E_ARG;
  *bytes = (int64_t)B * efttc_ws_per_instance(N, F);
  return 0;
}

extern "C" int neptune_efttc(int B, int N, int F, int kind, double alpha, const double* d, const double* w,
                             const double* r, const double* m, const double* Mj, const double* Kj,
                             const double* old, const double* cost, double budget, uint8_t* c_out,
                             uint8_t* n_out, int32_t* info_out, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || kind < 0 || kind > 2) return NEPTUNE_E_ARG;
  if (!d || !w || !r || !m || !Mj || !Kj || !old || !c_out || !n_out || !info_out || !workspace)
    return NEPTUNE_E_ARG;
  if (kind != NEPTUNE_KIND_MIN_DELAY && !cost) return NEPTUNE_E_ARG;
  const int64_t per = efttc_ws_per_instance(N, F);
  if (workspace_bytes < per * B) return NEPTUNE_E_NOMEM;
  EfttcArgs a{};
  a.N = N; a.F = F; a.kind = kind; a.alpha = alpha; a.budget = budget;
  a.d = d; a.w = w; a.r = r; a.m = m; a.Mj = Mj; a.Kj = Kj; a.old = old; a.cost = cost;
  a.c = c_out; a.n = n_out; a.info = info_out;
  const size_t sm = (size_t)(3 * N + F + 5 * (F + N) + 1) * 4 + (size_t)(F + N) + 16;
  if (sm > 200 * 1024) return NEPTUNE_E_SIZE;
  NEPTUNE_CUDA_OK(cudaFuncSetAttribute(k_efttc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  { k_efttc<<<B, 256, sm, (cudaStream_t)stream>>>(a, per, (char*)workspace); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  return 0;
}
