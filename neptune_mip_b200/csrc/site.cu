// site.cu -- (d2) round-robin delay-improvement greedy: a placement for instances too large for the per-instance
// EFTTC block and the shared-memory searches (BASELINE config 4: 2000 nodes x 200 functions).
//
// The move is EFTTC's own: "add the pod of f on the node with the largest delay improvement"
//   gain(f, j) = sum_i w[f,i] * max(0, cur[f,i] - d[i,j]),   cur[f,i] = delay from i to its nearest pod of f
// (reference core/solvers/efttc/efttc_step1.py:214-288, find_best_node_by_delay_improvement; memory test :290-312).
// EFTTC's trading cycles hand the nodes to whichever functions gain most and at 2000 x 200 leave 180 of the 200
// functions without a pod (profiles/r02_c4_placement.json); here EVERY function proposes its best node in every
// round, a node accepts proposals in order of gain while its memory lasts, and the rounds go on until no
// function gains or no memory is left.  One round = four launches over all (instance, function) pairs:
//   k_site_gain    : gain[f][j] for every node (thread per j, sources streamed through shared memory; d is read
//                    coalesced along j, N*N*8 bytes per function from L2)          -- the O(F N^2) part
//   k_site_pick    : per function, the node with the largest gain (lowest index among equals)
//   k_site_resolve : per instance, proposals in order of gain (ties: lowest function), memory accounting
//   k_site_apply   : cur[f][i] = min(cur[f][i], d[i][j]) for the accepted pods
// Deterministic.  CPU capacity is not part of the greedy: the placement goes to neptune_route_capacitated and the
// checkers afterwards (a function has ~ N*slots/F pods, so a pod carries a small share of its function's load).
#include "common.cuh"

namespace neptune {

constexpr int kSiteThreads = 256;
constexpr int kSiteChunk = 512;          // sources staged per shared-memory chunk

__global__ void __launch_bounds__(kSiteThreads)
k_site_gain(int N, int F, const double* __restrict__ d0, const double* __restrict__ w0, const double* __restrict__ m0,
            const double* __restrict__ cur0, const double* __restrict__ memfree0, const uint8_t* __restrict__ c0,
            double* __restrict__ gain0) {
  const int b = blockIdx.z, f = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  const double* __restrict__ d = d0 + (int64_t)b * N * N;
  const double* __restrict__ w = w0 + ((int64_t)b * F + f) * N;
  const double* __restrict__ cur = cur0 + ((int64_t)b * F + f) * N;
  __shared__ double s_w[kSiteChunk], s_c[kSiteChunk];
  double acc = 0.0;
  for (int i0 = 0; i0 < N; i0 += kSiteChunk) {
    const int n = min(kSiteChunk, N - i0);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) { s_w[t] = w[i0 + t]; s_c[t] = cur[i0 + t]; }
    __syncthreads();
    if (j < N) {
      const double* __restrict__ dj = d + (int64_t)i0 * N + j;
#pragma unroll 4
      for (int t = 0; t < n; ++t) {
        const double wv = s_w[t];
        if (wv == 0.0) continue;                              // warp-uniform: no workload, no gain
        acc = __dadd_rn(acc, __dmul_rn(wv, fmax(s_c[t] - dj[(int64_t)t * N], 0.0)));   // no FMA contraction: the oracle sums the same way
      }
    }
  }
  if (j < N) {
    const bool open = c0[((int64_t)b * F + f) * N + j] != 0;
    const bool fits = m0[(int64_t)b * F + f] <= memfree0[(int64_t)b * N + j] + 1e-9;
    gain0[((int64_t)b * F + f) * N + j] = (open || !fits) ? -1.0 : acc;
  }
}

__global__ void __launch_bounds__(kSiteThreads)
k_site_pick(int N, int F, const double* __restrict__ gain0, double* __restrict__ best_g, int* __restrict__ best_j) {
  const int b = blockIdx.y, f = blockIdx.x;
  const double* __restrict__ g = gain0 + ((int64_t)b * F + f) * N;
  double bv = -1.0; int bj = -1;
  for (int j = threadIdx.x; j < N; j += blockDim.x) { const double v = g[j]; if (v > bv) { bv = v; bj = j; } }   // ascending j per thread
  __shared__ double s_v[kSiteThreads];
  __shared__ int s_j[kSiteThreads];
  s_v[threadIdx.x] = bv; s_j[threadIdx.x] = bj;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double v2 = s_v[threadIdx.x + o]; const int j2 = s_j[threadIdx.x + o];
      if (v2 > s_v[threadIdx.x] || (v2 == s_v[threadIdx.x] && j2 >= 0 && (s_j[threadIdx.x] < 0 || j2 < s_j[threadIdx.x]))) {
        s_v[threadIdx.x] = v2; s_j[threadIdx.x] = j2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { best_g[(int64_t)b * F + f] = s_v[0]; best_j[(int64_t)b * F + f] = s_j[0]; }
}

// one block per instance; thread 0 walks the proposals in order of gain (F is a few hundred)
__global__ void k_site_resolve(int N, int F, const double* __restrict__ m0, double* __restrict__ memfree0,
                               double* __restrict__ best_g, const int* __restrict__ best_j, int* __restrict__ accepted,
                               uint8_t* __restrict__ c0, int* __restrict__ npods0,
                               int* __restrict__ info /* [B][2]: rounds with a pod, pods */, int* __restrict__ placed_total) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  double* g = best_g + (int64_t)b * F;
  const int* bj = best_j + (int64_t)b * F;
  int* acc = accepted + (int64_t)b * F;
  double* memfree = memfree0 + (int64_t)b * N;
  int* npods = npods0 + (int64_t)b * F;
  // a function without workload gains nothing anywhere, but the model wants every source routed (C3): while it has
  // no pod its best node counts as a (smallest possible) gain
  for (int f = 0; f < F; ++f) { acc[f] = -1; if (npods[f] == 0 && g[f] == 0.0) g[f] = 1e-300; }
  int placed = 0;
  for (int k = 0; k < F; ++k) {
    int bf = -1; double bv = 0.0;                                // strictly positive gain only
    for (int f = 0; f < F; ++f) if (g[f] > bv) { bv = g[f]; bf = f; }
    if (bf < 0) break;
    g[bf] = -1.0;
    const int j = bj[bf];
    const double mf = m0[(int64_t)b * F + bf];
    if (j < 0 || mf > memfree[j] + 1e-9) continue;               // the node filled up in this round: next round
    memfree[j] -= mf;
    c0[((int64_t)b * F + bf) * N + j] = 1;
    acc[bf] = j;
    npods[bf] += 1;
    ++placed;
  }
  if (placed) { info[b * 2] += 1; info[b * 2 + 1] += placed; atomicAdd(placed_total, placed); }
}

__global__ void __launch_bounds__(kSiteThreads)
k_site_apply(int N, int F, const double* __restrict__ d0, const int* __restrict__ accepted, double* __restrict__ cur0) {
  const int b = blockIdx.z, f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = accepted[(int64_t)b * F + f];
  if (j < 0 || i >= N) return;
  double* cur = cur0 + ((int64_t)b * F + f) * N;
  cur[i] = fmin(cur[i], d0[(int64_t)b * N * N + (int64_t)i * N + j]);
}

__global__ void k_site_init(int64_t n_cur, double* __restrict__ cur, double big, int64_t n_mem, const double* __restrict__ Mj,
                            double* __restrict__ memfree) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = t; k < n_cur; k += nt) cur[k] = big;
  for (int64_t k = t; k < n_mem; k += nt) memfree[k] = Mj[k];
}

// ---------------------------------------------------------------------------------------------------------------
// Two-choice routing of a placement at sizes where the per-instance routers (one block, serial marginal-flow loop)
// do not finish (C4: 400 000 sources): every source (f, i) sends the share theta[j1] to its nearest pod j1 and the rest
// to its second-nearest pod j2; theta[j] <= 1 is lowered until node j's CPU row holds:
//     P[j] theta[j] + Pfix[j] + S[j] <= K[j],   P = splittable primary load, Pfix = primaries without a second pod,
//     S[j] = spill that arrives from other nodes' sources = sum w r[f,j] (1 - theta[j1]).
// theta starts at 1 and only decreases (S only grows), so the iteration converges to the largest feasible theta.  Loads
// are accumulated in fixed point (integer adds commute): deterministic.  A pod that is nobody's nearest is closed
// first.  Rows of the model it answers: constrain_handle_all_requests + constrain_CPU_usage
// (constraints_step1.py:27-33, 57-65) with the objective of objectives.py:4-11 evaluated on the x it writes.
// ---------------------------------------------------------------------------------------------------------------
constexpr double kTcFx = 1073741824.0;           // 2^30

__global__ void __launch_bounds__(kSiteThreads)
k_tc_nearest(int N, int F, const double* __restrict__ d0, const uint8_t* __restrict__ c0, int* __restrict__ j1o,
             int* __restrict__ j2o, int* __restrict__ prim) {
  const int b = blockIdx.z, f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  extern __shared__ int s_list[];                // pods of f (any order: the choice below breaks ties by node index)
  __shared__ int s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const uint8_t* __restrict__ c = c0 + ((int64_t)b * F + f) * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) if (c[j]) s_list[atomicAdd(&s_n, 1)] = j;
  __syncthreads();
  if (i >= N) return;
  const double* __restrict__ di = d0 + (int64_t)b * N * N + (int64_t)i * N;
  double v1 = INFINITY, v2 = INFINITY; int a1 = -1, a2 = -1;
  const int n = s_n;
  for (int q = 0; q < n; ++q) {
    const int j = s_list[q];
    const double v = di[j];
    if (v < v1 || (v == v1 && j < a1)) { v2 = v1; a2 = a1; v1 = v; a1 = j; }
    else if (v < v2 || (v == v2 && j < a2)) { v2 = v; a2 = j; }
  }
  const int64_t o = ((int64_t)b * F + f) * N + i;
  j1o[o] = a1; j2o[o] = a2;
  if (a1 >= 0) atomicAdd(&prim[((int64_t)b * F + f) * N + a1], 1);
}

__global__ void k_tc_close(int64_t n, uint8_t* __restrict__ c, const int* __restrict__ prim) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && c[t] && prim[t] == 0) c[t] = 0;
}

// which = 0: P and Pfix (once); which = 1: the spill S at the current theta
__global__ void __launch_bounds__(kSiteThreads)
k_tc_loads(int N, int F, const double* __restrict__ w0, const double* __restrict__ r0, const int* __restrict__ j1o,
           const int* __restrict__ j2o, const double* __restrict__ theta0, unsigned long long* __restrict__ P,
           unsigned long long* __restrict__ Pfix, unsigned long long* __restrict__ S, int which) {
  const int b = blockIdx.z, f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t o = ((int64_t)b * F + f) * N + i;
  const double wv = w0[o];
  const int a1 = j1o[o], a2 = j2o[o];
  if (wv <= 0.0 || a1 < 0) return;
  const double* __restrict__ r = r0 + ((int64_t)b * F + f) * N;
  if (which == 0) {
    const unsigned long long q = (unsigned long long)(fmin(wv * r[a1], 1e9) * kTcFx + 0.5);
    atomicAdd((a2 >= 0 ? P : Pfix) + (int64_t)b * N + a1, q);
  } else if (a2 >= 0) {
    const double sp = wv * r[a2] * (1.0 - theta0[(int64_t)b * N + a1]);
    if (sp > 0.0) atomicAdd(S + (int64_t)b * N + a2, (unsigned long long)(fmin(sp, 1e9) * kTcFx + 0.5));
  }
}

__global__ void k_tc_theta(int64_t n, int N, const double* __restrict__ K, const unsigned long long* __restrict__ P,
                           const unsigned long long* __restrict__ Pfix, const unsigned long long* __restrict__ S,
                           double* __restrict__ theta, int* __restrict__ changed, int* __restrict__ over) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double p = (double)P[t] / kTcFx, pf = (double)Pfix[t] / kTcFx, sp = (double)S[t] / kTcFx;
  const double cap = K[t] * (1.0 - 1e-9);
  const double th = theta[t];
  double want = p > 0.0 ? (cap - sp - pf) / p : 1.0;
  want = fmin(fmax(want, 0.0), th);                            // only ever lowered
  if (want < th - 1e-12) { theta[t] = want; atomicAdd(changed, 1); }
  if (p * want + pf + sp > K[t] * (1.0 + 1e-9) + 1e-9) atomicAdd(over + t / N, 1);      // still above capacity at this theta (per instance)
}

__global__ void __launch_bounds__(kSiteThreads)
k_tc_write(int N, int F, const double* __restrict__ d0, const double* __restrict__ w0, const int* __restrict__ j1o,
           const int* __restrict__ j2o, const double* __restrict__ theta0, double* __restrict__ x0,
           double* __restrict__ part) {
  const int b = blockIdx.z, f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double red[32];
  double cost = 0.0;
  if (i < N) {
    const int64_t o = ((int64_t)b * F + f) * N + i;
    const int a1 = j1o[o], a2 = j2o[o];
    if (a1 >= 0) {
      const double th = a2 >= 0 ? theta0[(int64_t)b * N + a1] : 1.0;
      const double* __restrict__ di = d0 + (int64_t)b * N * N + (int64_t)i * N;
      if (x0) {
        double* __restrict__ xr = x0 + (((int64_t)b * N + i) * F + f) * N;
        xr[a1] = th;
        if (a2 >= 0 && th < 1.0) xr[a2] = 1.0 - th;
      }
      cost = w0[o] * (th * di[a1] + (a2 >= 0 ? (1.0 - th) * di[a2] : 0.0));
    }
  }
  const double s = block_sum(cost, red);
  if (threadIdx.x == 0) part[((int64_t)b * F + f) * gridDim.x + blockIdx.x] = s;
}

__global__ void k_tc_finish(int N, int F, int nblk, const double* __restrict__ part, const uint8_t* __restrict__ c0,
                            const int* __restrict__ over, double* __restrict__ n_out, double* __restrict__ obj,
                            int32_t* __restrict__ feas) {
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    int any = 0;
    for (int f = 0; f < F; ++f) any |= c0[((int64_t)b * F + f) * N + j];
    n_out[(int64_t)b * N + j] = any ? 1.0 : 0.0;
  }
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int64_t k = 0; k < (int64_t)F * nblk; ++k) s += part[(int64_t)b * F * nblk + k];      // fixed order
    obj[b] = s;
    feas[b] = over[b] == 0 ? 1 : 0;
  }
}

static inline int64_t site_align(int64_t v) { return (v + 255) & ~(int64_t)255; }

}  // namespace neptune

using namespace neptune;

extern "C" int neptune_site_greedy_workspace_bytes(int B, int N, int F, int64_t* bytes) {
  if (B <= 0 || N <= 0 || F <= 0 || !bytes) return NEPTUNE_E_ARG;
  const int64_t fn = (int64_t)B * F * N;
  *bytes = 2 * site_align(fn * 8) + site_align((int64_t)B * N * 8) + site_align((int64_t)B * F * 8) +
           3 * site_align((int64_t)B * F * 4) + 512;
  return 0;
}

extern "C" int neptune_site_greedy(int B, int N, int F, const double* d, const double* w, const double* m,
                                   const double* Mj, double unserved_delay, int max_rounds, uint8_t* c_out,
                                   int32_t* info_out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || !d || !w || !m || !Mj || !c_out || !info_out || !workspace || max_rounds <= 0 ||
      !(unserved_delay > 0.0))
    return NEPTUNE_E_ARG;
  int64_t need = 0;
  neptune_site_greedy_workspace_bytes(B, N, F, &need);
  if (workspace_bytes < need) return NEPTUNE_E_NOMEM;
  if (F > 65535 || B > 65535) return NEPTUNE_E_SIZE;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t fn = (int64_t)B * F * N;
  char* p = (char*)workspace;
  double* cur = (double*)p; p += site_align(fn * 8);
  double* gain = (double*)p; p += site_align(fn * 8);
  double* memfree = (double*)p; p += site_align((int64_t)B * N * 8);
  double* best_g = (double*)p; p += site_align((int64_t)B * F * 8);
  int* best_j = (int*)p; p += site_align((int64_t)B * F * 4);
  int* accepted = (int*)p; p += site_align((int64_t)B * F * 4);
  int* npods = (int*)p; p += site_align((int64_t)B * F * 4);
  int* d_placed = (int*)p;
  NEPTUNE_CUDA_OK(cudaMemsetAsync(c_out, 0, (size_t)fn, s));
  NEPTUNE_CUDA_OK(cudaMemsetAsync(info_out, 0, (size_t)B * 2 * sizeof(int32_t), s));
  NEPTUNE_CUDA_OK(cudaMemsetAsync(npods, 0, (size_t)B * F * sizeof(int), s));
  NEPTUNE_CUDA_OK(cudaMemsetAsync(d_placed, 0, 4, s));
  { k_site_init<<<4 * kNumSMs, 256, 0, s>>>(fn, cur, unserved_delay, (int64_t)B * N, Mj, memfree); NEPTUNE_COUNT(1); }
  const dim3 gfn((N + kSiteThreads - 1) / kSiteThreads, F, B);
  // the host looks at the number of pods placed every `look` rounds (one 4-byte copy): the greedy ends when a round
  // places nothing (no gain anywhere, or no memory left)
  const int look = 4;
  int h_prev = -1;
  for (int rnd = 0; rnd < max_rounds; ++rnd) {
    { k_site_gain<<<gfn, kSiteThreads, 0, s>>>(N, F, d, w, m, cur, memfree, c_out, gain); NEPTUNE_COUNT(1); }
    { k_site_pick<<<dim3(F, B), kSiteThreads, 0, s>>>(N, F, gain, best_g, best_j); NEPTUNE_COUNT(1); }
    { k_site_resolve<<<B, 32, 0, s>>>(N, F, m, memfree, best_g, best_j, accepted, c_out, npods, info_out, d_placed); NEPTUNE_COUNT(1); }
    { k_site_apply<<<gfn, kSiteThreads, 0, s>>>(N, F, d, accepted, cur); NEPTUNE_COUNT(1); }
    if ((rnd + 1) % look == 0 || rnd + 1 == max_rounds) {
      int h = 0;
      NEPTUNE_CUDA_OK(cudaMemcpyAsync(&h, d_placed, 4, cudaMemcpyDeviceToHost, s));
      NEPTUNE_CUDA_OK(cudaStreamSynchronize(s));
      if (h == h_prev) break;                                   // nothing placed in the last `look` rounds
      h_prev = h;
    }
  }
  NEPTUNE_LAUNCH_OK();
  return 0;
}

extern "C" int neptune_route_two_choice_workspace_bytes(int B, int N, int F, int64_t* bytes) {
  if (B <= 0 || N <= 0 || F <= 0 || !bytes) return NEPTUNE_E_ARG;
  const int64_t fn = (int64_t)B * F * N, nblk = (N + kSiteThreads - 1) / kSiteThreads;
  *bytes = 3 * site_align(fn * 4) + 4 * site_align((int64_t)B * N * 8) + site_align((int64_t)B * F * nblk * 8) + site_align(4 * ((int64_t)B + 1)) + 256;
  return 0;
}

extern "C" int neptune_route_two_choice(int B, int N, int F, const double* d, const double* w, const double* r,
                                        const double* Kj, const uint8_t* c, uint8_t* c_out, double* x, double* n_out,
                                        double* obj_out, int32_t* feas_out, int32_t* iters_out, int max_iters,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || !d || !w || !r || !Kj || !c || !c_out || !n_out || !obj_out || !feas_out ||
      !workspace || max_iters <= 0)
    return NEPTUNE_E_ARG;
  int64_t need = 0;
  neptune_route_two_choice_workspace_bytes(B, N, F, &need);
  if (workspace_bytes < need) return NEPTUNE_E_NOMEM;
  if (F > 65535 || B > 65535) return NEPTUNE_E_SIZE;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t fn = (int64_t)B * F * N, bn = (int64_t)B * N;
  const int nblk = (N + kSiteThreads - 1) / kSiteThreads;
  char* p = (char*)workspace;
  int* j1 = (int*)p; p += site_align(fn * 4);
  int* j2 = (int*)p; p += site_align(fn * 4);
  int* prim = (int*)p; p += site_align(fn * 4);
  double* theta = (double*)p; p += site_align(bn * 8);
  unsigned long long* P = (unsigned long long*)p; p += site_align(bn * 8);
  unsigned long long* Pfix = (unsigned long long*)p; p += site_align(bn * 8);
  unsigned long long* S = (unsigned long long*)p; p += site_align(bn * 8);
  double* part = (double*)p; p += site_align((int64_t)B * F * nblk * 8);
  int* d_flags = (int*)p;                        // [0] changed, [1 + b] nodes of instance b still above capacity
  const dim3 gfn(nblk, F, B);
  const size_t sm = (size_t)N * sizeof(int);
  if (sm > 200 * 1024) return NEPTUNE_E_SIZE;
  NEPTUNE_CUDA_OK(cudaFuncSetAttribute(k_tc_nearest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  NEPTUNE_CUDA_OK(cudaMemcpyAsync(c_out, c, (size_t)fn, cudaMemcpyDeviceToDevice, s));
  // nearest and second-nearest pod of every source; pods that are nobody's nearest are closed, then once more
  for (int pass = 0; pass < 2; ++pass) {
    NEPTUNE_CUDA_OK(cudaMemsetAsync(prim, 0, (size_t)fn * 4, s));
    { k_tc_nearest<<<gfn, kSiteThreads, sm, s>>>(N, F, d, c_out, j1, j2, prim); NEPTUNE_COUNT(1); }
    if (pass == 0) { k_tc_close<<<(int)((fn + 255) / 256), 256, 0, s>>>(fn, c_out, prim); NEPTUNE_COUNT(1); }
  }
  NEPTUNE_CUDA_OK(cudaMemsetAsync(P, 0, (size_t)bn * 8, s));
  NEPTUNE_CUDA_OK(cudaMemsetAsync(Pfix, 0, (size_t)bn * 8, s));
  { k_site_init<<<4 * kNumSMs, 256, 0, s>>>(bn, theta, 1.0, 0, nullptr, nullptr); NEPTUNE_COUNT(1); }
  { k_tc_loads<<<gfn, kSiteThreads, 0, s>>>(N, F, w, r, j1, j2, theta, P, Pfix, S, 0); NEPTUNE_COUNT(1); }
  int it = 0, h[1] = {1};
  for (; it < max_iters && h[0]; ++it) {
    NEPTUNE_CUDA_OK(cudaMemsetAsync(S, 0, (size_t)bn * 8, s));
    NEPTUNE_CUDA_OK(cudaMemsetAsync(d_flags, 0, 4 * ((size_t)B + 1), s));
    { k_tc_loads<<<gfn, kSiteThreads, 0, s>>>(N, F, w, r, j1, j2, theta, P, Pfix, S, 1); NEPTUNE_COUNT(1); }
    { k_tc_theta<<<(int)((bn + 255) / 256), 256, 0, s>>>(bn, N, Kj, P, Pfix, S, theta, d_flags, d_flags + 1); NEPTUNE_COUNT(1); }
    NEPTUNE_CUDA_OK(cudaMemcpyAsync(h, d_flags, 4, cudaMemcpyDeviceToHost, s));
    NEPTUNE_CUDA_OK(cudaStreamSynchronize(s));
  }
  if (iters_out) *iters_out = it;
  if (x) NEPTUNE_CUDA_OK(cudaMemsetAsync(x, 0, (size_t)B * N * F * N * 8, s));
  { k_tc_write<<<gfn, kSiteThreads, 0, s>>>(N, F, d, w, j1, j2, theta, x, part); NEPTUNE_COUNT(1); }
  { k_tc_finish<<<B, 256, 0, s>>>(N, F, nblk, part, c_out, d_flags + 1, n_out, obj_out, feas_out); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  return 0;
}
