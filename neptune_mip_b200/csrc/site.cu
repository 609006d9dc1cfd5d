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
