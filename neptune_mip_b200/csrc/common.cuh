// common.cuh -- shared helpers for the neptune_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/neptune_b200.h"

#define NEPTUNE_CUDA_OK(expr)                         \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return (int)_e;            \
  } while (0)

#define NEPTUNE_LAUNCH_OK()                           \
  do {                                                \
    cudaError_t _e = cudaGetLastError();              \
    if (_e != cudaSuccess) return (int)_e;            \
  } while (0)

extern "C" void neptune_count_launches(long long n);   // bench bookkeeping (host-side atomic)
#define NEPTUNE_COUNT(n) neptune_count_launches((long long)(n))

namespace neptune {

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr double kBigM = 1e6;         // reference constraints_step1.py:1
constexpr double kEps = 1e-6;         // reference constraints_step1.py:2

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct Layout {   // closed-form geometry of the canonical model (see include/neptune_b200.h)
  int N, F, kind, flags;
  int64_t X, C;                 // #x columns, #c columns
  int with_n, strengthen;
  int64_t cols, rows, nnz;
  // row offsets of the families
  int64_t r1, r2, r3, r4, r5, r6, rs;
  // nnz offsets of the families
  int64_t e1, e2, e3, e4, e5, e6, es;
  int cw;                       // entries per c column in A^T (3 or 5) (+N with strengthen)
  __host__ __device__ Layout() {}
  __host__ __device__ Layout(int N_, int F_, int kind_, int flags_) {
    N = N_; F = F_; kind = kind_; flags = flags_;
    int64_t n = N, f = F;
    X = f * n * n; C = f * n;
    with_n = kind != NEPTUNE_KIND_MIN_DELAY;
    strengthen = (flags & NEPTUNE_FLAG_STRENGTHEN) ? 1 : 0;
    cols = X + C + (with_n ? n : 0);
    r1 = 0; r2 = 2 * C; r3 = r2 + n; r4 = r3 + C; r5 = r4 + n;
    r6 = r5 + (with_n ? 2 * n : 0);
    rs = r6 + (with_n ? n : 0);
    rows = rs + (strengthen ? X : 0);
    e1 = 0; e2 = 2 * C * (n + 1); e3 = e2 + n * f; e4 = e3 + X; e5 = e4 + X;
    e6 = e5 + (with_n ? 2 * n * (f + 1) : 0);
    es = e6 + (with_n ? n : 0);
    nnz = es + (strengthen ? 2 * X : 0);
    cw = with_n ? 5 : 3;
  }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; result valid in thread 0.  `smem` needs >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_sum(v);
  return v;
}

// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }

}  // namespace neptune
