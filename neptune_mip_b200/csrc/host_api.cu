// host_api.cu -- host-buffer entry points (the e2e path: H2D copies, kernels, D2H copies in one call).
#include "common.cuh"

using namespace neptune;

namespace {
struct DevBuf {          // stream-ordered scratch that frees itself
  cudaStream_t s; void* p = nullptr;
  DevBuf(cudaStream_t s_) : s(s_) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 8, s); }
  ~DevBuf() { if (p) cudaFreeAsync(p, s); }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
}  // namespace

extern "C" int neptune_efttc_host(int B, int N, int F, int kind, double alpha, const double* d_h,
                                  const double* w_h, const double* r_h, const double* m_h, const double* Mj_h,
                                  const double* Kj_h, const double* old_h, const double* maxd_h,
                                  const double* cost_h, double budget, uint8_t* c_out_h, uint8_t* n_out_h,
                                  int32_t* info_out_h, int32_t* flags_out_h, double* scores_out_h,
                                  void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || !d_h || !w_h || !r_h || !m_h || !Mj_h || !Kj_h || !old_h || !maxd_h ||
      !cost_h || !c_out_h || !n_out_h || !info_out_h || !flags_out_h || !scores_out_h)
    return NEPTUNE_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = N, f = F, b = B;
  // one input slab: d | w | r | old | m | maxd | Mj | Kj | cost
  const size_t off_d = 0, off_w = off_d + b * n * n, off_r = off_w + b * f * n, off_old = off_r + b * f * n,
               off_m = off_old + b * f * n, off_maxd = off_m + b * f, off_Mj = off_maxd + b * f,
               off_Kj = off_Mj + b * n, off_cost = off_Kj + b * n, total = off_cost + b * n;
  DevBuf in(s), ws(s), out(s), xbuf(s);
  NEPTUNE_CUDA_OK(in.alloc(total * 8));
  double* D = in.as<double>();
  auto up = [&](size_t off, const double* src, size_t cnt) {
    return cudaMemcpyAsync(D + off, src, cnt * 8, cudaMemcpyHostToDevice, s);
  };
  NEPTUNE_CUDA_OK(up(off_d, d_h, b * n * n));
  NEPTUNE_CUDA_OK(up(off_w, w_h, b * f * n));
  NEPTUNE_CUDA_OK(up(off_r, r_h, b * f * n));
  NEPTUNE_CUDA_OK(up(off_old, old_h, b * f * n));
  NEPTUNE_CUDA_OK(up(off_m, m_h, b * f));
  NEPTUNE_CUDA_OK(up(off_maxd, maxd_h, b * f));
  NEPTUNE_CUDA_OK(up(off_Mj, Mj_h, b * n));
  NEPTUNE_CUDA_OK(up(off_Kj, Kj_h, b * n));
  NEPTUNE_CUDA_OK(up(off_cost, cost_h, b * n));
  int64_t wsb = 0;
  int rc = neptune_efttc_workspace_bytes(B, N, F, &wsb);
  if (rc) return rc;
  NEPTUNE_CUDA_OK(ws.alloc((size_t)wsb));
  // outputs: c | n | info | flags | scores | nd (n as double) | cd (c as double)
  const size_t o_c = 0, o_n = o_c + b * f * n, o_info = (o_n + b * n + 15) & ~(size_t)15,
               o_flags = o_info + b * 16, o_scores = (o_flags + b * 4 + 15) & ~(size_t)15,
               o_nd = o_scores + b * 24, o_cd = o_nd + b * n * 8, o_end = o_cd + b * f * n * 8;
  NEPTUNE_CUDA_OK(out.alloc(o_end));
  char* O = out.as<char>();
  NEPTUNE_CUDA_OK(xbuf.alloc(b * n * f * n * 8));
  rc = neptune_efttc(B, N, F, kind, alpha, D + off_d, D + off_w, D + off_r, D + off_m, D + off_Mj, D + off_Kj,
                     D + off_old, D + off_cost, budget, (uint8_t*)(O + o_c), (uint8_t*)(O + o_n),
                     (int32_t*)(O + o_info), ws.p, wsb, stream);
  if (rc) return rc;
  rc = neptune_route_placements(B, N, F, D + off_d, (uint8_t*)(O + o_c), xbuf.as<double>(),
                                (double*)(O + o_nd), stream);
  if (rc) return rc;
  rc = neptune_u8_to_f64((int64_t)(b * f * n), (const uint8_t*)(O + o_c), (double*)(O + o_cd), stream);
  if (rc) return rc;
  rc = neptune_check_solution(B, N, F, alpha, D + off_d, D + off_w, D + off_r, D + off_m, D + off_Mj,
                              D + off_Kj, D + off_maxd, D + off_cost, budget, xbuf.as<double>(),
                              (double*)(O + o_cd), (double*)(O + o_nd), (int32_t*)(O + o_flags),
                              (double*)(O + o_scores), stream);
  if (rc) return rc;
  NEPTUNE_CUDA_OK(cudaMemcpyAsync(c_out_h, O + o_c, b * f * n, cudaMemcpyDeviceToHost, s));
  NEPTUNE_CUDA_OK(cudaMemcpyAsync(n_out_h, O + o_n, b * n, cudaMemcpyDeviceToHost, s));
  NEPTUNE_CUDA_OK(cudaMemcpyAsync(info_out_h, O + o_info, b * 16, cudaMemcpyDeviceToHost, s));
  NEPTUNE_CUDA_OK(cudaMemcpyAsync(flags_out_h, O + o_flags, b * 4, cudaMemcpyDeviceToHost, s));
  NEPTUNE_CUDA_OK(cudaMemcpyAsync(scores_out_h, O + o_scores, b * 24, cudaMemcpyDeviceToHost, s));
  NEPTUNE_CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

namespace neptune {
__global__ void k_u8_to_f64(int64_t n, const uint8_t* __restrict__ in, double* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    out[k] = in[k] ? 1.0 : 0.0;
}
}  // namespace neptune

extern "C" int neptune_u8_to_f64(int64_t n, const uint8_t* in, double* out, void* stream) {
  if (n <= 0 || !in || !out) return NEPTUNE_E_ARG;
  int64_t g = (n + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  { k_u8_to_f64<<<(int)g, 256, 0, (cudaStream_t)stream>>>(n, in, out); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  return 0;
}

// ---- launch bookkeeping (bench.py reports how many of our kernels ran inside the timed region) ----
#include <atomic>
static std::atomic<long long> g_launches{0};
extern "C" void neptune_count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int neptune_launch_count(int64_t* count, int reset) {
  if (!count) return NEPTUNE_E_ARG;
  *count = g_launches.load(std::memory_order_relaxed);
  if (reset) g_launches.store(0, std::memory_order_relaxed);
  return 0;
}
