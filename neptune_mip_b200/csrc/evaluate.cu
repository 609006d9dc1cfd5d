// evaluate.cu -- (c1) exact feasibility + objective evaluation.
//
//  * neptune_check_solution : the reference's checkers (efttc/utils/constraints_step1.py:5-133) and
//    scorers (efttc/utils/objectives.py:23-98) on dense x/c/n.  Every checker sum runs in the
//    reference's order inside ONE thread (lanes differ in j, so global loads stay coalesced), with
//    explicit __dadd_rn/__dmul_rn so no FMA contraction changes a rounding: verdicts are bit-exact.
//  * neptune_route_placements : closed-form routing of change_x_one (efttc_step1.py:196-212).
//  * neptune_eval_placements  : P candidate placements per instance, never materialising x.
#include "common.cuh"
#include "route_cap.cuh"

namespace neptune {

struct Inst {
  int N, F;
  const double *d, *w, *r, *m, *Mj, *Kj, *maxd, *cost;
  double budget, alpha;
};

__device__ __forceinline__ Inst inst_at(const Inst& in, int b) {
  Inst o = in;
  const int64_t n = in.N, f = in.F;
  o.d = in.d + b * n * n; o.w = in.w + b * f * n; o.r = in.r + b * f * n; o.m = in.m + b * f;
  o.Mj = in.Mj + b * n; o.Kj = in.Kj + b * n;
  o.maxd = in.maxd ? in.maxd + b * f : nullptr;
  o.cost = in.cost ? in.cost + b * n : nullptr;
  return o;
}

// scratch per instance: [0] sum(w)  [1] wmax (masked with 0, objectives.py:73-76)
// [2] sum x*d*w  [3] sum float32(x)*w*d  [4] #active nodes
constexpr int kScr = 8;

__global__ void __launch_bounds__(256) k_check_pre(Inst in0, double* __restrict__ scratch,
                                                   int32_t* __restrict__ flags) {
  const int b = blockIdx.x;
  const Inst in = inst_at(in0, b);
  __shared__ double sm[32];
  double tw = 0.0, wm = 0.0;
  for (int fi = threadIdx.x; fi < in.F * in.N; fi += blockDim.x) {
    const int f = fi / in.N, i = fi - f * in.N;
    const double md = in.maxd ? in.maxd[f] : INFINITY;
    double best = 0.0;                     // np.where(mask, d, 0).max()
    bool first = true;
    for (int j = 0; j < in.N; ++j) {
      const double dij = in.d[(int64_t)i * in.N + j];
      const double v = (dij <= md) ? dij : 0.0;
      if (first || v > best) { best = v; first = false; }
    }
    tw += in.w[fi];
    wm += in.w[fi] * best;
  }
  tw = block_sum(tw, sm);
  wm = block_sum(wm, sm);
  if (threadIdx.x == 0) {
    double* s = scratch + (int64_t)b * kScr;
    s[0] = tw; s[1] = wm; s[2] = 0.0; s[3] = 0.0; s[4] = 0.0;
    flags[b] = NEPTUNE_OK_ALL;
  }
}

__global__ void __launch_bounds__(256) k_check(Inst in0, const double* __restrict__ x0,
                                               const double* __restrict__ c0, const double* __restrict__ n0,
                                               double* __restrict__ scratch, int32_t* __restrict__ flags) {
  const int b = blockIdx.y;
  const Inst in = inst_at(in0, b);
  const int N = in.N, F = in.F;
  const double* __restrict__ x = x0 + (int64_t)b * N * F * N;
  const double* __restrict__ c = c0 + (int64_t)b * F * N;
  const double* __restrict__ nn = n0 + (int64_t)b * N;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  int bad = 0;
  __shared__ double sm[32];

  // c <-> x : per (f, j), sum over sources i ascending          constraints_step1.py:5-18
  for (int64_t fj = tid; fj < (int64_t)F * N; fj += nth) {
    const int f = (int)(fj / N), j = (int)(fj - (int64_t)f * N);
    double s = 0.0;
    for (int i = 0; i < N; ++i) s = __dadd_rn(s, x[((int64_t)i * F + f) * N + j]);
    const bool on = c[fj] > 0.0;
    if (s > (on ? kBigM : 0.0)) bad |= NEPTUNE_OK_C_X;
    if (__dadd_rn(s, kEps) < (on ? 1.0 : 0.0)) bad |= NEPTUNE_OK_C_X;
  }
  // handle_all_requests : per (f, i), sum over j ascending, |total - 1| < 0.1        :37-47
  for (int64_t fi = tid; fi < (int64_t)F * N; fi += nth) {
    const int f = (int)(fi / N), i = (int)(fi - (int64_t)f * N);
    const double* row = x + ((int64_t)i * F + f) * N;
    double s = 0.0;
    for (int j = 0; j < N; ++j) s = __dadd_rn(s, row[j]);
    if (!(fabs(__dsub_rn(s, 1.0)) < 0.1)) bad |= NEPTUNE_OK_HANDLE;
  }
  // per node j: memory (:22-33), CPU (:70-80, f outer / i inner, (x*w)*r), n <-> c (:85-95)
  for (int64_t j = tid; j < N; j += nth) {
    double mem = 0.0;
    int cnt = 0;
    for (int f = 0; f < F; ++f)
      if (c[(int64_t)f * N + j] > 0.0) { mem = __dadd_rn(mem, in.m[f]); ++cnt; }
    if (mem > in.Mj[j]) bad |= NEPTUNE_OK_MEMORY;
    double tot = 0.0;
    for (int f = 0; f < F; ++f) {
      const double rfj = in.r[(int64_t)f * N + j];
      for (int i = 0; i < N; ++i) {
        const double v = x[((int64_t)i * F + f) * N + j];
        tot = __dadd_rn(tot, __dmul_rn(__dmul_rn(v, in.w[(int64_t)f * N + i]), rfj));
      }
    }
    if (tot > __dadd_rn(in.Kj[j], 1e-6)) bad |= NEPTUNE_OK_CPU;
    const int nv = nn[j] > 0.0 ? 1 : 0;
    if ((double)cnt > (double)nv * kBigM) bad |= NEPTUNE_OK_N_C;
    if (__dadd_rn((double)cnt, kEps) < (double)nv) bad |= NEPTUNE_OK_N_C;
  }
  // budget : one thread, sequential over j                                             :126-133
  if (tid == 0 && in.cost) {
    double tot = 0.0;
    int act = 0;
    for (int j = 0; j < N; ++j) {
      const double nv = nn[j] > 0.0 ? 1.0 : 0.0;
      tot = __dadd_rn(tot, __dmul_rn(nv, in.cost[j]));
      act += nn[j] > 0.0;
    }
    if (tot > __dadd_rn(in.budget, 1e-6)) bad |= NEPTUNE_OK_BUDGET;
    scratch[(int64_t)b * kScr + 4] = (double)act;
  }
  // scores: sum x*d*w in fp64 and with x rounded through float32 (objectives.py:86)
  double s64 = 0.0, s32 = 0.0;
  for (int64_t k = tid; k < (int64_t)N * F * N; k += nth) {
    const int j = (int)(k % N);
    const int64_t t = k / N;
    const int f = (int)(t % F), i = (int)(t / F);
    const double v = x[k];
    if (v != 0.0) {
      const double dw_a = in.d[(int64_t)i * N + j], wv = in.w[(int64_t)f * N + i];
      s64 += v * dw_a * wv;
      s32 += (double)(float)v * wv * dw_a;
    }
  }
  s64 = block_sum(s64, sm);
  s32 = block_sum(s32, sm);
  if (threadIdx.x == 0) {
    if (s64 != 0.0) atomicAdd(scratch + (int64_t)b * kScr + 2, s64);
    if (s32 != 0.0) atomicAdd(scratch + (int64_t)b * kScr + 3, s32);
  }
  if (bad) atomicAnd(flags + b, ~bad);
}

__global__ void k_check_fin(int B, int N, double alpha, const double* __restrict__ scratch,
                            double* __restrict__ scores) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* s = scratch + (int64_t)b * kScr;
  const double util = s[4];
  double comb = util * (alpha / (double)N);
  if (s[0] != 0.0 && s[1] != 0.0) comb = comb + s[3] * (1.0 - alpha) / s[1];
  scores[(int64_t)b * 3 + 0] = s[2];
  scores[(int64_t)b * 3 + 1] = util;
  scores[(int64_t)b * 3 + 2] = comb;
}

// ---- routing ---------------------------------------------------------------------------------------
// block per (function, instance); warps stride over sources i, lanes over destinations j.
__global__ void __launch_bounds__(256) k_route(int N, int F, const double* __restrict__ d0,
                                               const uint8_t* __restrict__ c0, double* __restrict__ x0,
                                               double* __restrict__ n0) {
  const int f = blockIdx.x, b = blockIdx.y;
  const double* __restrict__ d = d0 + (int64_t)b * N * N;
  const uint8_t* __restrict__ c = c0 + ((int64_t)b * F + f) * N;
  double* __restrict__ x = x0 + (int64_t)b * N * F * N;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wid; i < N; i += nw) {
    const double* di = d + (int64_t)i * N;
    double mn = INFINITY;
    for (int j = lane; j < N; j += 32) if (c[j]) mn = fmin(mn, di[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    int cnt = 0;
    for (int j = lane; j < N; j += 32) cnt += (c[j] && fabs(di[j] - mn) < 1e-6);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const double share = cnt ? 1.0 / (double)cnt : 0.0;
    double* xr = x + ((int64_t)i * F + f) * N;
    for (int j = lane; j < N; j += 32) xr[j] = (c[j] && fabs(di[j] - mn) < 1e-6) ? share : 0.0;
  }
  if (f == 0 && n0) {
    const uint8_t* cb = c0 + (int64_t)b * F * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      int any = 0;
      for (int g = 0; g < F; ++g) any |= cb[(int64_t)g * N + j];
      n0[(int64_t)b * N + j] = any ? 1.0 : 0.0;
    }
  }
}

// ---- batched candidate evaluation -------------------------------------------------------------------
// block per (candidate, instance).  Functions are processed one after the other; threads own sources.
// dynamic smem: load[N] + recv[N] doubles, act[N] ints
__global__ void __launch_bounds__(256) k_eval(Inst in0, int P, const uint8_t* __restrict__ c0,
                                              const double* __restrict__ pre,   // [B][kScr] from k_check_pre
                                              double* __restrict__ obj_out, int32_t* __restrict__ flags_out,
                                              double* __restrict__ over_out) {
  const int p = blockIdx.x, b = blockIdx.y;
  const Inst in = inst_at(in0, b);
  const int N = in.N, F = in.F;
  const uint8_t* __restrict__ c = c0 + ((int64_t)b * P + p) * F * N;
  extern __shared__ double smem[];
  double* load = smem;
  double* recv = smem + N;
  int* act = reinterpret_cast<int*>(smem + 2 * N);
  __shared__ double red[32];
  __shared__ int n_act, bad_s;
  if (threadIdx.x == 0) bad_s = 0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) load[j] = 0.0;
  double delay = 0.0;
  int bad = 0;
  for (int f = 0; f < F; ++f) {
    __syncthreads();
    if (threadIdx.x == 0) n_act = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) recv[j] = 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x)
      if (c[(int64_t)f * N + j]) act[atomicAdd(&n_act, 1)] = j;
    __syncthreads();
    const int na = n_act;
    if (na == 0) { bad |= NEPTUNE_OK_HANDLE; continue; }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double* di = in.d + (int64_t)i * N;
      double mn = INFINITY;
      for (int k = 0; k < na; ++k) mn = fmin(mn, di[act[k]]);
      int cnt = 0;
      for (int k = 0; k < na; ++k) cnt += fabs(di[act[k]] - mn) < 1e-6;
      const double share = 1.0 / (double)cnt;
      const double wfi = in.w[(int64_t)f * N + i];
      for (int k = 0; k < na; ++k) {
        const int j = act[k];
        const double dij = di[j];
        if (fabs(dij - mn) < 1e-6) {
          atomicAdd(recv + j, share);
          if (wfi != 0.0) {
            atomicAdd(load + j, share * wfi * in.r[(int64_t)f * N + j]);
            delay += share * dij * wfi;
          }
        }
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < na; k += blockDim.x)
      if (recv[act[k]] + kEps < 1.0) bad |= NEPTUNE_OK_C_X;
  }
  __syncthreads();
  // per-node checks
  double over = 0.0, util = 0.0, costsum = 0.0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    double mem = 0.0;
    int cnt = 0;
    for (int f = 0; f < F; ++f) if (c[(int64_t)f * N + j]) { mem += in.m[f]; ++cnt; }
    if (mem > in.Mj[j]) bad |= NEPTUNE_OK_MEMORY;
    const double ex = load[j] - (in.Kj[j] + 1e-6);
    if (ex > 0.0) { bad |= NEPTUNE_OK_CPU; over += load[j] - in.Kj[j]; }
    if (cnt) { util += 1.0; costsum += in.cost ? in.cost[j] : 0.0; }
  }
  delay = block_sum(delay, red);
  over = block_sum(over, red);
  util = block_sum(util, red);
  costsum = block_sum(costsum, red);
  if (bad) atomicOr(&bad_s, bad);
  __syncthreads();
  if (threadIdx.x == 0) {
    int fl = NEPTUNE_OK_ALL & ~bad_s;
    if (costsum > in.budget + 1e-6) fl &= ~NEPTUNE_OK_BUDGET;
    const double* s = pre + (int64_t)b * kScr;
    double comb = util * (in.alpha / (double)N);
    if (s[0] != 0.0 && s[1] != 0.0) comb += delay * (1.0 - in.alpha) / s[1];
    const int64_t o = (int64_t)b * P + p;
    obj_out[o * 3 + 0] = delay; obj_out[o * 3 + 1] = util; obj_out[o * 3 + 2] = comb;
    flags_out[o] = fl;
    if (over_out) over_out[o] = over;
  }
}

// ---- capacity-aware routing of one placement per instance (block per instance) -----------------------
__global__ void __launch_bounds__(256) k_route_cap(int N, int F, const double* __restrict__ d0,
                                                   const double* __restrict__ w0, const double* __restrict__ r0,
                                                   const double* __restrict__ K0, const uint8_t* __restrict__ c0,
                                                   uint8_t* __restrict__ cout0, double* __restrict__ x0,
                                                   double* __restrict__ n0, double* __restrict__ obj,
                                                   int32_t* __restrict__ feas, char* ws, int64_t ws_stride) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t fn = (int64_t)F * N;
  char* p = ws + (int64_t)b * ws_stride;
  CapRoute q;
  q.N = N; q.F = F;
  q.d = d0 + (int64_t)b * N * N; q.w = w0 + b * fn; q.r = r0 + b * fn; q.Kj = K0 + (int64_t)b * N;
  q.c = c0 + b * fn;
  q.th = (double*)p; p += fn * 8; q.rho = (double*)p; p += fn * 8;
  q.load = (double*)p; p += (int64_t)N * 8; q.lam = (double*)p; p += (int64_t)N * 8;
  q.ch = (int*)p; p += fn * 4; q.sec = (int*)p; p += fn * 4; q.alt = (int*)p;
  __shared__ double red[32];
  __shared__ double sh[32];
  __shared__ int bad;
  if (tid == 0) bad = 0;
  const CapResult cr = cap_route(q, 4 * N + 16, red, sh);
  uint8_t* cout = cout0 + b * fn;
  double* x = x0 + (int64_t)b * N * fn;
  // x arrives zero-filled (cudaMemsetAsync in the launcher: one block per instance cannot fill 6.4 GB at C4)
  for (int64_t k = tid; k < fn; k += blockDim.x) cout[k] = q.c[k];
  __syncthreads();
  if (cr.feasible) {
    for (int fi = tid; fi < (int)fn; fi += blockDim.x) {
      const int f = fi / N, i = fi - f * N;
      double* xr = x + ((int64_t)i * F + f) * N;
      xr[q.ch[fi]] = q.th[fi];
      if (q.sec[fi] >= 0) xr[q.sec[fi]] = 1.0 - q.th[fi];
    }
    __syncthreads();
    // A split can leave a pod with a total share below 1 - eps (C1b, constraints_step1.py:12-15).  Sources
    // without workload for f (w[f,i] == 0) route for free -- no delay cost, no CPU load -- so they top the
    // starved pod up, exactly what the MIP does with its free x columns.  One thread per function.
    for (int f = tid; f < F; f += blockDim.x) {
      for (int j = 0; j < N; ++j) {
        if (!q.c[(int64_t)f * N + j]) continue;
        double rc = 0.0;
        for (int i = 0; i < N; ++i) rc += x[((int64_t)i * F + f) * N + j];
        if (rc == 0.0 || rc + kEps >= 1.0) continue;
        double deficit = 1.0 - rc;
        for (int i = 0; i < N && deficit > 0.0; ++i) {
          if (q.w[(int64_t)f * N + i] != 0.0) continue;
          double* xr = x + ((int64_t)i * F + f) * N;
          for (int p2 = 0; p2 < N && deficit > 0.0; ++p2) {
            if (p2 == j || xr[p2] <= 0.0) continue;
            double rp = 0.0;                                  // donor pod must keep a share >= 1
            for (int i2 = 0; i2 < N; ++i2) rp += x[((int64_t)i2 * F + f) * N + p2];
            const double amt = fmin(fmin(xr[p2], rp - 1.0), deficit);
            if (amt <= 0.0) continue;
            xr[p2] -= amt; xr[j] += amt; deficit -= amt;
          }
        }
      }
    }
    __syncthreads();
    // pods that serve nobody are closed (free for the objective); a pod still short of its share is
    // reported (feas_out = 0) -- the caller falls back to another placement
    for (int fj = tid; fj < (int)fn; fj += blockDim.x) {
      if (!q.c[fj]) continue;
      const int f = fj / N, j = fj - f * N;
      double rc = 0.0;
      for (int i = 0; i < N; ++i) rc += x[((int64_t)i * F + f) * N + j];
      if (rc == 0.0) cout[fj] = 0;
      else if (rc + kEps < 1.0) bad = 1;
    }
  }
  __syncthreads();
  for (int j = tid; j < N; j += blockDim.x) {
    int any = 0;
    for (int f = 0; f < F; ++f) any |= cout[(int64_t)f * N + j];
    n0[(int64_t)b * N + j] = any ? 1.0 : 0.0;
  }
  if (tid == 0) { obj[b] = cr.cost; feas[b] = (cr.feasible && !bad) ? 1 : 0; }
}

}  // namespace neptune

using namespace neptune;

extern "C" int neptune_route_capacitated(int B, int N, int F, const double* d, const double* w, const double* r,
                                         const double* Kj, const uint8_t* c, uint8_t* c_out, double* x,
                                         double* n, double* obj_out, int32_t* feas_out, void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || !d || !w || !r || !Kj || !c || !c_out || !x || !n || !obj_out || !feas_out)
    return NEPTUNE_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t fn = (int64_t)F * N;
  int64_t stride = 2 * fn * 8 + 2 * (int64_t)N * 8 + 3 * fn * 4;
  stride = (stride + 255) & ~(int64_t)255;
  char* ws = nullptr;
  NEPTUNE_CUDA_OK(cudaMallocAsync(&ws, (size_t)stride * B, s));
  NEPTUNE_CUDA_OK(cudaMemsetAsync(x, 0, (size_t)B * N * fn * 8, s));
  { k_route_cap<<<B, 256, 0, s>>>(N, F, d, w, r, Kj, c, c_out, x, n, obj_out, feas_out, ws, stride); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  NEPTUNE_CUDA_OK(cudaFreeAsync(ws, s));
  return 0;
}

extern "C" int neptune_check_solution(int B, int N, int F, double alpha, const double* d, const double* w,
                                      const double* r, const double* m, const double* Mj, const double* Kj,
                                      const double* maxd, const double* cost, double budget, const double* x,
                                      const double* c, const double* n, int32_t* flags_out,
                                      double* scores_out, void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || !d || !w || !r || !m || !Mj || !Kj || !x || !c || !n || !flags_out ||
      !scores_out)
    return NEPTUNE_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  double* scratch = nullptr;
  NEPTUNE_CUDA_OK(cudaMallocAsync(&scratch, (size_t)B * kScr * 8, s));
  Inst in{N, F, d, w, r, m, Mj, Kj, maxd, cost, budget, alpha};
  { k_check_pre<<<B, 256, 0, s>>>(in, scratch, flags_out); NEPTUNE_COUNT(1); }
  int64_t work = (int64_t)N * F * N;
  int gx = (int)((work + 255) / 256);
  if (gx > kNumSMs * 8) gx = kNumSMs * 8;
  if (gx < 1) gx = 1;
  { k_check<<<dim3(gx, B), 256, 0, s>>>(in, x, c, n, scratch, flags_out); NEPTUNE_COUNT(1); }
  { k_check_fin<<<(B + 127) / 128, 128, 0, s>>>(B, N, alpha, scratch, scores_out); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  NEPTUNE_CUDA_OK(cudaFreeAsync(scratch, s));
  return 0;
}

extern "C" int neptune_route_placements(int B, int N, int F, const double* d, const uint8_t* c, double* x,
                                        double* n, void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || !d || !c || !x) return NEPTUNE_E_ARG;
  { k_route<<<dim3(F, B), 256, 0, (cudaStream_t)stream>>>(N, F, d, c, x, n); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  return 0;
}

extern "C" int neptune_eval_placements(int B, int P, int N, int F, double alpha, const double* d,
                                       const double* w, const double* r, const double* m, const double* Mj,
                                       const double* Kj, const double* maxd, const double* cost, double budget,
                                       const uint8_t* c, double* obj_out, int32_t* flags_out,
                                       double* overload_out, void* stream) {
  if (B <= 0 || P <= 0 || N <= 0 || F <= 0 || !d || !w || !r || !m || !Mj || !Kj || !c || !obj_out ||
      !flags_out)
    return NEPTUNE_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  double* scratch = nullptr;
  int32_t* dummy = nullptr;
  NEPTUNE_CUDA_OK(cudaMallocAsync(&scratch, (size_t)B * kScr * 8, s));
  NEPTUNE_CUDA_OK(cudaMallocAsync(&dummy, (size_t)B * 4, s));
  Inst in{N, F, d, w, r, m, Mj, Kj, maxd, cost, budget, alpha};
  { k_check_pre<<<B, 256, 0, s>>>(in, scratch, dummy); NEPTUNE_COUNT(1); }
  size_t sm = (size_t)N * (8 + 8 + 4);
  { k_eval<<<dim3(P, B), 256, sm, s>>>(in, P, c, scratch, obj_out, flags_out, overload_out); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  NEPTUNE_CUDA_OK(cudaFreeAsync(scratch, s));
  NEPTUNE_CUDA_OK(cudaFreeAsync(dummy, s));
  return 0;
}
