// assemble.cu -- (a) closed-form assembly of the placement model's constraint matrix on device.
//
// The reference builds the model through O(F*N^2) Python->SWIG calls
// (neptune/utils/variables.py:4-17, constraints_step1.py:5-103, objectives.py:4-52).  Every row's
// length, offset, column ids and coefficients are closed formulas of (N, F), so here each non-zero
// is written exactly once by index arithmetic: no atomics, no sort, fully coalesced stores.
// Algorithmic bytes: pattern 4*nnz (col_idx) + 4*nnz (colT_idx) + 8*(rows+cols); values
// 8*nnz (val) + 8*nnz (valT) + 8*(3*cols + 2*rows) -- pure HBM writes (inputs are O(F*N + N^2)).
#include "common.cuh"

namespace neptune {

// ---- pattern -------------------------------------------------------------------------------------
// One thread per non-zero of A (in CSR order).
__global__ void __launch_bounds__(256) k_pattern_csr(Layout L, int32_t* __restrict__ col_idx) {
  const int64_t n = L.N, f_ = L.F;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < L.nnz;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t col;
    if (e < L.e2) {                       // C1a / C1b : row = 2*(f*N+j)+ab, N x-entries then c
      int64_t row = e / (n + 1), k = e - row * (n + 1);
      int64_t fj = row >> 1, f = fj / n, j = fj - f * n;
      col = (k < n) ? (f * n * n + k * n + j) : (L.X + fj);
    } else if (e < L.e3) {                // C2 : row j, entries f
      int64_t t = e - L.e2, j = t / f_, f = t - j * f_;
      col = L.X + f * n + j;
    } else if (e < L.e4) {                // C3 : row (f,i), entries j  -> contiguous
      col = e - L.e3;
    } else if (e < L.e5) {                // C4 : row j, entries (f,i)
      int64_t t = e - L.e4, j = t / L.C, fi = t - j * L.C;
      col = fi * n + j;                   // f*N*N + i*N + j
    } else if (e < L.e6) {                // C5a / C5b : row = 2*j+ab, F c-entries then n
      int64_t t = e - L.e5, row = t / (f_ + 1), k = t - row * (f_ + 1), j = row >> 1;
      col = (k < f_) ? (L.X + k * n + j) : (L.X + L.C + j);
    } else if (e < L.es) {                // C6 : row j, one entry
      col = L.X + L.C + (e - L.e6);
    } else {                              // S : row q=(f,i,j): x col q, then c col (f,j)
      int64_t t = e - L.es, q = t >> 1;
      if ((t & 1) == 0) col = q;
      else { int64_t f = q / (n * n), j = q % n; col = L.X + f * n + j; }
    }
    col_idx[e] = (int32_t)col;
  }
}

// One thread per row: row_ptr.
__global__ void __launch_bounds__(256) k_pattern_rowptr(Layout L, int64_t* __restrict__ row_ptr) {
  const int64_t n = L.N, f_ = L.F;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= L.rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    int64_t p;
    if (r < L.r2) p = r * (n + 1);
    else if (r < L.r3) p = L.e2 + (r - L.r2) * f_;
    else if (r < L.r4) p = L.e3 + (r - L.r3) * n;
    else if (r < L.r5) p = L.e4 + (r - L.r4) * L.C;
    else if (r < L.r6) p = L.e5 + (r - L.r5) * (f_ + 1);
    else if (r < L.rs) p = L.e6 + (r - L.r6);
    else p = L.es + 2 * (r - L.rs);
    row_ptr[r] = p;
  }
}

// A^T: x column q has rows {C1a, C1b, C3, C4 [, S]}, c column has {C1a, C1b, C2 [, C5a, C5b] [, S x N]},
// n column has {C5a, C5b, C6}: all ascending.
__device__ __forceinline__ int64_t colT_offset(const Layout& L, int64_t col) {
  const int64_t xw = 4 + L.strengthen, cwT = L.cw + (L.strengthen ? L.N : 0);
  if (col <= L.X) return xw * col;
  if (col <= L.X + L.C) return xw * L.X + cwT * (col - L.X);
  return xw * L.X + cwT * L.C + 3 * (col - L.X - L.C);
}

__global__ void __launch_bounds__(256) k_pattern_T(Layout L, int64_t* __restrict__ rowT_ptr,
                                                   int32_t* __restrict__ colT_idx) {
  const int64_t n = L.N;
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col <= L.cols;
       col += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = colT_offset(L, col);
    if (rowT_ptr) rowT_ptr[col] = p;
    if (col == L.cols || !colT_idx) continue;
    if (col < L.X) {
      int64_t f = col / (n * n), rem = col - f * n * n, i = rem / n, j = rem - i * n;
      int64_t fj = f * n + j;
      colT_idx[p + 0] = (int32_t)(2 * fj);
      colT_idx[p + 1] = (int32_t)(2 * fj + 1);
      colT_idx[p + 2] = (int32_t)(L.r3 + f * n + i);
      colT_idx[p + 3] = (int32_t)(L.r4 + j);
      if (L.strengthen) colT_idx[p + 4] = (int32_t)(L.rs + col);
    } else if (col < L.X + L.C) {
      int64_t fj = col - L.X, f = fj / n, j = fj - f * n;
      colT_idx[p + 0] = (int32_t)(2 * fj);
      colT_idx[p + 1] = (int32_t)(2 * fj + 1);
      colT_idx[p + 2] = (int32_t)(L.r2 + j);
      int64_t q = p + 3;
      if (L.with_n) {
        colT_idx[q++] = (int32_t)(L.r5 + 2 * j);
        colT_idx[q++] = (int32_t)(L.r5 + 2 * j + 1);
      }
      if (L.strengthen)
        for (int64_t i = 0; i < n; ++i) colT_idx[q + i] = (int32_t)(L.rs + f * n * n + i * n + j);
    } else {
      int64_t j = col - L.X - L.C;
      colT_idx[p + 0] = (int32_t)(L.r5 + 2 * j);
      colT_idx[p + 1] = (int32_t)(L.r5 + 2 * j + 1);
      colT_idx[p + 2] = (int32_t)(L.r6 + j);
    }
  }
}

// ---- values ---------------------------------------------------------------------------------------
struct Inputs {
  const double *d, *w, *r, *m, *Mj, *Kj, *maxd, *cost;
  double budget, alpha;
};

__device__ __forceinline__ Inputs offset_inputs(const Inputs& in, int b, int N, int F) {
  Inputs o = in;
  const int64_t n = N, f = F;
  o.d = in.d + b * n * n; o.w = in.w + b * f * n; o.r = in.r + b * f * n; o.m = in.m + b * f;
  o.Mj = in.Mj + b * n; o.Kj = in.Kj + b * n;
  o.maxd = in.maxd ? in.maxd + b * f : nullptr;
  o.cost = in.cost ? in.cost + b * n : nullptr;
  return o;
}

__global__ void __launch_bounds__(256) k_values_csr(Layout L, Inputs in0, double* __restrict__ val) {
  const int b = blockIdx.y;
  const Inputs in = offset_inputs(in0, b, L.N, L.F);
  double* __restrict__ v = val + (int64_t)b * L.nnz;
  const int64_t n = L.N, f_ = L.F;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < L.nnz;
       e += (int64_t)gridDim.x * blockDim.x) {
    double out;
    if (e < L.e2) {
      int64_t row = e / (n + 1), k = e - row * (n + 1);
      out = (k < n) ? 1.0 : ((row & 1) ? -1.0 : -kBigM);
    } else if (e < L.e3) {
      int64_t t = e - L.e2, j = t / f_, f = t - j * f_;
      out = in.m[f];
    } else if (e < L.e4) {
      out = 1.0;
    } else if (e < L.e5) {
      int64_t t = e - L.e4, j = t / L.C, fi = t - j * L.C, f = fi / n;
      out = __dmul_rn(in.w[fi], in.r[f * n + j]);          // (x * w[f,i]) * r[f,j]
    } else if (e < L.e6) {
      int64_t t = e - L.e5, row = t / (f_ + 1), k = t - row * (f_ + 1);
      out = (k < f_) ? 1.0 : ((row & 1) ? -1.0 : -kBigM);
    } else if (e < L.es) {
      out = in.cost[e - L.e6];
    } else {
      out = ((e - L.es) & 1) ? -1.0 : 1.0;
    }
    v[e] = out;
  }
}

// wmax[b] = sum over (f,i) in reference order of w[f,i] * max{ d[i,j] : d[i,j] <= maxd[f] }
// (objectives.py:36-44).  Terms are produced in parallel, the sum itself is strictly left-to-right
// (one thread) so that non-integer data rounds exactly like the reference's Python loop.
__global__ void __launch_bounds__(256) k_wmax(int N, int F, Inputs in0, double* __restrict__ terms,
                                              double* __restrict__ wmax) {
  const int b = blockIdx.x;
  const Inputs in = offset_inputs(in0, b, N, F);
  double* t = terms + (int64_t)b * F * N;
  for (int fi = threadIdx.x; fi < F * N; fi += blockDim.x) {
    int f = fi / N, i = fi - f * N;
    double md = in.maxd[f], best = -INFINITY;
    for (int j = 0; j < N; ++j) {
      double dij = in.d[(int64_t)i * N + j];
      if (dij <= md && dij > best) best = dij;
    }
    t[fi] = __dmul_rn(in.w[fi], best);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int fi = 0; fi < F * N; ++fi) s = __dadd_rn(s, t[fi]);
    wmax[b] = s;
  }
}

// One thread per column: obj, bounds, integrality, and the values of A^T.
__global__ void __launch_bounds__(256) k_values_cols(Layout L, Inputs in0, const double* __restrict__ wmax,
                                                     double* __restrict__ valT, double* __restrict__ obj,
                                                     double* __restrict__ col_lb, double* __restrict__ col_ub,
                                                     uint8_t* __restrict__ col_int) {
  const int b = blockIdx.y;
  const Inputs in = offset_inputs(in0, b, L.N, L.F);
  const int64_t n = L.N;
  const double alpha = in.alpha;
  double wm = 0.0;
  if (L.kind == NEPTUNE_KIND_MIN_DELAY_UTIL) wm = wmax[b];
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < L.cols;
       col += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = colT_offset(L, col) + (int64_t)b * L.nnz;
    double o = 0.0, ub;
    if (col < L.X) {
      int64_t f = col / (n * n), rem = col - f * n * n, i = rem / n, j = rem - i * n;
      double wfi = in.w[f * n + i], dij = in.d[i * n + j];
      if (valT) {
        valT[p + 0] = 1.0; valT[p + 1] = 1.0; valT[p + 2] = 1.0;
        valT[p + 3] = __dmul_rn(wfi, in.r[f * n + j]);
        if (L.strengthen) valT[p + 4] = 1.0;
      }
      if (L.kind == NEPTUNE_KIND_MIN_DELAY) o = __dmul_rn(dij, wfi);
      else if (L.kind == NEPTUNE_KIND_MIN_DELAY_UTIL && wm != 0.0)   // delay term only if sum(w) != 0
        o = __ddiv_rn(__dmul_rn(__dmul_rn(1.0 - alpha, wfi), dij), wm);
      // x <= 1 is implied by C3 (sum_j x = 1, x >= 0); stating it in the strengthened model keeps the dual
      // objective finite, i.e. a valid lower bound at every PDHG iterate (the as-written model keeps +inf)
      ub = L.strengthen ? 1.0 : INFINITY;
    } else if (col < L.X + L.C) {
      int64_t fj = col - L.X, f = fj / n;
      if (valT) {
        valT[p + 0] = -kBigM; valT[p + 1] = -1.0; valT[p + 2] = in.m[f];
        int64_t q = p + 3;
        if (L.with_n) { valT[q++] = 1.0; valT[q++] = 1.0; }
        if (L.strengthen) for (int64_t i = 0; i < n; ++i) valT[q + i] = -1.0;
      }
      ub = 1.0;
    } else {
      int64_t j = col - L.X - L.C;
      if (valT) { valT[p + 0] = -kBigM; valT[p + 1] = -1.0; valT[p + 2] = in.cost[j]; }
      if (L.kind == NEPTUNE_KIND_MIN_UTIL) o = 1.0;
      else o = alpha / (double)L.N;
      ub = 1.0;
    }
    if (obj) obj[(int64_t)b * L.cols + col] = o;
    if (col_lb) col_lb[(int64_t)b * L.cols + col] = 0.0;
    if (col_ub) col_ub[(int64_t)b * L.cols + col] = ub;
    if (col_int && b == 0) col_int[col] = (col >= L.X) ? 1 : 0;
  }
}

// One thread per row: lo / hi.
__global__ void __launch_bounds__(256) k_values_rows(Layout L, Inputs in0, double* __restrict__ lo,
                                                     double* __restrict__ hi) {
  const int b = blockIdx.y;
  const Inputs in = offset_inputs(in0, b, L.N, L.F);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < L.rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    double l, h;
    // C1a (even rows) is implied by the S rows when strengthening: left as a free row there, because
    // its 1e6 coefficient is what stalls first-order methods (DESIGN.md "PDHG")
    if (r < L.r2) { if (r & 1) { l = -kEps; h = INFINITY; } else { l = -INFINITY; h = L.strengthen ? INFINITY : 0.0; } }
    else if (r < L.r3) { l = -INFINITY; h = in.Mj[r - L.r2]; }
    else if (r < L.r4) { l = 1.0; h = 1.0; }
    else if (r < L.r5) { l = -INFINITY; h = in.Kj[r - L.r4]; }
    else if (r < L.r6) { if ((r - L.r5) & 1) { l = -kEps; h = INFINITY; } else { l = -INFINITY; h = 0.0; } }
    else if (r < L.rs) { l = -INFINITY; h = in.budget; }
    else { l = -INFINITY; h = 0.0; }
    if (lo) lo[(int64_t)b * L.rows + r] = l;
    if (hi) hi[(int64_t)b * L.rows + r] = h;
  }
}

// sum(w) == 0 -> the combined objective has no delay term (objectives.py:34-35): encode as wmax = 0
__global__ void k_wmax_zero_if_no_workload(int N, int F, const double* __restrict__ w,
                                           double* __restrict__ wmax) {
  const int b = blockIdx.x;
  __shared__ double sm[32];
  double s = 0.0;
  for (int k = threadIdx.x; k < F * N; k += blockDim.x) s += w[(int64_t)b * F * N + k];
  s = block_sum(s, sm);
  if (threadIdx.x == 0 && s == 0.0) wmax[b] = 0.0;
}

static inline int grid_for(int64_t work, int threads = 256) {
  int64_t g = ceil_div(work, threads);
  int64_t cap = (int64_t)kNumSMs * 32;       // grid-stride beyond this
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace neptune

using namespace neptune;

extern "C" int neptune_abi_version(void) { return NEPTUNE_ABI_VERSION; }

static int check_shape(int N, int F, int kind, int flags) {
  if (N <= 0 || F <= 0) return NEPTUNE_E_ARG;
  if (kind < 0 || kind > 2) return NEPTUNE_E_ARG;
  if (flags & ~NEPTUNE_FLAG_STRENGTHEN) return NEPTUNE_E_ARG;
  Layout L(N, F, kind, flags);
  if (L.cols >= (int64_t)INT32_MAX || L.rows >= (int64_t)INT32_MAX) return NEPTUNE_E_SIZE;
  return 0;
}

extern "C" int neptune_model_sizes(int N, int F, int kind, int flags, int64_t* rows, int64_t* cols,
                                   int64_t* nnz) {
  int rc = check_shape(N, F, kind, flags);
  if (rc) return rc;
  Layout L(N, F, kind, flags);
  if (rows) *rows = L.rows;
  if (cols) *cols = L.cols;
  if (nnz) *nnz = L.nnz;
  return 0;
}

extern "C" int neptune_assemble_pattern(int N, int F, int kind, int flags, int64_t* row_ptr,
                                        int32_t* col_idx, int64_t* rowT_ptr, int32_t* colT_idx,
                                        void* stream) {
  int rc = check_shape(N, F, kind, flags);
  if (rc) return rc;
  Layout L(N, F, kind, flags);
  cudaStream_t s = (cudaStream_t)stream;
  if (col_idx) { k_pattern_csr<<<grid_for(L.nnz), 256, 0, s>>>(L, col_idx); NEPTUNE_COUNT(1); }
  if (row_ptr) { k_pattern_rowptr<<<grid_for(L.rows + 1), 256, 0, s>>>(L, row_ptr); NEPTUNE_COUNT(1); }
  if (rowT_ptr || colT_idx) { k_pattern_T<<<grid_for(L.cols + 1), 256, 0, s>>>(L, rowT_ptr, colT_idx); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  return 0;
}

extern "C" int neptune_assemble_values(int B, int N, int F, int kind, int flags, double alpha,
                                       const double* d, const double* w, const double* r,
                                       const double* m, const double* Mj, const double* Kj,
                                       const double* maxd, const double* cost, double budget,
                                       double* val, double* valT, double* obj, double* lo, double* hi,
                                       double* col_lb, double* col_ub, uint8_t* col_int, double* wmax,
                                       void* stream) {
  int rc = check_shape(N, F, kind, flags);
  if (rc) return rc;
  if (B <= 0 || !d || !w || !r || !m || !Mj || !Kj) return NEPTUNE_E_ARG;
  Layout L(N, F, kind, flags);
  if (L.with_n && !cost) return NEPTUNE_E_ARG;
  if (kind == NEPTUNE_KIND_MIN_DELAY_UTIL && (!maxd || !wmax || (!valT && !val))) return NEPTUNE_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  Inputs in{d, w, r, m, Mj, Kj, maxd, cost, budget, alpha};
  if (wmax && maxd) {
    // scratch for the F*N terms: borrow the head of val (or valT), overwritten afterwards
    double* scratch = val ? val : valT;
    if (!scratch) return NEPTUNE_E_ARG;
    if ((int64_t)F * N > L.nnz) return NEPTUNE_E_ARG;
    // terms for instance b live at scratch + b*F*N  (F*N <= nnz, so they fit in [B][nnz])
    { k_wmax<<<B, 256, 0, s>>>(N, F, in, scratch, wmax); NEPTUNE_COUNT(1); }
    { k_wmax_zero_if_no_workload<<<B, 256, 0, s>>>(N, F, w, wmax); NEPTUNE_COUNT(1); }
  }
  dim3 gc(grid_for(L.cols), B), gr(grid_for(L.rows), B), ge(grid_for(L.nnz), B);
  if (valT || obj || col_lb || col_ub || col_int)
    { k_values_cols<<<gc, 256, 0, s>>>(L, in, wmax, valT, obj, col_lb, col_ub, col_int); NEPTUNE_COUNT(1); }
  if (lo || hi) { k_values_rows<<<gr, 256, 0, s>>>(L, in, lo, hi); NEPTUNE_COUNT(1); }
  if (val) { k_values_csr<<<ge, 256, 0, s>>>(L, in, val); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  return 0;
}
