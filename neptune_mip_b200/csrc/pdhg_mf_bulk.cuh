// pdhg_mf_bulk.cuh -- the iteration pass of the matrix-free PDHG with the streams staged through shared memory by the
// bulk-copy engine (cp.async.bulk + mbarrier; SASS UBLKCP / UBLKRED), for instances whose function slab fits a stage:
// even N <= 64 (C2: 50 x 10), i.e. the shapes the pair pass k_mf_iter2 serves.  Included by pdhg_mf.cu.
//
// Why: the register passes keep every byte in flight in a register -- 2 blocks x 8 warps x 160 B per lane is ~64 KB per SM
// at N = 50 (25 of 32 lanes), and nothing is in flight while a warp computes; ncu shows 16 warps per SM waiting on their own
// loads (long scoreboard 7.8 of 9.6 stalled warps per issue, profiles/r02_pdhg_variants.md).  Here the bytes in flight live
// in shared memory (up to 200 KB per SM) and are requested by one thread:
//
//   * a persistent block per SM: one producer warp (one elected lane) + NW consumer warps (+ two small-vector warps, FUSE);
//     a tile is the whole N x N slab of one (instance, function), contiguous in every stream, so a tile is a handful of
//     one-dimensional bulk copies: x, yS (xsum, ysum), the delay matrix, w[f,:], r[f,:] (and, without FUSE, y3[f,:], y4,
//     cbar[f,:], the multiplier pairs of C1) plus a header {tau, sigma / 2, live} the producer writes -- on the tile path a
//     consumer reads NOTHING from global memory (with the vectors read from global a tile cost 4.5 us of dependent L2 round
//     trips, all warps in lockstep);
//   * STAGES stages, each with a `full` mbarrier (producer: arrive.expect_tx, the copies complete the transaction count)
//     and a `done` mbarrier (one arrival per consumer warp, after fence.proxy.async: the stores of the async proxy read
//     what the generic proxy wrote);
//   * consumers update x and yS IN PLACE in shared memory (same arithmetic as k_mf_iter2: pairs of adjacent columns per
//     lane, reciprocal by Newton steps); the producer writes the stage back with bulk stores and refills it after
//     cp.async.bulk.wait_group.read;
//   * RED = true: the running sums xsum, ysum are never loaded -- `cp.reduce.async.bulk ... add.f64` adds the new x / yS
//     from the same shared-memory buffers to them in L2 (one IEEE addition per element and iteration, so the value equals
//     xsum + x+ of the other passes bit for bit); a stage is then two read-write slabs, and more stages fit;
//     RED = false: xsum, ysum travel through the stage like x and yS (four read-write slabs);
//   * column sums: per-lane accumulators over the rows of a warp, then one cross-warp reduction per tile through a small
//     shared buffer in warp order (fixed summation order: runs are bit-reproducible); row sums by warp shuffle;
//   * FUSE = true: the small-vector update (k_mf_small's arithmetic) runs inside the launch -- two small-vector warps, one
//     instance ahead of the consumers, two shared buffers handed over by mbarriers; small state and partial sums alternate
//     between two sets from launch to launch (bulk_small_pre, BulkFuse); the chunk ends with k_mf_small_from.  An
//     iteration is then ONE launch: 87.6 us at C2 x 256 against 94.6 with k_mf_small launches and 125 for the pair pass;
//   * slabs are enumerated over the instances still iterating (`live`, rebuilt by k_mf_live at every KKT check).
// The partial-sum buffers (P1, P4, PS: one row tile per slab, P3i: one column segment) have the layout k_mf_iter2 fills, so
// k_mf_small and the KKT passes are shared.  Every spin on an mbarrier is bounded (trap after ~2 s of SM clocks).
// Measurements, step by step: profiles/r02c_pdhg_bulk.md.
#pragma once

namespace neptune {

constexpr int kBulkMaxStages = 8;
constexpr int kRows = 1;                          // rows of a consumer warp in flight together
constexpr int kBulkSmallWarps = 2;                // warps of the in-pass small-vector update (FUSE)
constexpr int kBulkMaxWarps = 15;                 // consumer warps
constexpr size_t kBulkSmemMax = 227 * 1024;       // dynamic shared memory a block may opt in to on sm_100a

struct BulkCfg {
  int ok;                 // the shape fits (>= 2 stages)
  int red;                // running sums by bulk reduction
  int nw;                 // consumer warps
  int stages;
  unsigned slab_bytes;    // N * N * 8
  unsigned slab_stride;   // slab_bytes rounded up to 128
  unsigned streams;       // read-write slabs of a stage: 2 (red) or 4; the d slab follows them
  unsigned vec_off;       // offset of the small vectors inside a stage: w[N] r[N] y3[N] y4[N] cbar[N] y1[2N] hdr[4]
  unsigned stage_bytes;
  size_t smem;            // dynamic shared memory of the launch
  int fuse;               // the small-vector update of an instance runs inside the pass (bulk_small_pre), see k_mf_iter_bulk
  unsigned inst_bytes;    // shared memory of the per-instance vectors y1[C] y3[C] cbar[C] y4[N] (fuse), rounded to 128
  int diag;               // measurements only (tools/mf_check.py): 1 = no copies (arithmetic on whatever the stage holds),
                          // 2 = no arithmetic (copy through), 3 = loads only; results are meaningless then
};

// consumer warps: rows go round robin over the warps, so take the fewest rounds a block of at most 512 threads (128
// registers each) allows and then the fewest warps that still need no more rounds (N = 50: 4 rounds, 13 warps)
static int bulk_pick_warps(int N) {
  const int rounds = (N + kBulkMaxWarps - 1) / kBulkMaxWarps;
  const int nw = (N + rounds - 1) / rounds;
  return nw < 4 ? 4 : nw;
}

static BulkCfg bulk_config(int N, int red, int nw_override, int stage_cap, int F = 0, int fuse = 0) {
  BulkCfg c{};
  c.red = red ? 1 : 0;
  if ((N & 1) || N > 64 || N < 2) return c;
  c.fuse = (fuse && F > 0) ? 1 : 0;
  c.inst_bytes = c.fuse ? (((unsigned)(3 * F * N + N) * 8u + 127u) & ~127u) : 0u;      // ONE of the two buffers
  c.nw = (nw_override >= 1 && nw_override <= kBulkMaxWarps) ? nw_override : bulk_pick_warps(N);
  if (c.fuse && c.nw > kBulkMaxWarps - kBulkSmallWarps) {        // the block also holds the small-vector warps
    const int cap = kBulkMaxWarps - kBulkSmallWarps, rounds = (N + cap - 1) / cap;
    c.nw = (N + rounds - 1) / rounds;
    if (c.nw < 4) c.nw = 4;
  }
  c.slab_bytes = (unsigned)N * (unsigned)N * 8u;
  c.slab_stride = (c.slab_bytes + 127u) & ~127u;
  c.streams = red ? 2u : 4u;
  c.vec_off = (c.streams + 1u) * c.slab_stride;
  c.stage_bytes = c.vec_off + (((unsigned)(7 * N + 4) * 8u + 127u) & ~127u);
  const size_t fixed = 256 /* barriers */ + (size_t)3 * c.nw * N * 8 /* column partials */ + 256 /* alignment slack */ + 2 * (size_t)c.inst_bytes;
  if (fixed + 2 * (size_t)c.stage_bytes > kBulkSmemMax) return c;
  int s = (int)((kBulkSmemMax - fixed) / c.stage_bytes);
  if (s > kBulkMaxStages) s = kBulkMaxStages;
  if (stage_cap >= 2 && s > stage_cap) s = stage_cap;
  if (s < 2) return c;
  c.stages = s;
  c.smem = fixed + (size_t)s * c.stage_bytes;
  c.ok = 1;
  return c;
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// wait for the phase of parity `parity` to complete; a spin that lasts ~2 s of SM clocks is a protocol error: trap
// rather than hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (int spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin == 64) t0 = clock64();
    if (spin > 64 && (spin & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_add_f64(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }

// ---- the small-vector update of ONE instance, inside the pass (FUSE) ---------------------------------------------
// Same arithmetic, term for term, as k_mf_small (POST of the pass that ran before, PREC and Y2 of the iteration that starts)
// for the min-delay model, run by the consumer threads of a block before its first tile of an instance.  Every block that
// holds tiles of the instance computes it (it only reads what the PREVIOUS launch left: the partial sums and the small
// state `src`), keeps y1, y3, y4, cbar in shared memory for its tiles, and the block that owns tile f = 0 (`writer`)
// also writes the new small state to `dst` and advances the running sums.  `src` and `dst` are different buffers (the
// launches of a chunk alternate between two), so a block that is still reading never sees a half-written state.
struct BulkFuse {
  const double* src_q;    // [B][qstride] small state read by this launch unless it is the first of a chunk
  double* dst_q;          // [B][qstride] small state written by this launch: y rows [0, rs) | c [C] | cbar [C]
  const int* first;       // device flag (graph node 0): non-zero -> first launch of a chunk
  int force_first;        // host-known: first launch of a chunk (read the canonical arrays, no POST)
  int64_t qstride;        // rs + 2 C
  // partial sums: a launch reads the set the launch before wrote and writes the other one (a block may still be reading
  // the sums of an instance while another block, further into the same launch, writes that instance's new ones)
  const double *rP1, *rP4, *rPS, *rP3;
  double *wP1, *wP4, *wPS, *wP3;
};

// The update is a chain of dependent global round trips (microseconds each while the copies saturate the memory system), so
// every load of a batch of items is issued before the first use (kItems items of a thread in flight), and the running sums
// advance by fire-and-forget reductions (one addition per element and launch: the value of a load-add-store, bit for bit).
// Row and column tiles: rt = ct = 1 on this path, so a "sum of partials" is one value (0.0 + v, as strided_sum returns it).
constexpr int kSmallItems = 4;
__device__ __forceinline__ void bulk_small_pre(const MfGeo& G, const MfIn& in, const MfSt& st, const BulkFuse& fz, int b, bool post, bool writer,
                                               double tau, double sigma, const double* __restrict__ sy, const double* __restrict__ sc,
                                               const double* __restrict__ scbar, double* __restrict__ dy, double* __restrict__ dc,
                                               double* __restrict__ dcbar, double* iy1, double* iy3, double* icb, double* iy4,
                                               int tid, int nth) {
  const int N = G.N, F = G.F;
  const int C = (int)G.C;
  double* __restrict__ ys = st.ysum + (int64_t)b * G.rows;
  double* __restrict__ cs = st.xsum + (int64_t)b * G.cols + G.X;
  const double* __restrict__ P1 = fz.rP1 + (int64_t)b * C;
  const double* __restrict__ P4 = fz.rP4 + (int64_t)b * C;
  const double* __restrict__ PS = fz.rPS + (int64_t)b * C;
  const double* __restrict__ P3 = fz.rP3 + (int64_t)b * C;
  const double* __restrict__ m = in.m + (int64_t)b * F;
  const double s1 = sigma / (double)(N + 1), s3 = sigma / (double)N;
  for (int q0 = tid; q0 < C; q0 += kSmallItems * nth) {
    double l_y1[kSmallItems], l_y3[kSmallItems], l_p1[kSmallItems], l_cb[kSmallItems], l_p3[kSmallItems], l_ps[kSmallItems],
        l_y2[kSmallItems], l_c[kSmallItems], l_m[kSmallItems];
#pragma unroll
    for (int u = 0; u < kSmallItems; ++u) {
      const int q = q0 + u * nth;
      const int qq = q < C ? q : 0;
      const int f = qq / N, j = qq - f * N;
      l_y1[u] = sy[2 * qq + 1]; l_y3[u] = sy[G.r3 + qq]; l_ps[u] = PS[qq]; l_y2[u] = sy[G.r2 + j]; l_c[u] = sc[qq]; l_m[u] = m[f];
      l_p1[u] = post ? P1[qq] : 0.0; l_cb[u] = post ? scbar[qq] : 0.0; l_p3[u] = post ? P3[qq] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kSmallItems; ++u) {
      const int q = q0 + u * nth;
      if (q < C) {
        double y1 = l_y1[u], y3n = l_y3[u];
        if (post) {
          const double a1 = (0.0 + l_p1[u]) - l_cb[u];
          const double a3 = 0.0 + l_p3[u];
          const double v1 = y1 + s1 * a1;
          y1 = v1 - s1 * fmax(v1 / s1, -kEps);
          y3n = y3n + s3 * a3 - s3;
          if (writer) { atomicAdd(ys + 2 * q + 1, y1); atomicAdd(ys + G.r3 + q, y3n); }
        }
        const double sS = 0.0 + l_ps[u];
        const double mf = l_m[u];
        const double gc = -y1 + mf * l_y2[u] - sS;
        const double co = l_c[u];
        double cn = co - tau * gc / (1.0 + mf + (double)N);
        cn = fmin(fmax(cn, 0.0), 1.0);
        const double cb = 2.0 * cn - co;
        iy1[q] = y1; iy3[q] = y3n; icb[q] = cb;
        if (writer) { dy[2 * q + 1] = y1; dy[G.r3 + q] = y3n; dcbar[q] = cb; dc[q] = cn; atomicAdd(cs + q, cn); }
      }
    }
  }
  {
    // C4 dual: the F partials of a column, all requested at once, added in function order (strided_sum's order)
    const double* __restrict__ Kj = in.Kj + (int64_t)b * N;
    for (int j = tid; j < N; j += nth) {
      double yn = sy[G.r4 + j];
      if (post) {
        double a = 0.0;
        if (F <= 16) {
          double v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = k < F ? P4[(int64_t)k * N + j] : 0.0;
          const double s4 = st.S4[(int64_t)b * N + j], kj = Kj[j];
#pragma unroll
          for (int k = 0; k < 16; ++k) if (k < F) a += v[k];
          const double s = sigma * s4;
          const double vv = yn + s * a;
          yn = vv - s * fmin(vv / s, kj);
        } else {
          a = strided_sum(P4 + j, F, N);
          const double s = sigma * st.S4[(int64_t)b * N + j];
          const double vv = yn + s * a;
          yn = vv - s * fmin(vv / s, Kj[j]);
        }
        if (writer) atomicAdd(ys + G.r4 + j, yn);
      }
      iy4[j] = yn;
      if (writer) dy[G.r4 + j] = yn;
    }
  }
}

// C2 dual of the instance from the cbar just computed (shared memory); only the writer needs it (the pass does not read y2)
__device__ __forceinline__ void bulk_small_y2(const MfGeo& G, const MfIn& in, const MfSt& st, int b, double sigma,
                                              const double* __restrict__ sy, double* __restrict__ dy, const double* icb, int tid, int nth) {
  const int N = G.N, F = G.F;
  const double* __restrict__ m = in.m + (int64_t)b * F;
  const double* __restrict__ Mj = in.Mj + (int64_t)b * N;
  double* __restrict__ ys = st.ysum + (int64_t)b * G.rows;
  const double s = sigma * st.S2[b];
  for (int j = tid; j < N; j += nth) {
    const double y2 = sy[G.r2 + j], mj = Mj[j];
    double a = 0.0;
    for (int f = 0; f < F; ++f) a += m[f] * icb[f * N + j];
    const double vv = y2 + s * a;
    const double yn = vv - s * fmin(vv / s, mj);
    dy[G.r2 + j] = yn; atomicAdd(ys + G.r2 + j, yn);
  }
}

// projections as compare-and-select (fmin / fmax on doubles cost six instructions each for their NaN rules; the iterates are
// finite, and for finite arguments the value is the same -- only the sign of a zero may differ)
__device__ __forceinline__ double clip01(double v) { v = v < 0.0 ? 0.0 : v; return v > 1.0 ? 1.0 : v; }
__device__ __forceinline__ double max0(double v) { return v < 0.0 ? 0.0 : v; }

// ---- the pass -------------------------------------------------------------------------------------------------
// grid: one block per SM (at most one block per tile); block: (cfg.nw + 1) warps, the LAST warp is the producer.
// Tiles: slab = b * F + f.  !FUSE: tile n of block k is slab k + n * gridDim.x.  FUSE: a block owns a CONTIGUOUS run of
// slabs (so it meets few instances, and the small-vector update of an instance runs once per block that holds tiles of
// it).  Slabs of converged instances go through the barrier protocol without copies or arithmetic, so stage and phase are
// closed formulas of n for both roles.
// EVERYTHING a consumer reads arrives through shared memory: the slabs of x, yS (xsum, ysum), the delay matrix, the vectors
// w[f,:], r[f,:] and a header {tau, sigma / 2, live} written by the producer; y3[f,:], y4, cbar[f,:] and the (C1a, C1b)
// multiplier pairs come through the stage (!FUSE) or from the block's own small-vector update (FUSE) -- a consumer's only
// global accesses on the tile path are the stores of the partial sums (measured: with the vectors read from global memory a
// tile cost 4.5 us of dependent L2 round trips, all warps in lockstep).
template <bool RED, bool FUSE>
__global__ void __launch_bounds__((kBulkMaxWarps + 1) * 32, 1)
k_mf_iter_bulk(MfGeo G, MfIn in, MfSt st, const Ctl* __restrict__ ctl, int B, BulkCfg cfg, BulkFuse fz,
               const int* __restrict__ live) {
  extern __shared__ __align__(128) unsigned char bulk_smem_raw[];
  const int N = G.N, F = G.F, NW = cfg.nw, S = cfg.stages;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t NN = (int64_t)N * N;
  // `live` (or null): live[0] = number of instances still iterating, live[1..] = their indices in increasing order (rebuilt
  // at every KKT check, constant during a chunk).  Slabs are enumerated over the live instances only, so the work of a
  // batch whose instances converge one by one stays spread over every SM (the last live instance of a batch of 48 would
  // otherwise sit in three blocks).
  const int n_inst = live ? live[0] : B;
  const int64_t total = (int64_t)n_inst * F;
  auto inst_of = [&](int64_t slab, int& f) { const int q = (int)(slab / F); f = (int)(slab - (int64_t)q * F); return live ? live[1 + q] : q; };
  // shared-memory carve: [barriers 256 B][column partials 3 * NW * N doubles][per-instance vectors (FUSE)][stages]
  unsigned char* base = bulk_smem_raw;
  double* colbuf = reinterpret_cast<double*>(base + 256);
  const uint32_t base_addr = smem_addr(base);
  const uint32_t inst0 = (base_addr + 256u + (uint32_t)(3 * NW * N * 8) + 127u) & ~127u;
  const uint32_t stage0 = inst0 + 2u * cfg.inst_bytes;
  const uint32_t stage_bytes = cfg.stage_bytes;
  const uint32_t bar_full = base_addr, bar_done = base_addr + 8u * (uint32_t)S;
  const uint32_t bar_ifull = base_addr + 16u * (uint32_t)S, bar_ifree = bar_ifull + 16u;      // two instance-vector buffers (FUSE)
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8u * s, 1u); mbar_init(bar_done + 8u * s, (uint32_t)NW); }
    for (int q = 0; q < 2; ++q) { mbar_init(bar_ifull + 8u * q, (uint32_t)kBulkSmallWarps); mbar_init(bar_ifree + 8u * q, (uint32_t)NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  // tiles of this block
  int64_t first, step; int T;
  if (FUSE) {
    const int64_t q = total / gridDim.x, rem = total - q * gridDim.x;
    first = (int64_t)blockIdx.x * q + (blockIdx.x < rem ? blockIdx.x : rem);
    T = (int)(q + (blockIdx.x < rem ? 1 : 0));
    step = 1;
  } else {
    first = blockIdx.x; step = gridDim.x;
    T = first < total ? (int)((total - first + gridDim.x - 1) / gridDim.x) : 0;
  }
  const uint32_t vecb = (uint32_t)N * 8u;          // bytes of one N-vector (a multiple of 16: N is even)

  if (warp == NW) {
    // ------------------------------------------------------------ producer: one lane moves every byte
    if (lane == 0) {
      struct Prep { double tau, shalf; int conv; };
      auto prep_tile = [&](int n) {          // the control block of the tile's instance (requested ahead of the waits)
        const int64_t slab = first + (int64_t)n * step;
        int f_; const int b = inst_of(slab, f_);
        Prep p; p.tau = ctl[b].tau; p.shalf = 0.5 * ctl[b].sigma; p.conv = ctl[b].converged;
        return p;
      };
      auto load_tile = [&](int n, const Prep& pr) {
        const int64_t slab = first + (int64_t)n * step;
        int f; const int b = inst_of(slab, f);
        const int s = n % S;
        const uint32_t bar = bar_full + 8u * s, dst = stage0 + (uint32_t)s * stage_bytes;
        double* hdr = reinterpret_cast<double*>(base + (dst - base_addr) + cfg.vec_off) + 7 * N;
        hdr[0] = pr.tau; hdr[1] = pr.shalf; hdr[2] = (!pr.conv && cfg.diag < 2) ? 1.0 : 0.0;
        if (pr.conv || cfg.diag == 1) { mbar_arrive(bar); return; }
        const double* xg = st.x + (int64_t)b * G.cols + (int64_t)f * NN;
        const double* yb = st.y + (int64_t)b * G.rows;
        const double* sg = yb + G.rs + (int64_t)f * NN;
        mbar_arrive_expect_tx(bar, (cfg.streams + 1u) * cfg.slab_bytes + (FUSE ? 2u : 7u) * vecb);
        bulk_load(dst, xg, cfg.slab_bytes, bar);
        bulk_load(dst + cfg.slab_stride, sg, cfg.slab_bytes, bar);
        if (!RED) {
          bulk_load(dst + 2u * cfg.slab_stride, st.xsum + (int64_t)b * G.cols + (int64_t)f * NN, cfg.slab_bytes, bar);
          bulk_load(dst + 3u * cfg.slab_stride, st.ysum + (int64_t)b * G.rows + G.rs + (int64_t)f * NN, cfg.slab_bytes, bar);
        }
        bulk_load(dst + cfg.streams * cfg.slab_stride, in.d + (int64_t)b * NN, cfg.slab_bytes, bar);
        const uint32_t v = dst + cfg.vec_off;
        bulk_load(v, in.w + ((int64_t)b * F + f) * N, vecb, bar);
        bulk_load(v + vecb, in.r + ((int64_t)b * F + f) * N, vecb, bar);
        if (!FUSE) {
          bulk_load(v + 2u * vecb, yb + G.r3 + (int64_t)f * N, vecb, bar);
          bulk_load(v + 3u * vecb, yb + G.r4, vecb, bar);
          bulk_load(v + 4u * vecb, st.cbar + (int64_t)b * G.C + (int64_t)f * N, vecb, bar);
          bulk_load(v + 5u * vecb, yb + 2 * (int64_t)f * N, 2u * vecb, bar);
        }
      };
      auto store_tile = [&](int n) {
        const int64_t slab = first + (int64_t)n * step;
        int f; const int b = inst_of(slab, f);
        if (!ctl[b].converged && cfg.diag != 1 && cfg.diag != 3) {
          double* xg = st.x + (int64_t)b * G.cols + (int64_t)f * NN;
          double* sg = st.y + (int64_t)b * G.rows + G.rs + (int64_t)f * NN;
          double* xsg = st.xsum + (int64_t)b * G.cols + (int64_t)f * NN;
          double* ssg = st.ysum + (int64_t)b * G.rows + G.rs + (int64_t)f * NN;
          const uint32_t src = stage0 + (uint32_t)(n % S) * stage_bytes;
          bulk_store(xg, src, cfg.slab_bytes);
          bulk_store(sg, src + cfg.slab_stride, cfg.slab_bytes);
          if (RED) {
            bulk_add_f64(xsg, src, cfg.slab_bytes);
            bulk_add_f64(ssg, src + cfg.slab_stride, cfg.slab_bytes);
          } else {
            bulk_store(xsg, src + 2u * cfg.slab_stride, cfg.slab_bytes);
            bulk_store(ssg, src + 3u * cfg.slab_stride, cfg.slab_bytes);
          }
        }
        bulk_commit();          // an empty group for a converged slab keeps the group count equal to the tile count
      };
      for (int n = 0; n < S && n < T; ++n) load_tile(n, prep_tile(n));
      // with three or more stages the refill of a stage waits for the stores of the PREVIOUS tile only, so the stores of
      // the tile just finished drain while the next loads are issued
      const bool lag = S >= 3;
      for (int k = 0; k < T; ++k) {
        const int freed = lag ? k - 1 : k, nl = freed + S;
        const bool refill = freed >= 0 && nl < T;                     // loads go in tile order
        Prep pr{0.0, 0.0, 1};
        if (refill) pr = prep_tile(nl);
        mbar_wait(bar_done + 8u * (k % S), (uint32_t)((k / S) & 1));
        store_tile(k);
        if (lag) bulk_wait_read<1>(); else bulk_wait_read<0>();
        if (refill) load_tile(nl, pr);
      }
      bulk_wait_all();
    }
    __syncwarp();
    return;
  }

  // -------------------------------------------------------------- small-vector warps (FUSE): one update per instance
  // of the block's run, into one of two shared buffers, AHEAD of the consumers (the update is a chain of dependent L2
  // round trips -- 7 us under load -- that must not sit on the consumers' path)
  if (FUSE && warp > NW) {
    const int tid = (warp - NW - 1) * 32 + lane, nth = kBulkSmallWarps * 32;
    const bool chunk_first = fz.force_first || (fz.first && *fz.first);
    int m = 0;
    for (int n = 0; n < T; ++n) {
      const int64_t slab = first + n;
      int f; const int b = inst_of(slab, f);
      if (!(n == 0 || f == 0)) continue;
      const uint32_t ib = (uint32_t)(m & 1);
      if (m >= 2) mbar_wait(bar_ifree + 8u * ib, (uint32_t)(((m - 2) >> 1) & 1));
      if (!ctl[b].converged && cfg.diag == 0) {
        double* iy1 = reinterpret_cast<double*>(base + (inst0 - base_addr) + ib * cfg.inst_bytes);
        double* iy3 = iy1 + G.C; double* icb = iy3 + G.C; double* iy4 = icb + G.C;
        const double tau_b = ctl[b].tau, sigma_b = ctl[b].sigma;
        const bool writer = f == 0;          // the block that owns the instance's first slab
        const double* sy = chunk_first ? st.y + (int64_t)b * G.rows : fz.src_q + (int64_t)b * fz.qstride;
        const double* sc = chunk_first ? st.x + (int64_t)b * G.cols + G.X : sy + G.rs;
        const double* scb = chunk_first ? st.cbar + (int64_t)b * G.C : sc + G.C;
        double* dy = fz.dst_q + (int64_t)b * fz.qstride;
        bulk_small_pre(G, in, st, fz, b, !chunk_first, writer, tau_b, sigma_b, sy, sc, scb, dy, dy + G.rs, dy + G.rs + G.C,
                       iy1, iy3, icb, iy4, tid, nth);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ifull + 8u * ib);          // the consumers need nothing of the C2 dual
        if (writer) {
          asm volatile("bar.sync 2, %0;" ::"r"(nth) : "memory");  // every cbar of the instance is in the buffer
          bulk_small_y2(G, in, st, b, sigma_b, sy, dy, icb, tid, nth);
        }
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ifull + 8u * ib);
      }
      ++m;
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  const int nthr = NW * 32;
  const unsigned jc = (unsigned)(2 * lane);
  const bool vj = (int)jc < N;
  const unsigned jl = vj ? jc : 0u;
  // per-instance vectors of the fused small-vector update (written by the small-vector warps)
  double *iy1 = nullptr, *iy3 = nullptr, *icb = nullptr, *iy4 = nullptr;
  int m_inst = 0; uint32_t cur_ib = 0;
  for (int n = 0; n < T; ++n) {
    const int64_t slab = first + (int64_t)n * step;
    int f; const int b = inst_of(slab, f);
    const int s = n % S;
    const uint32_t ph = (uint32_t)((n / S) & 1);
    const bool inst_last = FUSE && (f == F - 1 || n == T - 1);
    if (FUSE && (n == 0 || f == 0)) {          // first tile of an instance in this block: its vectors
      cur_ib = (uint32_t)(m_inst & 1);
      mbar_wait(bar_ifull + 8u * cur_ib, (uint32_t)((m_inst >> 1) & 1));
      iy1 = reinterpret_cast<double*>(base + (inst0 - base_addr) + cur_ib * cfg.inst_bytes);
      iy3 = iy1 + G.C; icb = iy3 + G.C; iy4 = icb + G.C;
      ++m_inst;
    }
    mbar_wait(bar_full + 8u * s, ph);
    unsigned char* stage_ptr = base + (stage0 - base_addr) + (size_t)s * stage_bytes;
    double* xs_ = reinterpret_cast<double*>(stage_ptr);
    double* ss_ = reinterpret_cast<double*>(stage_ptr + cfg.slab_stride);
    double* xsum_ = reinterpret_cast<double*>(stage_ptr + 2u * cfg.slab_stride);     // only when !RED
    double* ysum_ = reinterpret_cast<double*>(stage_ptr + 3u * cfg.slab_stride);
    const double* d_ = reinterpret_cast<const double*>(stage_ptr + cfg.streams * cfg.slab_stride);
    const double* vw = reinterpret_cast<const double*>(stage_ptr + cfg.vec_off);
    const double *vr = vw + N, *hdr = vw + 7 * N;
    const double *vy3 = FUSE ? iy3 + f * N : vw + 2 * N, *vy4 = FUSE ? iy4 : vw + 3 * N, *vcb = FUSE ? icb + f * N : vw + 4 * N;
    const double* vy1 = FUSE ? iy1 + f * N : vw + 5 * N;
    if (hdr[2] == 0.0) {          // converged instance (or a copy-only measurement): release the stage untouched
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_done + 8u * s); if (inst_last) mbar_arrive(bar_ifree + 8u * cur_ib); }
      continue;
    }
    const double tau = hdr[0], shalf = hdr[1];
    double* __restrict__ P3 = (FUSE ? fz.wP3 : st.P3i) + ((int64_t)b * G.C + (int64_t)f * N) * G.cti;
    double y1j[2], rj[2], rr4[2], cb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const unsigned je = jl + e;
      y1j[e] = FUSE ? vy1[je] : vy1[2 * je + 1];
      rj[e] = vr[je];
      rr4[e] = rj[e] * vy4[je];
      cb[e] = vcb[je];
    }
    const double r_tot = (int)threadIdx.x < N ? vr[threadIdx.x] : 0.0;      // for the totals below: read before the stage is released
    double a1[2] = {0.0, 0.0}, a4[2] = {0.0, 0.0}, aS[2] = {0.0, 0.0};
    // kRows rows of a warp are in flight together (independent dependency chains: the arithmetic is latency-bound);
    // a column's sums still take its rows in increasing order
    for (int i0 = warp; i0 < N; i0 += kRows * NW) {
      double wfi[kRows], ty[kRows]; double2 xv[kRows], sv[kRows], dv[kRows], xq[kRows], sq[kRows];
      bool ok[kRows]; unsigned o[kRows];
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        const int i = i0 + u * NW;
        ok[u] = i < N;
        const int ir = ok[u] ? i : i0;
        wfi[u] = vw[ir]; ty[u] = vy3[ir];
        o[u] = (unsigned)ir * (unsigned)N + jl;
        xv[u] = *reinterpret_cast<const double2*>(xs_ + o[u]);
        sv[u] = *reinterpret_cast<const double2*>(ss_ + o[u]);
        dv[u] = *reinterpret_cast<const double2*>(d_ + o[u]);
        if (!RED) {
          xq[u] = *reinterpret_cast<const double2*>(xsum_ + o[u]);
          sq[u] = *reinterpret_cast<const double2*>(ysum_ + o[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        double rsum = 0.0;
        if (vj && ok[u]) {
          double xn[2], sn[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double xo = e ? xv[u].y : xv[u].x, so = e ? sv[u].y : sv[u].x, de = e ? dv[u].y : dv[u].x;
            const double wr = fabs(wfi[u] * rj[e]);
            const double g = __dmul_rn(de, wfi[u]) + y1j[e] + ty[u] + wfi[u] * rr4[e] + so;
            double x1 = xo - tau * g * mf_rcp(3.0 + wr);
            x1 = clip01(x1);
            const double xb = 2.0 * x1 - xo;
            const double s1 = max0(so + shalf * (xb - cb[e]));
            xn[e] = x1; sn[e] = s1;
            a1[e] += xb; a4[e] += wfi[u] * xb; aS[e] += s1; rsum += xb;
          }
          *reinterpret_cast<double2*>(xs_ + o[u]) = make_double2(xn[0], xn[1]);
          *reinterpret_cast<double2*>(ss_ + o[u]) = make_double2(sn[0], sn[1]);
          if (!RED) {
            *reinterpret_cast<double2*>(xsum_ + o[u]) = make_double2(xq[u].x + xn[0], xq[u].y + xn[1]);
            *reinterpret_cast<double2*>(ysum_ + o[u]) = make_double2(sq[u].x + sn[0], sq[u].y + sn[1]);
          }
        }
        if (ok[u]) {          // uniform over the warp
          rsum = warp_sum(rsum);
          if (lane == 0) P3[(int64_t)(i0 + u * NW) * G.cti] = rsum;
        }
      }
    }
    // this warp's writes to the stage are done: make them visible to the async proxy, then release the stage
    fence_async_smem();
    __syncwarp();
    if (lane == 0) { mbar_arrive(bar_done + 8u * s); if (inst_last) mbar_arrive(bar_ifree + 8u * cur_ib); }
    // column sums: per-warp partials -> totals in warp order
    if (vj) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        colbuf[(0 * NW + warp) * N + jc + e] = a1[e];
        colbuf[(1 * NW + warp) * N + jc + e] = a4[e];
        colbuf[(2 * NW + warp) * N + jc + e] = aS[e];
      }
    }
    consumer_sync(nthr);
    if ((int)threadIdx.x < N) {
      const int j = threadIdx.x;
      double t1 = 0.0, t4 = 0.0, tS = 0.0;
      for (int q = 0; q < NW; ++q) {
        t1 += colbuf[(0 * NW + q) * N + j]; t4 += colbuf[(1 * NW + q) * N + j]; tS += colbuf[(2 * NW + q) * N + j];
      }
      const int64_t o = ((int64_t)b * F + f) * (int64_t)G.rt * N + j;      // rt = 1: one row tile per slab
      (FUSE ? fz.wP1 : st.P1)[o] = t1;
      (FUSE ? fz.wP4 : st.P4)[o] = r_tot * t4;
      (FUSE ? fz.wPS : st.PS)[o] = tS;
    }
    consumer_sync(nthr);
  }
}

// end of a chunk of fused launches: POST of the last pass from the small state `q` (same arithmetic as k_mf_small), and the
// whole small state back into the canonical arrays (y rows [0, rs), the c columns, cbar).  One block per instance.
__global__ void __launch_bounds__(256)
k_mf_small_from(MfGeo G, MfIn in, MfSt st, const Ctl* __restrict__ ctl, const double* __restrict__ qbuf, int64_t qstride,
                const double* __restrict__ rP1, const double* __restrict__ rP4, const double* __restrict__ rP3) {
  const int b = blockIdx.x;
  if (ctl[b].converged) return;
  const int N = G.N, F = G.F, rt = G.rt, ct = G.cti;
  const int C = (int)G.C;
  const double sigma = ctl[b].sigma;
  const double* __restrict__ sy = qbuf + (int64_t)b * qstride;
  const double* __restrict__ sc = sy + G.rs;
  const double* __restrict__ scbar = sc + C;
  double* __restrict__ y = st.y + (int64_t)b * G.rows;
  double* __restrict__ ys = st.ysum + (int64_t)b * G.rows;
  double* __restrict__ c = st.x + (int64_t)b * G.cols + G.X;
  double* __restrict__ cbar = st.cbar + (int64_t)b * C;
  const double* __restrict__ P1 = rP1 + (int64_t)b * F * rt * N;
  const double* __restrict__ P4 = rP4 + (int64_t)b * F * rt * N;
  const double* __restrict__ P3 = rP3 + (int64_t)b * C * ct;
  const double* __restrict__ Kj = in.Kj + (int64_t)b * N;
  const int K4 = F * rt;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const double a = strided_sum(P4 + j, K4, N);
    const double s = sigma * st.S4[(int64_t)b * N + j];
    const double v = sy[G.r4 + j] + s * a;
    const double yn = v - s * fmin(v / s, Kj[j]);
    y[G.r4 + j] = yn; ys[G.r4 + j] += yn;
    y[G.r2 + j] = sy[G.r2 + j];
  }
  const double s1 = sigma / (double)(N + 1), s3 = sigma / (double)N;
  for (int q = threadIdx.x; q < C; q += blockDim.x) {
    const int f = q / N, j = q - f * N;
    const int64_t po = (int64_t)f * rt * N + j;
    const double a1 = strided_sum(P1 + po, rt, N) - scbar[q];
    const double a3 = strided_sum(P3 + (int64_t)q * ct, ct, 1);
    const double v1 = sy[2 * q + 1] + s1 * a1;
    const double y1 = v1 - s1 * fmax(v1 / s1, -kEps);
    const double y3n = sy[G.r3 + q] + s3 * a3 - s3;
    y[2 * q + 1] = y1; ys[2 * q + 1] += y1;
    y[G.r3 + q] = y3n; ys[G.r3 + q] += y3n;
    c[q] = sc[q]; cbar[q] = scbar[q];
  }
}

}  // namespace neptune
