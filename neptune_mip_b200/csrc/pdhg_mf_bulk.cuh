// pdhg_mf_bulk.cuh -- the iteration pass of the matrix-free PDHG with the streams staged through shared memory by the
// bulk-copy engine (cp.async.bulk + mbarrier; SASS UBLKCP / UBLKRED), for instances whose function slab fits a stage:
// even N <= 64 (C2: 50 x 10), i.e. the shapes the pair pass k_mf_iter2 serves.  Included by pdhg_mf.cu.
//
// Why: the register passes keep every byte in flight in a register -- 2 blocks x 8 warps x 160 B per lane is ~64 KB per SM
// at N = 50 (25 of 32 lanes), and nothing is in flight while a warp computes; ncu shows 16 warps per SM waiting on their own
// loads (long scoreboard 7.8 of 9.6 stalled warps per issue, profiles/r02_pdhg_variants.md).  Here the bytes in flight live
// in shared memory (up to 200 KB per SM) and are requested by one thread:
//
//   * a persistent block per SM: one producer warp (one elected lane) + NW consumer warps; a tile is the whole N x N slab
//     of one (instance, function), contiguous in every stream, so a tile is 4 (or 2) one-dimensional bulk copies;
//   * STAGES stages, each with a `full` mbarrier (producer: arrive.expect_tx, the copies complete the transaction count)
//     and a `done` mbarrier (one arrival per consumer warp, after fence.proxy.async: the stores of the async proxy read
//     what the generic proxy wrote);
//   * consumers update x and yS IN PLACE in shared memory (same arithmetic as k_mf_iter2: pairs of adjacent columns per
//     lane, reciprocal by Newton steps); the producer writes the stage back with bulk stores and refills it after
//     cp.async.bulk.wait_group.read;
//   * RED = true: the running sums xsum, ysum are never loaded -- `cp.reduce.async.bulk ... add.f64` adds the new x / yS
//     from the same shared-memory buffers to them in L2 (one IEEE addition per element and iteration, so the value equals
//     xsum + x+ of the other passes bit for bit); a stage is then two streams, and twice as many stages fit;
//     RED = false: xsum, ysum travel through the stage like x and yS (four streams);
//   * column sums: per-lane accumulators over the rows of a warp, then one cross-warp reduction per tile through a small
//     shared buffer in warp order (fixed summation order: runs are bit-reproducible); row sums by warp shuffle.
// The partial-sum buffers (P1, P4, PS: one row tile per slab, P3i: one column segment) are the ones k_mf_iter2 fills, so
// k_mf_small and the KKT passes are shared.  Every spin on an mbarrier is bounded (trap after ~2 s of SM clocks).
#pragma once

namespace neptune {

constexpr int kBulkMaxStages = 8;
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

static BulkCfg bulk_config(int N, int red, int nw_override, int stage_cap) {
  BulkCfg c{};
  c.red = red ? 1 : 0;
  if ((N & 1) || N > 64 || N < 2) return c;
  c.nw = (nw_override >= 1 && nw_override <= kBulkMaxWarps) ? nw_override : bulk_pick_warps(N);
  c.slab_bytes = (unsigned)N * (unsigned)N * 8u;
  c.slab_stride = (c.slab_bytes + 127u) & ~127u;
  c.streams = red ? 2u : 4u;
  c.vec_off = (c.streams + 1u) * c.slab_stride;
  c.stage_bytes = c.vec_off + (((unsigned)(7 * N + 4) * 8u + 127u) & ~127u);
  const size_t fixed = 256 /* barriers */ + (size_t)3 * c.nw * N * 8 /* column partials */ + 128 /* alignment slack */;
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

// ---- the pass -------------------------------------------------------------------------------------------------
// grid: one block per SM (at most one block per tile); block: (cfg.nw + 1) warps, the LAST warp is the producer.
// tile n of block k is slab  k + n * gridDim.x  (slab = b * F + f); slabs of converged instances go through the barrier
// protocol without copies or arithmetic, so stage and phase are closed formulas of n for both roles.
// EVERYTHING a consumer reads arrives through the stage: the slabs of x, yS (xsum, ysum), the delay matrix, the vectors
// w[f,:], r[f,:], y3[f,:], y4, cbar[f,:], the (C1a, C1b) multiplier pairs of the function, and a header {tau, sigma / 2,
// live} written by the producer -- a consumer's only global accesses are the stores of the partial sums (measured: with
// the vectors read from global memory a tile cost 4.5 us of dependent L2 round trips, all warps in lockstep).
template <bool RED>
__global__ void __launch_bounds__((kBulkMaxWarps + 1) * 32, 1)
k_mf_iter_bulk(MfGeo G, MfIn in, MfSt st, const Ctl* __restrict__ ctl, int B, BulkCfg cfg) {
  extern __shared__ __align__(128) unsigned char bulk_smem_raw[];
  const int N = G.N, F = G.F, NW = cfg.nw, S = cfg.stages;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t NN = (int64_t)N * N;
  const int64_t total = (int64_t)B * F;
  // shared-memory carve: [barriers 256 B][column partials 3 * NW * N doubles][stages], stages 128-byte aligned
  unsigned char* base = bulk_smem_raw;
  double* colbuf = reinterpret_cast<double*>(base + 256);
  const uint32_t base_addr = smem_addr(base);
  const uint32_t stage0 = (base_addr + 256u + (uint32_t)(3 * NW * N * 8) + 127u) & ~127u;
  const uint32_t stage_bytes = cfg.stage_bytes;
  const uint32_t bar_full = base_addr, bar_done = base_addr + 8u * (uint32_t)S;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8u * s, 1u); mbar_init(bar_done + 8u * s, (uint32_t)NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  // number of tiles of this block
  const int64_t first = blockIdx.x;
  const int T = first < total ? (int)((total - first + gridDim.x - 1) / gridDim.x) : 0;
  const uint32_t vecb = (uint32_t)N * 8u;          // bytes of one N-vector (a multiple of 16: N is even)

  if (warp == NW) {
    // ------------------------------------------------------------ producer: one lane moves every byte
    if (lane == 0) {
      struct Prep { double tau, shalf; int conv; };
      auto prep_tile = [&](int n) {          // the control block of the tile's instance (requested ahead of the waits)
        const int64_t slab = first + (int64_t)n * gridDim.x;
        const int b = (int)(slab / F);
        Prep p; p.tau = ctl[b].tau; p.shalf = 0.5 * ctl[b].sigma; p.conv = ctl[b].converged;
        return p;
      };
      auto load_tile = [&](int n, const Prep& pr) {
        const int64_t slab = first + (int64_t)n * gridDim.x;
        const int b = (int)(slab / F), f = (int)(slab - (int64_t)b * F);
        const int s = n % S;
        const uint32_t bar = bar_full + 8u * s, dst = stage0 + (uint32_t)s * stage_bytes;
        double* hdr = reinterpret_cast<double*>(base + (dst - base_addr) + cfg.vec_off) + 7 * N;
        hdr[0] = pr.tau; hdr[1] = pr.shalf; hdr[2] = (!pr.conv && cfg.diag < 2) ? 1.0 : 0.0;
        if (pr.conv || cfg.diag == 1) { mbar_arrive(bar); return; }
        const double* xg = st.x + (int64_t)b * G.cols + (int64_t)f * NN;
        const double* yb = st.y + (int64_t)b * G.rows;
        const double* sg = yb + G.rs + (int64_t)f * NN;
        mbar_arrive_expect_tx(bar, (cfg.streams + 1u) * cfg.slab_bytes + 7u * vecb);
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
        bulk_load(v + 2u * vecb, yb + G.r3 + (int64_t)f * N, vecb, bar);
        bulk_load(v + 3u * vecb, yb + G.r4, vecb, bar);
        bulk_load(v + 4u * vecb, st.cbar + (int64_t)b * G.C + (int64_t)f * N, vecb, bar);
        bulk_load(v + 5u * vecb, yb + 2 * (int64_t)f * N, 2u * vecb, bar);
      };
      auto store_tile = [&](int n) {
        const int64_t slab = first + (int64_t)n * gridDim.x;
        const int b = (int)(slab / F), f = (int)(slab - (int64_t)b * F);
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

  // -------------------------------------------------------------- consumers
  const int nthr = NW * 32;
  const unsigned jc = (unsigned)(2 * lane);
  const bool vj = (int)jc < N;
  const unsigned jl = vj ? jc : 0u;
  for (int n = 0; n < T; ++n) {
    const int64_t slab = first + (int64_t)n * gridDim.x;
    const int b = (int)(slab / F), f = (int)(slab - (int64_t)b * F);
    const int s = n % S;
    const uint32_t ph = (uint32_t)((n / S) & 1);
    mbar_wait(bar_full + 8u * s, ph);
    unsigned char* stage_ptr = base + (stage0 - base_addr) + (size_t)s * stage_bytes;
    double* xs_ = reinterpret_cast<double*>(stage_ptr);
    double* ss_ = reinterpret_cast<double*>(stage_ptr + cfg.slab_stride);
    double* xsum_ = reinterpret_cast<double*>(stage_ptr + 2u * cfg.slab_stride);     // only when !RED
    double* ysum_ = reinterpret_cast<double*>(stage_ptr + 3u * cfg.slab_stride);
    const double* d_ = reinterpret_cast<const double*>(stage_ptr + cfg.streams * cfg.slab_stride);
    const double* vw = reinterpret_cast<const double*>(stage_ptr + cfg.vec_off);
    const double *vr = vw + N, *vy3 = vw + 2 * N, *vy4 = vw + 3 * N, *vcb = vw + 4 * N, *vy1 = vw + 5 * N, *hdr = vw + 7 * N;
    if (hdr[2] == 0.0) {          // converged instance (or a copy-only measurement): release the stage untouched
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_done + 8u * s);
      continue;
    }
    const double tau = hdr[0], shalf = hdr[1];
    double* __restrict__ P3 = st.P3i + ((int64_t)b * G.C + (int64_t)f * N) * G.cti;
    double y1j[2], rj[2], rr4[2], cb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const unsigned je = jl + e;
      y1j[e] = vy1[2 * je + 1];
      rj[e] = vr[je];
      rr4[e] = rj[e] * vy4[je];
      cb[e] = vcb[je];
    }
    const double r_tot = (int)threadIdx.x < N ? vr[threadIdx.x] : 0.0;      // for the totals below: read before the stage is released
    double a1[2] = {0.0, 0.0}, a4[2] = {0.0, 0.0}, aS[2] = {0.0, 0.0};
    for (int i = warp; i < N; i += NW) {
      const double wfi = vw[i], ty = vy3[i];
      double rsum = 0.0;
      if (vj) {
        const unsigned o = (unsigned)i * (unsigned)N + jc;
        const double2 xv = *reinterpret_cast<const double2*>(xs_ + o);
        const double2 sv = *reinterpret_cast<const double2*>(ss_ + o);
        const double2 dv = *reinterpret_cast<const double2*>(d_ + o);
        double xn[2], sn[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double xo = e ? xv.y : xv.x, so = e ? sv.y : sv.x, de = e ? dv.y : dv.x;
          const double wr = fabs(wfi * rj[e]);
          const double g = __dmul_rn(de, wfi) + y1j[e] + ty + wfi * rr4[e] + so;
          double x1 = xo - tau * g * mf_rcp(3.0 + wr);
          x1 = fmin(fmax(x1, 0.0), 1.0);
          const double xb = 2.0 * x1 - xo;
          const double s1 = fmax(so + shalf * (xb - cb[e]), 0.0);
          xn[e] = x1; sn[e] = s1;
          a1[e] += xb; a4[e] += wfi * xb; aS[e] += s1; rsum += xb;
        }
        *reinterpret_cast<double2*>(xs_ + o) = make_double2(xn[0], xn[1]);
        *reinterpret_cast<double2*>(ss_ + o) = make_double2(sn[0], sn[1]);
        if (!RED) {
          const double2 xq = *reinterpret_cast<const double2*>(xsum_ + o);
          const double2 sq = *reinterpret_cast<const double2*>(ysum_ + o);
          *reinterpret_cast<double2*>(xsum_ + o) = make_double2(xq.x + xn[0], xq.y + xn[1]);
          *reinterpret_cast<double2*>(ysum_ + o) = make_double2(sq.x + sn[0], sq.y + sn[1]);
        }
      }
      rsum = warp_sum(rsum);
      if (lane == 0) P3[(int64_t)i * G.cti] = rsum;
    }
    // this warp's writes to the stage are done: make them visible to the async proxy, then release the stage
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_done + 8u * s);
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
      st.P1[o] = t1;
      st.P4[o] = r_tot * t4;
      st.PS[o] = tS;
    }
    consumer_sync(nthr);
  }
}

}  // namespace neptune
