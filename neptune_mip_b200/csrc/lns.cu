// lns.cu -- (c2) LP-guided large-neighbourhood search over placements c[F][N], one WARP per search chain.
//
// Stands in for SCIP's branch and bound inside pywraplp.Solver.Solve() (reference core/solvers/solver.py:37)
// on the step-1 models whose objective has a delay term (NeptuneStep1CPUMinDelay / ...MinDelayAndUtilization,
// neptune_step1.py:63-77) when every function needs the same memory, so that the memory row of a node
// (constraints_step1.py:18-23) is a slot count  sum_f c[f,j] <= floor(M_j / m).
//
// State of a chain: a placement c and node prices lam[j] >= 0 on the CPU rows (constraints_step1.py:57-65).
// For fixed prices every source (f,i) goes to the open pod of f with the smallest PRICED delay
// d[i,j] + lam[j] r[f,j]; the priced cost  g(c, lam) = sum w min_j(d + lam r) - lam.K  is a lower bound of the
// routing LP of c and equals it at the best prices (LP duality for fixed c), so
//   * the prices are kept at the coordinate-wise maximiser for the current placement (dual ascent: the price of
//     an overloaded node is raised to the threshold at which enough flows leave, exactly the fractional
//     knapsack of route_cap.cuh, but only its dual side -- no flow is split here), and
//   * the placement moves are EXACT re-optimisations of all pods of k nodes at once: with the pods outside
//     the k nodes fixed, the priced cost separates over functions, cost_f(T) for the 2^k subsets T of the k
//     nodes is a min-reduction, and the slot limits couple the functions only through the k slot counters,
//     so a dynamic programme over functions with (slots+1)^k states finds the best of ~10^7 joint
//     reassignments (k = 3, F = 10) in a few thousand instructions.
// The DP is run on perturbed costs (noise proportional to an annealed temperature) and its result is always
// accepted; a chain records the best placement it visits (by g, only when no node is overloaded at the
// current prices).  The caller prices the chains' records exactly (route_lp.cu) and returns the best.
// Chains start from randomised roundings of the LP relaxation's c-bar (pdhg_mf.cu) or from seed placements.
// No atomics on floating point, fixed evaluation orders, counter-based random numbers: runs are
// bit-reproducible for a given rng seed.
#include "common.cuh"

namespace neptune {

struct LnsArgs {
  int B, N, F, kind, chains, rounds, k, smax, wpb, S, maxslots;
  double alpha, noise_coef;
  uint64_t rng;
  const double *d, *w, *r, *m, *Mj, *Kj, *maxd;
  const double* guide;     // [B][F][N] c-bar of the relaxation, or null
  const double* lam0;      // [B][N] CPU-row duals of the relaxation (delay units), or null
  const uint8_t* seeds;    // [B][S][F][N] or null
  uint8_t* out_c;          // [B][2*chains][F][N]: per chain one record by the upper bound, one by the lower bound
  double* out_g;           // [B][2*chains] [0,chains): whole-flow objective of the record (upper bound of its true objective);
                           //               [chains,2*chains): priced objective (lower bound); +inf: none
  double* out_lb;          // [B][2*chains] the other end of each record's bracket; may be null
  int32_t* out_round;      // [B][2*chains] round of the record
};

constexpr double kLnsBig = 1e9;                 // priced delay of "no pod"
constexpr double kFxScale = 1073741824.0;       // 2^30: CPU loads are accumulated in fixed point (integer adds commute)
constexpr int kLnsMaxK = 4;
// chains per block: two blocks per SM, either 8 warps at 128 registers or 12 warps at 80 registers.  A chain-round is
// 3-10 % slower at 80 registers, so 8 is used while all blocks are resident at once (2 * 148); beyond that the 24
// resident warps per SM of the 12-warp build win (the kernel is latency-bound: 30 % issue-active at 16 warps per SM).
constexpr int kLnsWarps = 8, kLnsWarpsWide = 12;

__device__ __forceinline__ uint64_t lns_mix(uint64_t z) {        // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double lns_u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ uint64_t lns_next(uint64_t& s) { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return s * 2685821657736338717ull; }

// ---- shared-memory layout -----------------------------------------------------------------------------------
// block: dT[N][N] (delay, transposed: the hot loops fix a destination j and run the lanes over sources i),
//        w[F][N], r[F][N], K[N], slots[N]
// warp : lam[N] | loadfx[N] | bestv[F*N] | scratch (costT | cur | nxt  /  th  /  exception list) | ints | bytes
__host__ __device__ inline size_t lns_block_shared(int N, int F) {
  return ((size_t)N * N + 2 * (size_t)F * N + (size_t)N) * 8 + (size_t)N * 4 + 64;
}
__host__ __device__ inline size_t lns_scratch_bytes(int N, int F, int k, int smax, int maxslots) {
  const size_t fn = (size_t)F * N, nT = (size_t)1 << k, tn = (size_t)(maxslots < F ? maxslots : F) * N;
  size_t dp = (size_t)F * nT * 8 + 2 * (size_t)smax * 8;      // costT, cur, nxt
  size_t th = tn * 4 + tn * 2;                                // thresholds (float) + their scan list (u16) of one node's sources
  if (th < fn * 2) th = fn * 2;                               // exception list (u16)
  return ((dp > th ? dp : th) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t lns_warp_shared(int N, int F, int k, int smax, int maxslots) {
  const size_t fn = (size_t)F * N;
  size_t b = 0;
  b += (size_t)N * 8;               // lam
  b += (size_t)N * 8;               // loadfx
  b += (fn * 4 + 15) & ~(size_t)15; // bestv (float: it ranks candidates, the records are priced exactly afterwards)
  b += lns_scratch_bytes(N, F, k, smax, maxslots);
  b += ((size_t)F + 64 + (size_t)F) * 4;   // npods, misc (J, rad, mul, incT), fl
  b += 3 * fn;                      // asg, c, podlist
  b += (size_t)F * smax;            // choice
  b += (size_t)smax + 2 * (size_t)F + 16;  // supp table, oldT, newT
  return (b + 15) & ~(size_t)15;
}

// reduce NT per-lane accumulators over the warp: afterwards lane L holds the warp total of accumulator
// t(L) = the top log2(NT) bits of L's 5-bit index read as a number (NT = 8: L >> 2; NT = 16: L >> 1)
template <int NT>
__device__ __forceinline__ double lns_multi_reduce(double (&acc)[NT], int lane) {
  int off = 16;
#pragma unroll
  for (int n = NT; n > 1; n >>= 1, off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int t = 0; t < n / 2; ++t) {
      const double keep = hi ? acc[t + n / 2] : acc[t];
      const double send = hi ? acc[t] : acc[t + n / 2];
      acc[t] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  double v = acc[0];
  for (; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

template <int KK, int WPB>
__global__ void __launch_bounds__(WPB * 32, 2) k_lns(LnsArgs a) {
  constexpr int NT = 1 << KK;
  const int N = a.N, F = a.F, b = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int chain = blockIdx.x * a.wpb + wid;
  const int fn = F * N, smax = a.smax;
  extern __shared__ __align__(16) unsigned char smem[];
  double* s_dT = (double*)smem;
  double* s_w = s_dT + (size_t)N * N;
  double* s_r = s_w + fn;
  double* s_K = s_r + fn;
  int* s_slots = (int*)(s_K + N);
  __shared__ double s_scal[4];          // noise0, u (utilisation weight in delay units), a_d
  {
    const double* d = a.d + (int64_t)b * N * N;
    const double* w = a.w + (int64_t)b * fn;
    const double* r = a.r + (int64_t)b * fn;
    for (int q = threadIdx.x; q < N * N; q += blockDim.x) { const int i = q / N, j = q - i * N; s_dT[j * N + i] = d[q]; }
    for (int q = threadIdx.x; q < fn; q += blockDim.x) { s_w[q] = w[q]; s_r[q] = r[q]; }
    const double m0 = a.m[(int64_t)b * F];
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      s_K[j] = a.Kj[(int64_t)b * N + j];
      int sl = m0 > 0.0 ? (int)floor(a.Mj[(int64_t)b * N + j] / m0 + 1e-9) : F;
      s_slots[j] = sl > a.maxslots ? a.maxslots : (sl < 0 ? 0 : sl);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // instance scalars: mean workload, mean off-diagonal delay (noise scale), objective weights
    double sw = 0.0, sd = 0.0, wm = 0.0;
    for (int q = lane; q < fn; q += 32) sw += s_w[q];
    for (int q = lane; q < N * N; q += 32) sd += s_dT[q];
    for (int fi = lane; fi < fn; fi += 32) {
      const int f = fi / N, i = fi - f * N;
      const double md = a.maxd ? a.maxd[(int64_t)b * F + f] : INFINITY;
      double best = -INFINITY;
      for (int j = 0; j < N; ++j) { const double v = s_dT[j * N + i]; if (v <= md && v > best) best = v; }
      wm += s_w[fi] * best;
    }
    sw = warp_sum(sw); sd = warp_sum(sd); wm = warp_sum(wm);
    if (lane == 0) {
      double a_d = 1.0, a_u = 0.0;
      if (a.kind == NEPTUNE_KIND_MIN_DELAY_UTIL) { a_u = a.alpha / (double)N; a_d = (sw != 0.0 && wm != 0.0) ? (1.0 - a.alpha) / wm : 0.0; }
      const double dmean = N > 1 ? sd / ((double)N * (N - 1)) : 1.0;
      s_scal[0] = a.noise_coef * (sw / fn) * dmean;
      s_scal[1] = a_d > 0.0 ? a_u / a_d : 0.0;
      s_scal[2] = a_d;
    }
  }
  __syncthreads();
  if (chain >= a.chains) return;                 // (no block-wide barrier below this line)
  // chains anneal from different start temperatures (x0.6, 0.9, 1.2, 1.5 by chain index): no single temperature
  // suits every instance, and a mixed population finds the optimum of more of them than any uniform one
  const double noise0 = s_scal[0] * (0.6 + 0.3 * (double)(chain & 3)), u = s_scal[1], a_d = s_scal[2];

  // ---- per-warp state ------------------------------------------------------------------------------------
  unsigned char* wp = smem + ((lns_block_shared(N, F) + 15) & ~(size_t)15) + (size_t)wid * lns_warp_shared(N, F, KK, smax, a.maxslots);
  double* lam = (double*)wp; wp += (size_t)N * 8;
  unsigned long long* loadfx = (unsigned long long*)wp; wp += (size_t)N * 8;
  float* bestv = (float*)wp; wp += ((size_t)fn * 4 + 15) & ~(size_t)15;
  unsigned char* scratch = wp; wp += lns_scratch_bytes(N, F, KK, smax, a.maxslots);
  double* costT = (double*)scratch;                       // [F][NT]
  double* cur = costT + (size_t)F * NT;
  double* nxt = cur + smax;
  float* th = (float*)scratch;                            // aliases the DP arrays (never live together)
  unsigned short* thl = (unsigned short*)(scratch + (size_t)(a.maxslots < F ? a.maxslots : F) * N * 4);   // sources whose threshold needs a pod scan
  unsigned short* exl = (unsigned short*)scratch;         // exception list, ditto
  int* npods = (int*)wp; wp += (size_t)F * 4;
  int* misc = (int*)wp; wp += 64 * 4;
  int* fl = (int*)wp; wp += (size_t)F * 4;
  uint8_t* asg = wp; wp += fn;
  uint8_t* c = wp; wp += fn;
  uint8_t* podlist = wp; wp += fn;
  uint8_t* choice = wp; wp += (size_t)F * smax;
  uint8_t* suppT = wp; wp += smax;
  uint8_t* oldT = wp; wp += F;
  uint8_t* newT = wp;
  int* Jn = misc; int* rad = misc + 4; int* mul = misc + 8; int* incT = misc + 12;      // incT[16]

  uint64_t rs = a.rng ^ (0x9E3779B97F4A7C15ull * (uint64_t)(chain + 1)) ^ (0xD1B54A32D192ED03ull * (uint64_t)(b + 1));
  lns_next(rs); lns_next(rs);

  // ---- start placement -----------------------------------------------------------------------------------
  for (int q = lane; q < fn; q += 32) c[q] = 0;
  for (int j = lane; j < N; j += 32) lam[j] = a.lam0 ? fmax(a.lam0[(int64_t)b * N + j], 0.0) : 0.0;
  __syncwarp();
  const bool from_seed = a.seeds && (!a.guide || chain < a.S);
  if (from_seed) {
    const uint8_t* seed = a.seeds + ((int64_t)b * a.S + (chain % a.S)) * fn;
    for (int q = lane; q < fn; q += 32) c[q] = seed[q] ? 1 : 0;
    __syncwarp();
    // seeds may exceed the slot count of a node (never for EFTTC output): drop the highest-numbered extras
    for (int j = lane; j < N; j += 32) { int cnt = 0; for (int f = 0; f < F; ++f) { if (c[f * N + j]) { if (cnt >= s_slots[j]) c[f * N + j] = 0; else ++cnt; } } }
  } else {
    // randomised rounding: per node the `slots` functions with the largest perturbed c-bar (or random without guide)
    const double amp = a.guide ? 0.02 + 0.12 * (double)((chain / 2) % 4) : 1.0;
    const double* g = a.guide ? a.guide + (int64_t)b * fn : nullptr;
    for (int j = lane; j < N; j += 32) {
      for (int cnt = 0; cnt < s_slots[j]; ++cnt) {
        int bf = -1; double bg = -INFINITY;
        for (int f = 0; f < F; ++f) {
          if (c[f * N + j]) continue;
          const double v = (g ? g[f * N + j] : 0.0) + amp * lns_u01(lns_mix(rs ^ ((uint64_t)(f * N + j) << 20)));
          if (v > bg) { bg = v; bf = f; }
        }
        if (bf < 0) break;
        c[bf * N + j] = 1;
      }
    }
  }
  __syncwarp();
  // every function needs a pod: put a missing one on the node where it has the most workload and a free slot,
  // else replace the pod of the function with the most pods there (lane 0; rare)
  if (lane == 0) {
    for (int f = 0; f < F; ++f) {
      int any = 0; for (int j = 0; j < N; ++j) any |= c[f * N + j];
      if (any) continue;
      int bj = -1; double bw = -1.0;
      for (int j = 0; j < N; ++j) { if (s_slots[j] <= 0) continue; if (s_w[f * N + j] > bw) { bw = s_w[f * N + j]; bj = j; } }
      if (bj < 0) continue;
      int cnt = 0; for (int f2 = 0; f2 < F; ++f2) cnt += c[f2 * N + bj];
      if (cnt >= s_slots[bj]) {
        int vf = -1, vp = 1;
        for (int f2 = 0; f2 < F; ++f2) { if (!c[f2 * N + bj]) continue; int p2 = 0; for (int j = 0; j < N; ++j) p2 += c[f2 * N + j]; if (p2 > vp) { vp = p2; vf = f2; } }
        if (vf < 0) continue;
        c[vf * N + bj] = 0;
      }
      c[f * N + bj] = 1;
    }
  }
  __syncwarp();

  // pod list of one function, ascending (ballot compaction; warp-uniform call)
  auto rebuild_pods = [&](int f) {
    int n = 0;
    for (int jb = 0; jb < N; jb += 32) {
      const int j = jb + lane;
      const bool on = j < N && c[f * N + j];
      const unsigned bal = __ballot_sync(0xffffffffu, on);
      if (on) podlist[f * N + n + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)j;
      n += __popc(bal);
    }
    if (lane == 0) npods[f] = n;
    __syncwarp();
  };
  // nearest priced pod of every source of f (lanes over sources, pods in a warp-uniform loop); `skip` < 0: none excluded
  auto scan_function = [&](int f) {
    const int np = npods[f];
    const uint8_t* pl = podlist + f * N;
    const double* rf = s_r + f * N;
    for (int ib = 0; ib < N; ib += 32) {
      const int i = ib + lane;
      double best = kLnsBig; int bj = 255;
      if (i < N) {
#pragma unroll 2
        for (int q = 0; q < np; ++q) {
          const int j = pl[q];
          const double v = s_dT[j * N + i] + lam[j] * rf[j];
          if (v < best) { best = v; bj = j; }
        }
        bestv[f * N + i] = (float)best; asg[f * N + i] = (uint8_t)bj;
      }
    }
  };
  for (int f = 0; f < F; ++f) rebuild_pods(f);
  for (int f = 0; f < F; ++f) scan_function(f);
  __syncwarp();

  uint64_t dirtyF = ~0ull;             // functions whose routing changed since the node prices were last settled
  double bestu = INFINITY, bestg = -INFINITY; int best_round = -1;
  double bestg2 = INFINITY, bestu2 = INFINITY; int best_round2 = -1;
  // records of an instance: [0, chains) by the whole-flow (upper-bound) objective, [chains, 2 chains) by the priced one
  uint8_t* outc = a.out_c + ((int64_t)b * 2 * a.chains + chain) * fn;
  uint8_t* outc2 = a.out_c + ((int64_t)b * 2 * a.chains + a.chains + chain) * fn;
  const int anneal_rounds = a.rounds - a.rounds / 8;

  #pragma unroll 1
  for (int round = 0; round <= a.rounds; ++round) {
    // ---- neighbourhood: a random node and k-1 nodes near it (tournaments on d), or k random nodes ----------
    int kk = KK;
    uint64_t jm0 = 0, jm1 = 0;
    {
      int J[KK];
      const uint64_t h0 = lns_next(rs);
      J[0] = (int)(((h0 & 0xffffffffull) * (uint64_t)N) >> 32);
      const bool local = ((h0 >> 40) & 3) != 0;
      int have = 1;
      #pragma unroll 1
      for (int tries = 0; have < KK && tries < 16 * KK; ++tries) {
        int bj = -1; double bd = INFINITY;
        const uint64_t h = lns_next(rs);
        const int tsize = local ? 1 + (int)((h >> 50) & 3) : 1;
        uint64_t hh = h;
        #pragma unroll 1
        for (int t = 0; t < tsize; ++t) {
          hh = lns_mix(hh + t);
          const int j = (int)(((hh & 0xffffffffull) * (uint64_t)N) >> 32);
          bool dup = false;
#pragma unroll
          for (int q = 0; q < KK; ++q) dup = dup || (q < have && J[q] == j);
          if (dup) continue;
          const int pick = (int)((((hh >> 32) & 0xffffull) * (uint64_t)have) >> 16);
          int jp = J[0];
#pragma unroll
          for (int q = 1; q < KK; ++q) if (q == pick) jp = J[q];
          const double dv = s_dT[j * N + jp];
          if (dv < bd) { bd = dv; bj = j; }
        }
        if (bj >= 0) {
#pragma unroll
          for (int q = 1; q < KK; ++q) if (q == have) J[q] = bj;
          ++have;
        }
      }
      kk = have;
      // state space of the slot counters must fit: drop trailing nodes otherwise
      int S_ = 1, fit = 0;
#pragma unroll
      for (int q = 0; q < KK; ++q) {
        if (q < kk && fit == q) { const int rd = s_slots[J[q]] + 1; if ((int64_t)S_ * rd <= smax) { S_ *= rd; ++fit; } }
      }
      kk = fit;
      if (lane == 0) {
        int mu_ = 1;
#pragma unroll
        for (int q = 0; q < KK; ++q) if (q < kk) { Jn[q] = J[q]; rad[q] = s_slots[J[q]] + 1; mul[q] = mu_; mu_ *= rad[q]; }
        misc[32] = mu_;                                    // number of states
      }
#pragma unroll
      for (int q = 0; q < KK; ++q) if (q < kk) { if (J[q] < 64) jm0 |= 1ull << J[q]; else jm1 |= 1ull << (J[q] - 64); }
    }
    __syncwarp();
    auto inJ = [&](int j) -> bool { return j < 64 ? (jm0 >> j) & 1 : (j < 128 ? (jm1 >> (j - 64)) & 1 : false); };

    // ---- CPU loads at the current prices; dual ascent on the prices of overloaded / priced nodes -----------------
    // A node is (re)priced when it is overloaded or carries a price.  Price of node jj = the threshold at which enough of its flows leave:  a source (f,i) uses jj
    // while  lam_jj < th = (best priced alternative - d[i,jj]) / r[f,jj];  flows leave in ascending th until the
    // rest fits K_jj; ties with the alternative (th = 0) leave first, at a price of 1e-9.
    bool overloaded = false, settled = false;
    double ucost = 0.0;
    // the state of a node depends only on the functions with a pod on it: a price is re-derived when the node is
    // overloaded, or when one of its functions changed pods (last apply) or was re-routed by another node's new price
    uint64_t dirty_now = dirtyF;
    const int max_pass = N <= 32 ? 12 : 4;              // small instances with most CPU rows binding need more sweeps of the ascent
    #pragma unroll 1
    for (int pass = 0; pass < max_pass; ++pass) {
      #pragma unroll 1
      for (int j = lane; j < N; j += 32) loadfx[j] = 0ull;
      __syncwarp();
      double up = 0.0;
      #pragma unroll 1
      for (int fi = lane; fi < fn; fi += 32) {
        const int bj = asg[fi];
        if (bj == 255) continue;
        const int f = fi / N, i = fi - f * N;
        const double wv = s_w[fi];
        const double ld = wv * s_r[f * N + bj];
        if (ld > 0.0) atomicAdd(&loadfx[bj], (unsigned long long)(fmin(ld, 1e9) * kFxScale + 0.5));
        if (wv > 0.0) up += wv * s_dT[bj * N + i];
      }
      ucost = warp_sum(up);                              // delay of the whole-flow routing at the current prices
      __syncwarp();
      bool changed = false; overloaded = false;
      #pragma unroll 1
      for (int jb = 0; jb < N; jb += 32) {
        const int j = jb + lane;
        bool need = false, over = false;
        if (j < N) {
          const double ld = (double)loadfx[j] * (1.0 / kFxScale);
          over = ld > s_K[j] + 1e-7;
          uint64_t on = 0;
          if (F <= 64) { for (int f = 0; f < F; ++f) if (c[f * N + j]) on |= 1ull << f; } else on = ~0ull;
          need = over || (lam[j] > 0.0 && (on & dirty_now) != 0);
        }
        unsigned todo = __ballot_sync(0xffffffffu, need);
        overloaded = overloaded || __any_sync(0xffffffffu, over);
        if (pass == max_pass - 1) continue;                // last pass only evaluates
        #pragma unroll 1
        while (todo) {
          const int jj = jb + (__ffs(todo) - 1); todo &= todo - 1;
          int nf = 0;
          #pragma unroll 1
          for (int fb = 0; fb < F; fb += 32) {
            const int f = fb + lane;
            const bool on = f < F && c[f * N + jj];
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            if (on) fl[nf + __popc(bal & ((1u << lane) - 1u))] = f;
            nf += __popc(bal);
          }
          __syncwarp();
          const int M = nf * N;
          // thresholds: a source whose nearest priced pod is not jj has that pod as its alternative; the sources
          // currently ON jj need their second-best pod -- a pod scan, done for the compacted list of them only
          double tot = 0.0; int nscan = 0;
          #pragma unroll 1
          for (int tb = 0; tb < M; tb += 32) {
            const int t = tb + lane;
            bool scan = false;
            if (t < M) {
              const int z = t / N, i = t - z * N, f = fl[z];
              const double rfj = s_r[f * N + jj], av = s_w[f * N + i] * rfj;
              float tv = -1.0f;                              // never uses jj
              if (av > 0.0) {
                if (asg[f * N + i] != jj) {
                  const double tq = ((double)bestv[f * N + i] - s_dT[jj * N + i]) / rfj;
                  if (tq >= 0.0) { tv = bestv[f * N + i] >= 0.5f * (float)kLnsBig ? INFINITY : (float)fmin(tq, 1e30); tot += av; }
                } else scan = true;
              }
              th[t] = tv;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, scan);
            if (scan) thl[nscan + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)t;
            nscan += __popc(bal);
          }
          __syncwarp();
          #pragma unroll 1
          for (int e = lane; e < nscan; e += 32) {
            const int t = thl[e], z = t / N, i = t - z * N, f = fl[z];
            const double* rf = s_r + f * N; const uint8_t* pl = podlist + f * N; const int np = npods[f];
            const double rfj = rf[jj];
            double alt = kLnsBig;
            #pragma unroll 1
            for (int q = 0; q < np; ++q) { const int j2 = pl[q]; if (j2 == jj) continue; const double v = s_dT[j2 * N + i] + lam[j2] * rf[j2]; if (v < alt) alt = v; }
            const double tq = (alt - s_dT[jj * N + i]) / rfj;
            if (tq >= 0.0) { th[t] = alt >= 0.5 * kLnsBig ? INFINITY : (float)fmin(tq, 1e30); tot += s_w[f * N + i] * rfj; }   // no other pod: cannot leave
          }
          tot = warp_sum(tot);
          __syncwarp();
          double nl = 0.0;
          if (tot > s_K[jj] + 1e-9) {
            float last = -1.0f;
            #pragma unroll 1
            for (int it = 0; it < M && tot > s_K[jj] + 1e-9; ++it) {
              unsigned mn = 0xffffffffu;
              #pragma unroll 1
              for (int t = lane; t < M; t += 32) { const float tv = th[t]; if (tv > last) mn = min(mn, __float_as_uint(tv)); }
              mn = __reduce_min_sync(0xffffffffu, mn);
              if (mn >= 0x7f800000u) break;                 // nothing movable left: the node stays overloaded (never recorded)
              last = __uint_as_float(mn);
              double rem = 0.0;
              #pragma unroll 1
              for (int t = lane; t < M; t += 32) if (th[t] == last) { const int z = t / N; rem += s_w[fl[z] * N + (t - z * N)] * s_r[fl[z] * N + jj]; }
              tot -= warp_sum(rem);
            }
            nl = (double)fmaxf(last, 0.0f) * (1.0 + 2e-6) + 1e-9;
          }
          const bool ch = fabs(nl - lam[jj]) > 1e-10 * (1.0 + nl);
          __syncwarp();
          if (ch) {
            changed = true;
            if (lane == 0) lam[jj] = nl;
            __syncwarp();
            #pragma unroll 1
            for (int z = 0; z < nf; ++z) { scan_function(fl[z]); dirty_now |= F <= 64 ? 1ull << fl[z] : ~0ull; }   // their sources see a new price on jj
            __syncwarp();
          }
        }
      }
      if (!changed) { settled = true; break; }
    }
    dirtyF = settled ? 0ull : dirty_now;
    // ---- record ---------------------------------------------------------------------------------------------------
    // With no node overloaded, routing every source WHOLE to its nearest priced pod is a feasible point of the MIP
    // (pods that end up unused are closed, the others serve at least one whole source: C1b holds), so its delay
    // `ucost` is an upper bound of the placement's true objective and the priced cost g a lower bound; they differ
    // by sum_j lam_j (K_j - load_j).  Chains record by the upper bound: it can only be improved by the exact routing
    // LP afterwards (split flows), never be a mirage of loose prices.
    {
      double gp = 0.0;
      #pragma unroll 1
      for (int fi = lane; fi < fn; fi += 32) { const double wv = s_w[fi]; if (wv > 0.0) gp += wv * (double)bestv[fi]; }
      #pragma unroll 1
      for (int j = lane; j < N; j += 32) gp -= lam[j] * s_K[j];
      double g = warp_sum(gp), uval = ucost;
      if (u != 0.0) {
        int act = 0;
        #pragma unroll 1
        for (int j = lane; j < N; j += 32) { int any = 0; for (int f = 0; f < F; ++f) any |= c[f * N + j]; act += any; }
        act = __reduce_add_sync(0xffffffffu, act);
        g += u * (double)act; uval += u * (double)act;
      }
      bool unserved = false;
      #pragma unroll 1
      for (int fi = lane; fi < fn; fi += 32) unserved = unserved || asg[fi] == 255;
      unserved = __any_sync(0xffffffffu, unserved);
      if (!overloaded && !unserved && uval < bestu - 1e-9 * (1.0 + fabs(uval))) {
        bestu = uval; bestg = g; best_round = round;
        #pragma unroll 1
        for (int q = lane; q < fn; q += 32) outc[q] = c[q];
      }
      // second record, by the lower bound: where the optimal routing splits flows, g is the LP value and the whole-flow
      // value overstates it by sum_j lam_j (K_j - load_j) -- 15 % on a 20-node instance whose optimum splits a heavy flow.
      // Only with settled prices; both records are priced exactly by the caller.
      if (!overloaded && !unserved && settled && g < bestg2 - 1e-9 * (1.0 + fabs(g))) {
        bestg2 = g; bestu2 = uval; best_round2 = round;
        #pragma unroll 1
        for (int q = lane; q < fn; q += 32) outc2[q] = c[q];
      }
    }
    if (round == a.rounds) break;
    if (kk == 0) continue;

    // ---- sources served from the k nodes: their nearest pod OUTSIDE the k nodes (compacted list, then a scan) ------
    int nex = 0;
    #pragma unroll 1
    for (int fb = 0; fb < fn; fb += 32) {
      const int fi = fb + lane;
      const bool ex = fi < fn && asg[fi] != 255 && inJ(asg[fi]);
      const unsigned bal = __ballot_sync(0xffffffffu, ex);
      if (ex) exl[nex + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)fi;
      nex += __popc(bal);
    }
    __syncwarp();
    #pragma unroll 1
    for (int e = lane; e < nex; e += 32) {
      const int fi = exl[e], f = fi / N, i = fi - f * N;
      const uint8_t* pl = podlist + f * N; const double* rf = s_r + f * N; const int np = npods[f];
      double best = kLnsBig; int bj = 255;
      #pragma unroll 1
      for (int q = 0; q < np; ++q) {
        const int j = pl[q];
        if (inJ(j)) continue;
        const double v = s_dT[j * N + i] + lam[j] * rf[j];
        if (v < best) { best = v; bj = j; }
      }
      bestv[fi] = (float)best; asg[fi] = (uint8_t)bj;      // (bestv, asg) now describe "the k nodes closed" for every source
    }
    __syncwarp();

    // ---- cost of every subset T of the k nodes, per function (lanes over sources, NT accumulators per lane) ------
    const double temp = round < anneal_rounds ? noise0 * (1.0 - (double)round / (double)anneal_rounds) : 0.0;
    const uint64_t rround = lns_next(rs);
    const int Tvalid = (1 << kk) - 1;
    int Jr[KK];
#pragma unroll
    for (int q = 0; q < KK; ++q) Jr[q] = q < kk ? Jn[q] : 0;
    #pragma unroll 1
    for (int f = lane; f < F; f += 32) {
      int ot = 0;
#pragma unroll
      for (int q = 0; q < KK; ++q) if (q < kk && c[f * N + Jr[q]]) ot |= 1 << q;
      oldT[f] = (uint8_t)ot;
    }
    __syncwarp();
    #pragma unroll 1
    for (int f = 0; f < F; ++f) {
      double pj[KK];
#pragma unroll
      for (int q = 0; q < KK; ++q) pj[q] = q < kk ? lam[Jr[q]] * s_r[f * N + Jr[q]] : kLnsBig;
      double acc[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[t] = 0.0;
      #pragma unroll 1
      for (int ib = 0; ib < N; ib += 32) {
        const int i = ib + lane;
        if (i >= N) continue;
        const double wv = s_w[f * N + i];
        if (wv == 0.0) continue;
        double val[NT];
        val[0] = (double)bestv[f * N + i];
        double vq[KK];
#pragma unroll
        for (int q = 0; q < KK; ++q) vq[q] = s_dT[Jr[q] * N + i] + pj[q];
#pragma unroll
        for (int t = 1; t < NT; ++t) val[t] = fmin(val[t & (t - 1)], vq[__ffs(t) - 1]);
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t] += wv * val[t];
      }
      const double tot = lns_multi_reduce<NT>(acc, lane);
      const int T = lane >> (5 - KK);
      if ((lane & ((32 >> KK) - 1)) == 0) {
        double v = tot;
        if (temp > 0.0) v += temp * lns_u01(lns_mix(rround ^ ((uint64_t)(f * 64 + T) << 8))) * (double)__popc(T);
        const bool outside = npods[f] - __popc((unsigned)oldT[f]) > 0;
        if ((T & ~Tvalid) != 0 || (T == 0 && !outside)) v = INFINITY;     // dropped node / the function would lose its last pod
        costT[f * NT + T] = v;
      }
    }
    __syncwarp();

    // ---- DP over functions, state = slot counters of the k nodes -----------------------------------------------
    const int S_ = misc[32];
    #pragma unroll 1
    for (int T = lane; T < NT; T += 32) { int s = 0; for (int q = 0; q < kk; ++q) if ((T >> q) & 1) s += mul[q]; incT[T] = s; }
    #pragma unroll 1
    for (int s = lane; s < S_; s += 32) {
      int supp = 0;
      #pragma unroll 1
      for (int q = 0; q < kk; ++q) if ((s / mul[q]) % rad[q] > 0) supp |= 1 << q;
      suppT[s] = (uint8_t)supp;
      cur[s] = s == 0 ? 0.0 : INFINITY;
    }
    __syncwarp();
    double* pc = cur; double* pn = nxt;
    #pragma unroll 1
    for (int f = 0; f < F; ++f) {
      const double* ct = costT + f * NT;
      #pragma unroll 1
      for (int s2 = lane; s2 < S_; s2 += 32) {
        const int supp = suppT[s2];
        double best = INFINITY; int bt = 0;
        int Ts = supp;
        #pragma unroll 1
        while (true) {
          const double v = pc[s2 - incT[Ts]] + ct[Ts];
          if (v < best) { best = v; bt = Ts; }
          if (Ts == 0) break;
          Ts = (Ts - 1) & supp;
        }
        pn[s2] = best; choice[f * smax + s2] = (uint8_t)bt;
      }
      __syncwarp();
      double* t_ = pc; pc = pn; pn = t_;
    }
    // best final state (utilisation: an empty node costs nothing, a used one u)
    double bv = INFINITY; int bs = 0x7fffffff;
    #pragma unroll 1
    for (int s = lane; s < S_; s += 32) {
      double v = pc[s];
      if (u != 0.0) v += u * (double)__popc((unsigned)suppT[s]);
      if (v < bv) { bv = v; bs = s; }
    }
    #pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
      const double v2 = __shfl_xor_sync(0xffffffffu, bv, o); const int s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      if (v2 < bv || (v2 == bv && s2 < bs)) { bv = v2; bs = s2; }
    }
    if (lane == 0) {
      if (bv < INFINITY) {
        int s = bs;
        #pragma unroll 1
        for (int f = F - 1; f >= 0; --f) { const int Tc = choice[f * smax + s]; newT[f] = (uint8_t)Tc; s -= incT[Tc]; }
      } else {
        #pragma unroll 1
        for (int f = 0; f < F; ++f) newT[f] = oldT[f];
      }
    }
    __syncwarp();
    // ---- apply: pods of the k nodes, then the nearest pod of every source of the functions that use them ---------
    #pragma unroll 1
    for (int f = 0; f < F; ++f) {
      const int nt_ = newT[f], ot_ = oldT[f];
      if (nt_ != ot_) {
        dirtyF |= F <= 64 ? 1ull << f : ~0ull;
        if (lane < kk) c[f * N + Jn[lane]] = (nt_ >> lane) & 1;
        __syncwarp();
        rebuild_pods(f);
      }
      if (nt_ == 0) continue;
      double pj[KK];
#pragma unroll
      for (int q = 0; q < KK; ++q) pj[q] = q < kk ? lam[Jr[q]] * s_r[f * N + Jr[q]] : kLnsBig;
      #pragma unroll 1
      for (int ib = 0; ib < N; ib += 32) {
        const int i = ib + lane;
        if (i >= N) continue;
        double best = (double)bestv[f * N + i]; int bj = asg[f * N + i];
#pragma unroll
        for (int q = 0; q < KK; ++q) {
          if (q < kk && ((nt_ >> q) & 1)) { const double v = s_dT[Jr[q] * N + i] + pj[q]; if (v < best || (v == best && Jr[q] < bj)) { best = v; bj = Jr[q]; } }
        }
        bestv[f * N + i] = (float)best; asg[f * N + i] = (uint8_t)bj;
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    const int64_t o = (int64_t)b * 2 * a.chains + chain;
    a.out_g[o] = bestu < INFINITY ? a_d * bestu : INFINITY;
    a.out_g[o + a.chains] = bestg2 < INFINITY ? a_d * bestg2 : INFINITY;
    if (a.out_lb) { a.out_lb[o] = bestu < INFINITY ? a_d * bestg : -INFINITY; a.out_lb[o + a.chains] = bestg2 < INFINITY ? a_d * bestu2 : INFINITY; }
    a.out_round[o] = best_round; a.out_round[o + a.chains] = best_round2;
  }
  if (best_round < 0) for (int q = lane; q < fn; q += 32) outc[q] = c[q];
  if (best_round2 < 0) for (int q = lane; q < fn; q += 32) outc2[q] = c[q];
}

}  // namespace neptune

using namespace neptune;

static int g_lns_block_mode = 0;      // 0: by the number of blocks, 1: always 8 chains per block, 2: always 12
extern "C" int neptune_lns_block_mode(int mode) {
  if (mode < 0 || mode > 2) return NEPTUNE_E_ARG;
  g_lns_block_mode = mode;
  return 0;
}

extern "C" int neptune_lns_search(int B, int N, int F, int kind, double alpha, int chains, int rounds, int k,
                                  double noise_coef, uint64_t rng_seed, const double* d, const double* w,
                                  const double* r, const double* m, const double* Mj, const double* Kj,
                                  const double* maxd, const double* guide, const double* lam0, int S,
                                  const uint8_t* seeds, int max_slots, uint8_t* out_c, double* out_g, double* out_lb,
                                  int32_t* out_round, void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || chains <= 0 || rounds < 0 || k < 2 || k > kLnsMaxK || !d || !w || !r || !m || !Mj ||
      !Kj || !out_c || !out_g || !out_round)
    return NEPTUNE_E_ARG;
  if (kind != NEPTUNE_KIND_MIN_DELAY && kind != NEPTUNE_KIND_MIN_DELAY_UTIL) return NEPTUNE_E_ARG;
  if (seeds && S <= 0) return NEPTUNE_E_ARG;
  if (N > 128 || (int64_t)F * N > 4096) return NEPTUNE_E_SIZE;       // the chain state lives in shared memory
  LnsArgs a{};
  a.B = B; a.N = N; a.F = F; a.kind = kind; a.chains = chains; a.rounds = rounds; a.k = k; a.S = seeds ? S : 0;
  a.alpha = alpha; a.noise_coef = noise_coef; a.rng = rng_seed ? rng_seed : 0x1234567ull;
  a.d = d; a.w = w; a.r = r; a.m = m; a.Mj = Mj; a.Kj = Kj; a.maxd = maxd; a.guide = guide; a.lam0 = lam0; a.seeds = seeds;
  a.out_c = out_c; a.out_g = out_g; a.out_lb = out_lb; a.out_round = out_round;
  // states of the slot counters: (slots+1)^k for the common 3-slot nodes, capped by shared memory
  int smax = 1; for (int q = 0; q < k; ++q) smax *= 4;
  const size_t blk = (lns_block_shared(N, F) + 15) & ~(size_t)15;
  a.maxslots = (max_slots > 0 && max_slots < F) ? max_slots : F;       // bound on floor(Mj / m) the caller vouches for (sizes the scratch)
  const size_t per = lns_warp_shared(N, F, k, smax, a.maxslots);
  const int wide_mode = g_lns_block_mode;
  const size_t budget = 112 * 1024;                                     // two blocks per SM (228 KB, 1 KB reserved per block)
  int wpb = kLnsWarps;
  const int64_t blocks8 = (int64_t)B * ((chains + kLnsWarps - 1) / kLnsWarps);
  const bool wide = blocks8 > 2 * kNumSMs && blk + kLnsWarpsWide * per <= budget && !(wide_mode & 1);
  if (wide || (wide_mode & 2)) wpb = kLnsWarpsWide;
  const int cap = wpb;
  while (wpb > 1 && blk + wpb * per > budget) --wpb;
  if (blk + wpb * per > 200 * 1024) return NEPTUNE_E_SIZE;
  if (wpb > chains) wpb = chains;
  a.smax = smax; a.wpb = wpb;
  const size_t sm = blk + wpb * per;
  const dim3 grid((chains + wpb - 1) / wpb, B);
#define NEPTUNE_LNS_LAUNCH(KK_, W_)                                                                                   \
  do {                                                                                                                \
    NEPTUNE_CUDA_OK(cudaFuncSetAttribute(k_lns<KK_, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
    k_lns<KK_, W_><<<grid, wpb * 32, sm, (cudaStream_t)stream>>>(a);                                                  \
  } while (0)
  if (cap == kLnsWarpsWide) {
    if (k == 3) NEPTUNE_LNS_LAUNCH(3, kLnsWarpsWide); else if (k == 4) NEPTUNE_LNS_LAUNCH(4, kLnsWarpsWide); else NEPTUNE_LNS_LAUNCH(2, kLnsWarpsWide);
  } else {
    if (k == 3) NEPTUNE_LNS_LAUNCH(3, kLnsWarps); else if (k == 4) NEPTUNE_LNS_LAUNCH(4, kLnsWarps); else NEPTUNE_LNS_LAUNCH(2, kLnsWarps);
  }
#undef NEPTUNE_LNS_LAUNCH
  NEPTUNE_COUNT(1);
  NEPTUNE_LAUNCH_OK();
  return 0;
}
