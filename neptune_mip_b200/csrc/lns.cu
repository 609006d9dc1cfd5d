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
  int B, N, F, kind, chains, rounds, k, smax, wpb, S;
  double alpha, noise_coef;
  uint64_t rng;
  const double *d, *w, *r, *m, *Mj, *Kj, *maxd;
  const double* guide;     // [B][F][N] c-bar of the relaxation, or null
  const double* lam0;      // [B][N] CPU-row duals of the relaxation (delay units), or null
  const uint8_t* seeds;    // [B][S][F][N] or null
  uint8_t* out_c;          // [B][chains][F][N]
  double* out_g;           // [B][chains] priced objective of the record (+inf: none)
  int32_t* out_round;      // [B][chains] round of the record
};

constexpr double kLnsBig = 1e9;                 // priced delay of "no pod"
constexpr double kFxScale = 1073741824.0;       // 2^30: CPU loads are accumulated in fixed point (integer adds commute)
constexpr int kLnsMaxK = 4;

__device__ __forceinline__ uint64_t lns_mix(uint64_t z) {        // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double lns_u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ uint64_t lns_next(uint64_t& s) { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return s * 2685821657736338717ull; }

struct LnsWarp {            // per-warp shared-memory state
  double *lam, *base, *costT, *cur, *nxt;
  unsigned long long* loadfx;
  float* th;
  uint8_t *asg, *c, *podlist, *choice;
  int *npods, *misc;        // misc: J[4] rad[4] mul[4] incT[16] fl[...]
};

__host__ __device__ inline size_t lns_block_shared(int N, int F) {
  return ((size_t)N * N + 2 * (size_t)F * N + (size_t)N) * 8 + (size_t)N * 4 + 64;
}
__host__ __device__ inline size_t lns_warp_shared(int N, int F, int k, int smax) {
  const size_t fn = (size_t)F * N, nT = (size_t)1 << k;
  size_t b = 0;
  b += (size_t)N * 8;               // lam
  b += (size_t)N * 8;               // loadfx
  b += fn * 8;                      // base
  b += (size_t)F * nT * 8;          // costT
  b += 2 * (size_t)smax * 8;        // cur, nxt
  b += fn * 4;                      // th
  b += ((size_t)F + 64 + (size_t)F) * 4;   // npods, misc (J, rad, mul, incT, fl)
  b += 3 * fn;                      // asg, c, podlist
  b += (size_t)F * smax;            // choice
  return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(256) k_lns(LnsArgs a) {
  const int N = a.N, F = a.F, b = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int chain = blockIdx.x * a.wpb + wid;
  const int fn = F * N, K_ = a.k, nT = 1 << K_, smax = a.smax;
  extern __shared__ __align__(16) unsigned char smem[];
  double* s_d = (double*)smem;
  double* s_w = s_d + (size_t)N * N;
  double* s_r = s_w + fn;
  double* s_K = s_r + fn;
  int* s_slots = (int*)(s_K + N);
  __shared__ double s_scal[4];          // noise0, u (utilisation weight in delay units), a_d
  {
    const double* d = a.d + (int64_t)b * N * N;
    const double* w = a.w + (int64_t)b * fn;
    const double* r = a.r + (int64_t)b * fn;
    for (int q = threadIdx.x; q < N * N; q += blockDim.x) s_d[q] = d[q];
    for (int q = threadIdx.x; q < fn; q += blockDim.x) { s_w[q] = w[q]; s_r[q] = r[q]; }
    const double m0 = a.m[(int64_t)b * F];
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      s_K[j] = a.Kj[(int64_t)b * N + j];
      int sl = m0 > 0.0 ? (int)floor(a.Mj[(int64_t)b * N + j] / m0 + 1e-9) : F;
      s_slots[j] = sl > F ? F : (sl < 0 ? 0 : sl);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // instance scalars: mean workload, mean off-diagonal delay (noise scale), objective weights
    double sw = 0.0, sd = 0.0, wm = 0.0;
    for (int q = lane; q < fn; q += 32) sw += s_w[q];
    for (int q = lane; q < N * N; q += 32) sd += s_d[q];
    for (int fi = lane; fi < fn; fi += 32) {
      const int f = fi / N, i = fi - f * N;
      const double md = a.maxd ? a.maxd[(int64_t)b * F + f] : INFINITY;
      double best = -INFINITY;
      for (int j = 0; j < N; ++j) { const double v = s_d[i * N + j]; if (v <= md && v > best) best = v; }
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
  const double noise0 = s_scal[0], u = s_scal[1], a_d = s_scal[2];

  // ---- per-warp state ------------------------------------------------------------------------------------
  unsigned char* wp = smem + ((lns_block_shared(N, F) + 15) & ~(size_t)15) + (size_t)wid * lns_warp_shared(N, F, K_, smax);
  double* lam = (double*)wp; wp += (size_t)N * 8;
  unsigned long long* loadfx = (unsigned long long*)wp; wp += (size_t)N * 8;
  double* base = (double*)wp; wp += (size_t)fn * 8;
  double* costT = (double*)wp; wp += (size_t)F * nT * 8;
  double* cur = (double*)wp; wp += (size_t)smax * 8;
  double* nxt = (double*)wp; wp += (size_t)smax * 8;
  float* th = (float*)wp; wp += (size_t)fn * 4;
  int* npods = (int*)wp; wp += (size_t)F * 4;
  int* misc = (int*)wp; wp += 64 * 4;
  int* fl = (int*)wp; wp += (size_t)F * 4;
  uint8_t* asg = wp; wp += fn;
  uint8_t* c = wp; wp += fn;
  uint8_t* podlist = wp; wp += fn;
  uint8_t* choice = wp;
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
  auto rebuild_pods = [&]() {
    for (int f = lane; f < F; f += 32) {
      int n = 0;
      for (int j = 0; j < N; ++j) if (c[f * N + j]) podlist[f * N + n++] = (uint8_t)j;
      npods[f] = n;
    }
    __syncwarp();
  };
  rebuild_pods();

  double bestg = INFINITY; int best_round = -1;
  uint8_t* outc = a.out_c + ((int64_t)b * a.chains + chain) * fn;
  const int anneal_rounds = a.rounds - a.rounds / 8;

  for (int round = 0; round <= a.rounds; ++round) {
    // ---- neighbourhood: a random node and k-1 nodes near it (tournaments on d), or k random nodes ----------
    int kk = K_;
    uint64_t jm0 = 0, jm1 = 0;
    {
      int J[kLnsMaxK];
      const uint64_t h0 = lns_next(rs);
      J[0] = (int)(h0 % (uint64_t)N);
      const bool local = ((h0 >> 40) & 3) != 0;
      int have = 1;
      for (int tries = 0; have < K_ && tries < 16 * K_; ++tries) {
        int bj = -1; double bd = INFINITY;
        const uint64_t h = lns_next(rs);
        const int tsize = local ? 1 + (int)((h >> 50) & 3) : 1;
        uint64_t hh = h;
        for (int t = 0; t < tsize; ++t) {
          hh = lns_mix(hh + t);
          const int j = (int)(hh % (uint64_t)N);
          bool dup = false;
          for (int q = 0; q < have; ++q) dup = dup || J[q] == j;
          if (dup) continue;
          const double dv = s_d[J[(hh >> 33) % (uint64_t)have] * N + j];
          if (dv < bd) { bd = dv; bj = j; }
        }
        if (bj >= 0) J[have++] = bj;
      }
      kk = have;
      // state space of the slot counters must fit: drop trailing nodes otherwise
      int S_ = 1, fit = 0;
      for (int q = 0; q < kk; ++q) { const int rd = s_slots[J[q]] + 1; if ((int64_t)S_ * rd > smax) break; S_ *= rd; ++fit; }
      kk = fit;
      if (lane == 0) {
        int mu_ = 1;
        for (int q = 0; q < kk; ++q) { Jn[q] = J[q]; rad[q] = s_slots[J[q]] + 1; mul[q] = mu_; mu_ *= rad[q]; }
        misc[32] = mu_;                                    // number of states
      }
      for (int q = 0; q < kk; ++q) { if (J[q] < 64) jm0 |= 1ull << J[q]; else jm1 |= 1ull << (J[q] - 64); }
    }
    __syncwarp();
    if (kk == 0) continue;
    const int nTk = 1 << kk;
    for (int T = lane; T < nTk; T += 32) { int s = 0; for (int q = 0; q < kk; ++q) if ((T >> q) & 1) s += mul[q]; incT[T] = s; }

    // ---- priced routing of the current placement + dual ascent on the node prices ---------------------------
    double g = 0.0; bool overloaded = false;
    for (int pass = 0; pass < 4; ++pass) {
      for (int j = lane; j < N; j += 32) loadfx[j] = 0ull;
      __syncwarp();
      double gp = 0.0;
      for (int fi = lane; fi < fn; fi += 32) {
        const int f = fi / N, i = fi - f * N;
        const double* di = s_d + i * N;
        const double* rf = s_r + f * N;
        const uint8_t* pl = podlist + f * N;
        const int np = npods[f];
        double best = kLnsBig, bout = kLnsBig; int bj = 0;
        for (int q = 0; q < np; ++q) {
          const int j = pl[q];
          const double v = di[j] + lam[j] * rf[j];
          if (v < best) { best = v; bj = j; }
          const bool inJ = j < 64 ? (jm0 >> j) & 1 : (jm1 >> (j - 64)) & 1;
          if (!inJ && v < bout) bout = v;
        }
        asg[fi] = (uint8_t)bj; base[fi] = bout;
        const double wv = s_w[fi];
        if (wv > 0.0) {
          gp += wv * best;
          const double ld = wv * rf[bj];
          if (ld > 0.0) atomicAdd(&loadfx[bj], (unsigned long long)(fmin(ld, 1e9) * kFxScale + 0.5));
        }
      }
      g = warp_sum(gp);
      __syncwarp();
      // nodes to (re)price: overloaded at the current prices, or carrying a price
      bool changed = false; overloaded = false;
      for (int jb = 0; jb < N; jb += 32) {
        const int j = jb + lane;
        bool need = false, over = false;
        if (j < N) {
          const double ld = (double)loadfx[j] * (1.0 / kFxScale);
          over = ld > s_K[j] + 1e-7;
          need = over || lam[j] > 0.0;
        }
        unsigned todo = __ballot_sync(0xffffffffu, need);
        overloaded = overloaded || __any_sync(0xffffffffu, over);
        if (pass == 3) continue;                           // last pass only evaluates
        while (todo) {
          const int jj = jb + (__ffs(todo) - 1); todo &= todo - 1;
          // thresholds of the sources that can use node jj:  th = (priced alternative - d[i,jj]) / r[f,jj]
          int nf = 0;
          for (int f = 0; f < F; ++f) if (c[f * N + jj]) { if (lane == 0) fl[nf] = f; ++nf; }
          __syncwarp();
          const int M = nf * N;
          double tot = 0.0;
          for (int t = lane; t < M; t += 32) {
            const int f = fl[t / N], i = t % N;
            const double rfj = s_r[f * N + jj], av = s_w[f * N + i] * rfj;
            float tv = 0.0f;
            if (av > 0.0) {
              const double* di = s_d + i * N; const double* rf = s_r + f * N; const uint8_t* pl = podlist + f * N;
              double alt = kLnsBig;
              for (int q = 0; q < npods[f]; ++q) { const int j2 = pl[q]; if (j2 == jj) continue; const double v = di[j2] + lam[j2] * rf[j2]; if (v < alt) alt = v; }
              const double tq = (alt - di[jj]) / rfj;
              if (tq > 0.0) { tv = (float)fmin(tq, 1e30); tot += av; }
            }
            th[t] = tv;                                     // 0: never uses jj
          }
          tot = warp_sum(tot);
          __syncwarp();
          double nl = 0.0;
          if (tot > s_K[jj] + 1e-9) {
            // flows leave in ascending threshold order until the rest fits; the price is the last threshold
            float last = 0.0f;
            for (int it = 0; it < 4 * N * 8 && tot > s_K[jj] + 1e-9; ++it) {
              unsigned mn = 0xffffffffu;
              for (int t = lane; t < M; t += 32) { const float tv = th[t]; if (tv > last) mn = min(mn, __float_as_uint(tv)); }
              mn = __reduce_min_sync(0xffffffffu, mn);
              if (mn == 0xffffffffu) break;                 // nothing left to move: stays overloaded
              last = __uint_as_float(mn);
              double rem = 0.0;
              for (int t = lane; t < M; t += 32) if (th[t] == last) rem += s_w[fl[t / N] * N + t % N] * s_r[fl[t / N] * N + jj];
              tot -= warp_sum(rem);
            }
            nl = (double)last * (1.0 + 2e-6) + 1e-12;
          }
          if (fabs(nl - lam[jj]) > 1e-10 * (1.0 + nl)) changed = true;
          __syncwarp();
          if (lane == 0) lam[jj] = nl;
          __syncwarp();
        }
      }
      if (!changed) break;
    }
    {
      double lk = 0.0;
      for (int j = lane; j < N; j += 32) lk += lam[j] * s_K[j];
      g -= warp_sum(lk);
      if (u != 0.0) {
        int act = 0;
        for (int j = lane; j < N; j += 32) { int any = 0; for (int f = 0; f < F; ++f) any |= c[f * N + j]; act += any; }
        act = __reduce_add_sync(0xffffffffu, act);
        g += u * (double)act;
      }
    }
    // ---- record ------------------------------------------------------------------------------------------
    if (!overloaded && g < bestg - 1e-9 * (1.0 + fabs(g))) {
      bestg = g; best_round = round;
      for (int q = lane; q < fn; q += 32) outc[q] = c[q];
    }
    if (round == a.rounds) break;

    // ---- cost of every subset T of the k nodes, per function -------------------------------------------------
    const double temp = round < anneal_rounds ? noise0 * (1.0 - (double)round / (double)anneal_rounds) : 0.0;
    const int groups = 32 / nTk, T = lane & (nTk - 1), ig = lane / nTk;
    const uint64_t rround = lns_next(rs);
    for (int f = 0; f < F; ++f) {
      double pj[kLnsMaxK];
#pragma unroll
      for (int q = 0; q < kLnsMaxK; ++q) pj[q] = q < kk ? lam[Jn[q]] * s_r[f * N + Jn[q]] : 0.0;
      double acc = 0.0; int outside = 0;
      for (int q = 0; q < npods[f]; ++q) { const int j = podlist[f * N + q]; outside |= !(j < 64 ? (jm0 >> j) & 1 : (jm1 >> (j - 64)) & 1); }
      for (int i = ig; i < N; i += groups) {
        const double wv = s_w[f * N + i];
        if (wv == 0.0) continue;
        double val = base[f * N + i];
#pragma unroll
        for (int q = 0; q < kLnsMaxK; ++q) if (q < kk && ((T >> q) & 1)) val = fmin(val, s_d[i * N + Jn[q]] + pj[q]);
        acc += wv * val;
      }
      for (int o = nTk; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane < nTk) {
        double v = acc;
        if (temp > 0.0) v += temp * lns_u01(lns_mix(rround ^ ((uint64_t)(f * 64 + T) << 8))) * (double)__popc(T);
        if (T == 0 && !outside) v = INFINITY;               // the function would lose its last pod
        costT[f * nT + T] = v;
      }
    }
    __syncwarp();

    // ---- DP over functions, state = slot counters of the k nodes -----------------------------------------------
    const int S_ = misc[32];
    for (int s = lane; s < S_; s += 32) cur[s] = s == 0 ? 0.0 : INFINITY;
    __syncwarp();
    double* pc = cur; double* pn = nxt;
    for (int f = 0; f < F; ++f) {
      for (int s2 = lane; s2 < S_; s2 += 32) {
        int supp = 0;
        for (int q = 0; q < kk; ++q) if ((s2 / mul[q]) % rad[q] > 0) supp |= 1 << q;
        double best = INFINITY; int bt = 0;
        int Ts = supp;
        while (true) {
          const double v = pc[s2 - incT[Ts]] + costT[f * nT + Ts];
          if (v < best || (v == best && Ts < bt)) { best = v; bt = Ts; }
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
    for (int s = lane; s < S_; s += 32) {
      double v = pc[s];
      if (u != 0.0) { int used = 0; for (int q = 0; q < kk; ++q) used += (s / mul[q]) % rad[q] > 0; v += u * (double)used; }
      if (v < bv) { bv = v; bs = s; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double v2 = __shfl_xor_sync(0xffffffffu, bv, o); const int s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      if (v2 < bv || (v2 == bv && s2 < bs)) { bv = v2; bs = s2; }
    }
    if (bv < INFINITY && lane == 0) {
      int s = bs;
      for (int f = F - 1; f >= 0; --f) {
        const int Tc = choice[f * smax + s];
        for (int q = 0; q < kk; ++q) c[f * N + Jn[q]] = (Tc >> q) & 1;
        s -= incT[Tc];
      }
    }
    __syncwarp();
    rebuild_pods();
  }
  if (lane == 0) {
    a.out_g[(int64_t)b * a.chains + chain] = bestg < INFINITY ? a_d * bestg : INFINITY;
    a.out_round[(int64_t)b * a.chains + chain] = best_round;
  }
  if (best_round < 0) for (int q = lane; q < fn; q += 32) outc[q] = c[q];
}

}  // namespace neptune

using namespace neptune;

extern "C" int neptune_lns_search(int B, int N, int F, int kind, double alpha, int chains, int rounds, int k,
                                  double noise_coef, uint64_t rng_seed, const double* d, const double* w,
                                  const double* r, const double* m, const double* Mj, const double* Kj,
                                  const double* maxd, const double* guide, const double* lam0, int S,
                                  const uint8_t* seeds, uint8_t* out_c, double* out_g, int32_t* out_round,
                                  void* stream) {
  if (B <= 0 || N <= 0 || F <= 0 || chains <= 0 || rounds < 0 || k < 1 || k > kLnsMaxK || !d || !w || !r || !m || !Mj ||
      !Kj || !out_c || !out_g || !out_round)
    return NEPTUNE_E_ARG;
  if (kind != NEPTUNE_KIND_MIN_DELAY && kind != NEPTUNE_KIND_MIN_DELAY_UTIL) return NEPTUNE_E_ARG;
  if (seeds && S <= 0) return NEPTUNE_E_ARG;
  if (N > 128 || (int64_t)F * N > 4096) return NEPTUNE_E_SIZE;       // the chain state lives in shared memory
  LnsArgs a{};
  a.B = B; a.N = N; a.F = F; a.kind = kind; a.chains = chains; a.rounds = rounds; a.k = k; a.S = seeds ? S : 0;
  a.alpha = alpha; a.noise_coef = noise_coef; a.rng = rng_seed ? rng_seed : 0x1234567ull;
  a.d = d; a.w = w; a.r = r; a.m = m; a.Mj = Mj; a.Kj = Kj; a.maxd = maxd; a.guide = guide; a.lam0 = lam0; a.seeds = seeds;
  a.out_c = out_c; a.out_g = out_g; a.out_round = out_round;
  // states of the slot counters: (slots+1)^k for the common 3-slot nodes, capped by shared memory
  int smax = 1; for (int q = 0; q < k; ++q) smax *= 4;
  const size_t blk = (lns_block_shared(N, F) + 15) & ~(size_t)15;
  int wpb = 8;
  size_t per = lns_warp_shared(N, F, k, smax);
  while (wpb > 1 && blk + wpb * per > 100 * 1024) wpb >>= 1;          // aim at two blocks per SM
  if (blk + wpb * per > 200 * 1024) return NEPTUNE_E_SIZE;
  if (wpb > chains) { wpb = 1; while (wpb * 2 <= chains) wpb *= 2; }
  a.smax = smax; a.wpb = wpb;
  const size_t sm = blk + wpb * per;
  NEPTUNE_CUDA_OK(cudaFuncSetAttribute(k_lns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  { k_lns<<<dim3((chains + wpb - 1) / wpb, B), wpb * 32, sm, (cudaStream_t)stream>>>(a); NEPTUNE_COUNT(1); }
  NEPTUNE_LAUNCH_OK();
  return 0;
}
