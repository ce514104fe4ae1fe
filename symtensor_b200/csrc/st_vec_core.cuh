// Core of the tail-table vector contraction (see st_vec.cu): strategy record and the per-warp walk.
// Host+device so that the index arithmetic can be exercised on the CPU by tests/emu (test-only harness that
// is NOT part of libsymtensor_b200.so).
#pragma once

#include "st_common.cuh"

namespace st {

// Per-class strategy of the tail-table kernel (device copy lives beside the plan).
struct TailStrategy {
  int32_t tau;   // tail length, 1 <= tau <= g_t
  int32_t hn;    // g_t - tau: head values inside the last run
  int32_t gt;    // length of the last run
  int32_t nE;    // values in earlier runs
  int32_t Rt;    // dim - nE: values available to the last run
  int32_t mu;    // multiplicity of the last run
  int32_t direct;  // ring kernel, tau == 2: hybrid pair walk -- rows k < k0 of the pair table are walked column-wise with
                   // weights xr[k] * xr[l] computed on the fly, rows k >= k0 come from a suffix table in shared memory
  int32_t k0;      // direct: first row held by the suffix table (Rt: no table at all)
  int64_t tbl_n; // C(Rt, tau)
  int64_t seg;   // C(Rt, g_t): components per fixed assignment of the earlier runs
};

// ------------------------------------------------------------------------------------------------------
// tail-table kernel
// ------------------------------------------------------------------------------------------------------
// A class is a sequence of SEGMENTS (one per assignment E of the earlier runs, `seg` components each); inside
// a segment the last run is an increasing g_t-combination u of the Rt relabelled values not in E; its first
// hn values are the "head", the last tau the "tail".  For a fixed head the tail components are one contiguous
// BLOCK of memory whose weights are a contiguous slice of the table
//     T[q] = prod_{u in q-th tau-combination of range(Rt)} xr[u],      xr[u] = x[actual(u)]^mu .
// Two table ownerships per class (host cost model):
//   shared  (tau >= 2)  T and xr live once per CTA and are rebuilt (with __syncthreads) when E changes;
//   private (tau == 1)  T == xr, one copy per warp, rebuilt by the warp itself when it enters a new segment --
//                       used when segments are too short to amortise a CTA-wide rebuild.
#define ST_MAX_RED 26  // 1 + the classes a RingSched can hold (more classes: several passes)
struct TailCtrl {
  double red[32];
  double wE;               // gamma * prod over earlier runs x[v]^m
  int32_t E[ST_MAX_RANK];  // values of the earlier runs, ascending
  int32_t cur_cls;
  int64_t cur_seg;
  int32_t last;            // this CTA is the last one to finish (fused finalize)
  // final reduction (ring kernel): the slot ranges to add, as one concatenated index space
  int32_t red_n, item_next;
  int64_t red_start[ST_MAX_RED + 1];
  const double* red_ptr[ST_MAX_RED];
};

// Warp-uniform small arrays of the walk.  On the device one copy per warp lives in SHARED memory: as thread-local
// arrays they would be indexed dynamically, i.e. sit in local memory, 32 identical copies per warp, and with
// most of the L1 carved out for the tail table every access would go to L2 (measured: a single unrank took
// tens of microseconds).  All lanes of a converged warp write the same values; on the host (tests/emu) it is a
// plain local.
struct alignas(16) DirEntry {  // tile directory entry, see dir_decode
  uint16_t v[ST_MAX_RANK];
};

struct WarpScratch {
  DirEntry de;
  double pref[ST_MAX_RANK + 1];
  int64_t dig[ST_MAX_RANK];
  int32_t u[ST_MAX_RANK];
  int32_t E[ST_MAX_RANK];
  int32_t u0[ST_MAX_RANK];
  int32_t vals[ST_MAX_RANK];
};

// earlier runs of segment `sidx`: values (ascending, in E) and the weight gamma * prod x[v]^m
template <typename T>
ST_HD double unrank_earlier(const PlanView& P, const ClassDesc& C, int64_t sidx, const T* __restrict__ xs,
                                                 int32_t* E, WarpScratch& ws) {
  double w = (double)C.gamma;
  if (C.nruns <= 1) return w;
  int32_t* vals = ws.vals;
  int64_t* dig = ws.dig;
  int nused = 0;
  for (int j = C.nruns - 2; j >= 0; --j) { const int64_t q = sidx / C.radix[j]; dig[j] = sidx - q * C.radix[j]; sidx = q; }
  for (int j = 0; j < C.nruns - 1; ++j) {
    const int g = C.run_len[j], s = C.run_start[j];
    comb_unrank(P.binom, P.rank, dig[j], P.dim - nused, g, vals + s);
    for (int i = 0; i < g; ++i) {
      int32_t v = vals[s + i];
      for (int e = 0; e < nused; ++e) v += (v >= E[e]);
      vals[s + i] = v;
      const double xv = (double)xs[v];
      for (int m = 0; m < C.run_mult[j]; ++m) w *= xv;
    }
    for (int i = 0; i < g; ++i) {
      const int32_t v = vals[s + i];
      int e = nused++;
      while (e > 0 && E[e - 1] > v) { E[e] = E[e - 1]; --e; }
      E[e] = v;
    }
  }
  return w;
}

template <typename T>
ST_HD T xrel_pow(const T* __restrict__ xs, const int32_t* E, int nE, int mu, int32_t u) {
  int32_t v = u;
  for (int e = 0; e < nE; ++e) v += (v >= E[e]);
  const T xv = xs[v];
  T p = xv;
  for (int m = 1; m < mu; ++m) p *= xv;
  return p;
}

#ifdef __CUDA_ARCH__
#define ST_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define ST_ASSUME_SHARED(p)
#endif

#ifdef __CUDACC__
// Warp-cooperative inverse of comb_rank: every lane tests one candidate value per step (ballot), so a
// combination is unranked in a few steps instead of g binary searches of ~40 64-bit instructions per probe.
// All lanes return the same combination.  Must be called by a full, converged warp.
__device__ __forceinline__ void comb_unrank_warp(const int64_t* __restrict__ tbl, int rank, int64_t r, int n, int g, int32_t* c, int lane) {
  int prev = -1;
  for (int i = 0; i < g; ++i) {
    const int k = g - i;
    const int64_t all = binom_at(tbl, rank, n - 1 - prev, k);
    // largest v in [prev + 1, n - k] with  all - C(n - v, k) <= r ; the predicate is monotone (true ... true false ...)
    int v = prev + 1;
    while (true) {
      const int cand = v + lane;
      const bool ok = cand <= n - k && all - binom_at(tbl, rank, n - cand, k) <= r;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      const int cnt = __popc(m);
      v += cnt;
      if (cnt < 32) break;
    }
    v -= 1;  // last candidate that satisfied the predicate (v = prev + 1 always does)
    r -= all - binom_at(tbl, rank, n - v, k);
    c[i] = v;
    prev = v;
  }
}
#endif

// ------------------------------------------------------------------------------------------------------
// walk_tile: one warp's walk over positions [q0, q0 + n) of one segment (vec_ring_kernel, st_vec.cu).  The odometer
// hands out "pieces" (block ∩ range, positions relative to q0) with  weight(idx) = hw * tbl[toff + idx]; its common
// step -- increment the last head value -- costs a handful of warp-uniform instructions, carries take the general
// path over the arrays u[] / pref[] of the warp's scratch.  `u_init` is the last run's combination at q0 when the
// caller knows it (tile directory, start of a segment); otherwise it is unranked here.  The
// components are NOT fetched here: they arrive in shared memory through the warp's ring of cp.async.bulk
// copies (the memory system runs ahead on its own), and this function only multiplies what is already
// on chip.  `src.chunk(k)` hands out a pointer to the elements [k * Bel, (k + 1) * Bel) of the warp's
// current tile (it waits for that copy, and releases -- i.e. refills -- the slots of earlier sub-chunks);
// the range walked is [q0, q0 + n) of the segment, whose first element sits at position `t0` of the tile.
// Per element: one LDS of the component, one LDS of the table, one FMA; a piece costs one odometer step.
// On the host (tests/emu) `src` is a plain array view and the function is replayed lane by lane.
// ------------------------------------------------------------------------------------------------------
// DIRECT (strategy `direct`, tau == 2) is the hybrid pair walk: the pair table's rows (fixed k, l = k+1 .. Rt-1,
// row k starts at table index rs(k) = k Rt - k (k + 1) / 2) with k >= k0 come from a SUFFIX TABLE in shared
// memory, the rows below k0 -- the long ones -- are walked row by row with the weights xr[k] * xr[l] formed on
// the fly:  racc += xr[k] * sum_l A[(k, l)] xr[l]  (per component one LDS of the component, one of xr, one FMA;
// per row a handful of warp-uniform instructions).  A 96 KB suffix table covers 97 % of the components of the
// rank-4 dim-200 class (1,1,1,1) (the full table would be 159 KB and leave no room for the rings), and the
// dimension is no longer capped by the table.
// GAP (direct classes with earlier runs, last-run multiplicity 1): `xr` is x itself, indexed by ACTUAL values, and the
// relabelling u -> u + #{e : u + e >= E[e]} (skip the values E of the earlier runs, kept ascending in ws.E) is applied
// on the fly -- a row's columns split into at most nE + 1 contiguous pieces of x -- so that these classes need no
// per-warp relabelled copy of x in shared memory.
template <typename T, bool DIRECT, bool GAP, typename Src>
ST_HD double walk_tile(const PlanView& P, const TailStrategy& S, const T* __restrict__ tbl, const T* __restrict__ xr,
                       const int32_t* __restrict__ blen, double wE, int64_t q0, int n, int t0, int lane, const int32_t* u_init, Src& src,
                       WarpScratch& ws) {
  ST_ASSUME_SHARED(tbl);
  ST_ASSUME_SHARED(xr);
  const int64_t* __restrict__ bt = P.binom;
  const int rk = P.rank;
  const int gt = S.gt, hn = S.hn, tau = S.tau, Rt = S.Rt;
  const int tbl_n = (int)S.tbl_n;
  const int nE = GAP ? S.nE : 0;
  const int32_t* __restrict__ Ev = ws.E;
  const int e0 = GAP ? Ev[0] : 0;
  // relabelled value -> table entry (GAP: skip the earlier runs' values)
  auto xat = [&](int uu) -> T {
    if (GAP) {
      if (nE == 1) uu += (uu >= e0);
      else for (int e = 0; e < nE; ++e) uu += (uu >= Ev[e]);
    }
    return xr[uu];
  };
  // ---- odometer state at q0
  int32_t* u = ws.u;
  double* pref = ws.pref;
  if (u_init) {
    if (u_init != u) for (int i = 0; i < gt; ++i) u[i] = u_init[i];
  } else {
#ifdef __CUDA_ARCH__
    comb_unrank_warp(bt, rk, q0, Rt, gt, u, lane);
#else
    comb_unrank(bt, rk, q0, Rt, gt, u);
#endif
  }
  int row_k = DIRECT ? u[hn] : 0;                                 // row walk: row of the current position ...
  int row_s = DIRECT ? row_k * Rt - row_k * (row_k + 1) / 2 : 0;  // ... and its first table index rs(row_k)
  const int tq = DIRECT   ? row_s + (u[hn + 1] - row_k - 1)
                 : tau == 2 ? u[hn] * Rt - u[hn] * (u[hn] + 1) / 2 + (u[hn + 1] - u[hn] - 1)
                 : tau == 1 ? u[hn]
                            : (int)comb_rank(bt, rk, u + hn, Rt, tau);
  pref[0] = wE;
  for (int i = 0; i < hn; ++i) pref[i + 1] = pref[i] * (double)xat(u[i]);
  int u_last = hn ? u[hn - 1] : -1;
  double pref_prev = hn ? pref[hn - 1] : wE;
  double hw = pref[hn];
  int pb = 0;
  int pe = (tbl_n - tq < n) ? tbl_n - tq : n;
  int toff = tq;
  T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
  double total = 0.0;
  auto advance = [&]() {
    pb = pe;
    if (hn == 0) { pe = n; return; }
    if (u_last + 1 <= Rt - tau - 1) {
      ++u_last;
    } else {
      u[hn - 1] = u_last;
      int j = hn - 1;
      while (j >= 0 && u[j] + 1 > Rt - (gt - j)) --j;
      if (j < 0) { pe = n; return; }
      ++u[j];
      for (int k = j + 1; k < hn; ++k) u[k] = u[k - 1] + 1;
      for (int k = j; k < hn - 1; ++k) pref[k + 1] = pref[k] * (double)xat(u[k]);
      u_last = u[hn - 1];
      pref_prev = pref[hn - 1];
    }
    hw = pref_prev * (double)xat(u_last);
    const int m = Rt - 1 - u_last;
    const int bl = DIRECT ? m * (m - 1) / 2 : tau == 1 ? m : (blen ? blen[u_last] : (int)binom_at(bt, rk, m, tau));
    toff = tbl_n - bl - pb;
    pe = (bl < n - pb) ? pb + bl : n;
    if (DIRECT) { row_k = u_last + 1; row_s = tbl_n - bl; }  // the block starts with row u_last + 1
  };
  const int qs = DIRECT ? (S.k0 >= Rt - 1 ? tbl_n : S.k0 * Rt - S.k0 * (S.k0 + 1) / 2) : 0;  // first table index held by the suffix table
  // row walk over the range-relative positions [a, b) of the current piece (all of them in rows below k0); dp[e] = component e.
  // Two forms.  Long rows (classes without earlier runs: the rows below k0 are the long ones): row by row, the warp
  // along the row.  GAP classes (every row of the pair table, average length Rt / 3, each cut in two by the
  // relabelling): 32 consecutive positions at a time, one per lane; the row of the group's first position is
  // tracked warp-uniformly and a lane steps at most one row further (more only in the last 32 rows).
  auto row_range = [&](const T* __restrict__ dp, int a, int b) {
    if (GAP) {
      for (int g0 = a; g0 < b; g0 += 128) {  // four groups of 32 positions per step (independent chains)
        T w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w[j] = T(0);
          if (g0 + 32 * j < b) {  // warp-uniform
            const int qj = toff + g0 + 32 * j;
            int len = Rt - 1 - row_k;
            while (qj >= row_s + len) { row_s += len; ++row_k; --len; }  // row of the group's first position
            const int q = qj + lane;
            if (g0 + 32 * j + lane < b) {  // this lane's position exists
              int k = row_k, rs = row_s;
              if (len >= 32) {  // warp-uniform: at most one row boundary inside the group
                const int ns = row_s + len;
                if (q >= ns) { k = row_k + 1; rs = ns; }
              } else {          // the last 32 rows of the pair table: several short rows per group
                int ln = len;
                while (q >= rs + ln) { rs += ln; ++k; --ln; }
              }
              int l = q - rs + k + 1;
              if (nE == 1) {  // the common case: one earlier value, kept in a register
                k += (k >= e0);
                l += (l >= e0);
              } else {
                for (int e = 0; e < nE; ++e) { k += (k >= Ev[e]); l += (l >= Ev[e]); }
              }
              w[j] = xr[k] * xr[l];
            }
          }
        }
        const T* __restrict__ d0 = dp + (g0 + lane);
        const int left = b - g0 - lane;  // this lane's positions g0 + lane + 32 j exist for 32 j < left
        if (left > 0) s0 += d0[0] * w[0];
        if (left > 32) s1 += d0[32] * w[1];
        if (left > 64) s2 += d0[64] * w[2];
        if (left > 96) s3 += d0[96] * w[3];
      }
      return;
    }
    int qa = toff + a;
    const int qb = toff + b;
    while (qa < qb) {
      const int rend = row_s + (Rt - 1 - row_k);
      const int se = qb < rend ? qb : rend;
      const T xv = xr[row_k];
      const T* __restrict__ xp = xr + (row_k + 1 - row_s + toff);  // xp[e] = xr[l] for the component (row_k, l) at position e
      T r0 = T(0), r1 = T(0);
      const int e = qa - toff + lane;
      const T* __restrict__ d0 = dp + e;
      const T* __restrict__ x0 = xp + e;
      int left = se - toff - e;  // this lane's components of the row: offsets 0, 32, 64, ... below `left`
      for (; left > 32; left -= 64) {
        r0 += d0[0] * x0[0];
        r1 += d0[32] * x0[32];
        d0 += 64;
        x0 += 64;
      }
      if (left > 0) r0 += d0[0] * x0[0];
      s3 += xv * (r0 + r1);
      qa = se;
      if (se == rend) { row_s = rend; ++row_k; }
    }
  };
  const int Bel = src.bel();
  int kc = t0 ? t0 / Bel : 0;  // sub-chunk of the tile that holds the range start
  int cb = kc * Bel - t0;      // range-relative position of its first element (<= 0)
  while (true) {
    const T* __restrict__ dp = src.chunk(kc) - cb;  // dp[e]: component at range-relative position e
    ST_ASSUME_SHARED(dp);
    const int ce = (cb + Bel < n) ? cb + Bel : n;
    bool done = false;
    while (true) {
      const int a = pb > cb ? pb : cb;
      const int b = pe < ce ? pe : ce;
      int as = a;
      if (DIRECT) {
        // rows below k0 row by row, the rest against the suffix table
        as = qs - toff;  // first position of the piece held by the table
        as = as < a ? a : (as > b ? b : as);
        if (a < as) row_range(dp, a, as);
      }
      {
        const T* __restrict__ tp = tbl + (toff - qs);  // tp[e]: weight of range-relative position e
        for (int f = as + lane; f < b; f += 128) {  // four independent chains; the ragged end is predicated, not looped
          s0 += dp[f] * tp[f];
          if (f + 32 < b) s1 += dp[f + 32] * tp[f + 32];
          if (f + 64 < b) s2 += dp[f + 64] * tp[f + 64];
          if (f + 96 < b) s3 += dp[f + 96] * tp[f + 96];
        }
      }
      if (pe > ce) break;  // the piece continues in the next sub-chunk
      total += hw * (((double)s0 + (double)s1) + ((double)s2 + (double)s3));
      s0 = T(0);
      s1 = T(0);
      s2 = T(0);
      s3 = T(0);
      if (pe >= n) { done = true; break; }
      advance();
      if (pb >= ce) break;
    }
    if (done) break;
    ++kc;
    cb += Bel;
  }
  return total;
}

template <typename T, typename Src>
ST_HD double walk_tile_any(const PlanView& P, const TailStrategy& S, const T* __restrict__ tbl, const T* __restrict__ xr,
                           const int32_t* __restrict__ blen, double wE, int64_t q0, int n, int t0, int lane, const int32_t* u_init, Src& src,
                           WarpScratch& ws) {
  if (S.direct) {
    if (S.nE != 0 && S.mu == 1) return walk_tile<T, true, true, Src>(P, S, tbl, xr, blen, wE, q0, n, t0, lane, u_init, src, ws);
    return walk_tile<T, true, false, Src>(P, S, tbl, xr, blen, wE, q0, n, t0, lane, u_init, src, ws);
  }
  return walk_tile<T, false, false, Src>(P, S, tbl, xr, blen, wE, q0, n, t0, lane, u_init, src, ws);
}

// host-side source of walk_tile (tests/emu): the tile is a plain array
template <typename T>
struct ArraySrc {
  const T* tile;  // first element of the tile
  int bel_;
  ST_HD int bel() const { return bel_; }
  ST_HD const T* chunk(int k) { return tile + (int64_t)k * bel_; }
};

// ------------------------------------------------------------------------------------------------------
// ring kernel plumbing (vec_ring_kernel, st_vec.cu): per-warp TMA rings
// ------------------------------------------------------------------------------------------------------
// Per-class record kept in shared memory (everything the scheduling loop needs, so that it never waits on a
// global load).
struct ClsInfo {
  int64_t offset, size;
  int64_t tile_base;  // tiles of the classes before this one (whole tensor): index into the tile directory
  int64_t sbase;      // components of the SMALL classes before this one: index into the per-component directory
  TailStrategy S;
};

// Per-class record of one launch (computed on the host and passed in the kernel parameters when the class count
// allows, else by one thread of the CTA): the part of the class inside the launch range, cut into tiles and
// chunks of NW tiles.  Static schedule: chunk c of the class belongs to CTA (c + rot) mod G.
struct ClsRun {
  int64_t lo, hi;    // positions of the class inside the launch range
  int64_t k0, k1;    // tiles k0 .. k1-1 (tile k = positions [k * tile, (k + 1) * tile))
  int64_t ch0, ch1;  // chunks ch0 .. ch1-1 (chunk c = tiles [c * NW, (c + 1) * NW))
  int64_t ntail;     // dynamic deal: the class's last ntail tiles (many tiny blocks: the costly ones) are dealt FIRST, last tile first
  int64_t s0, ns;    // dynamic deal: the class's first ns deal entries go to the warps s0 .. s0 + ns - 1 of the grid without a claim
  int64_t nd;        // deal entries of the class (== k1 - k0 unless tiles are grouped)
  int64_t kf, ke;    // grouped deal: tiles kf .. ke-1 are dealt one by one (last), the tiles on either side in groups of m
  int32_t m;         // group size (1: no grouping)
  int32_t rot;       // chunks of the earlier classes, mod G
  int32_t mode;      // 0: not walked (small class / empty range); 1: tiles per warp (mode A); 2: chunks per CTA (mode B)
  int32_t pad_;
};

// Deal order of a mode-A class (guided, costly pieces first).  A tile's cost is highest at the START of a class (long
// rows below the suffix table: row walk) and at its END (short blocks: many pieces per tile), and lowest and most even in
// between.  Entry n < ntail -> the costly tail tiles, last tile first; then GROUPS of m consecutive tiles over [k0, kf),
// ascending; then groups over [ke, k1 - ntail), last group first; then the single tiles kf .. ke-1 -- the cheap, even
// ones -- which make the launch's tail short.  Returns the first tile of the entry (>= k1: exhausted), its tile count in tn.
ST_HD int64_t deal_to_tile(const ClsRun& r, int64_t n, int32_t& tn) {
  tn = 1;
  if (n < r.ntail) return r.k1 - 1 - n;
  n -= r.ntail;
  const int64_t nl = (r.kf - r.k0 + r.m - 1) / r.m;
  if (n < nl) {
    const int64_t t = r.k0 + n * r.m;
    tn = (int32_t)(r.kf - t < r.m ? r.kf - t : r.m);
    return t;
  }
  n -= nl;
  const int64_t E = r.k1 - r.ntail;
  const int64_t nu = (E - r.ke + r.m - 1) / r.m;
  if (n < nu) {
    const int64_t hi = E - n * r.m;
    const int64_t lo = hi - r.m > r.ke ? hi - r.m : r.ke;
    tn = (int32_t)(hi - lo);
    return lo;
  }
  const int64_t t = r.kf + (n - nu);
  return t < r.ke ? t : r.k1;
}
// inverse: deal entry whose first tile is tk (the slot of the entry's sum), and its tile count
ST_HD int64_t tile_to_deal(const ClsRun& r, int64_t tk, int32_t& tn) {
  tn = 1;
  const int64_t E = r.k1 - r.ntail;
  if (tk >= E) return r.k1 - 1 - tk;
  const int64_t nl = (r.kf - r.k0 + r.m - 1) / r.m;
  if (tk < r.kf) {
    tn = (int32_t)(r.kf - tk < r.m ? r.kf - tk : r.m);
    return r.ntail + (tk - r.k0) / r.m;
  }
  if (tk >= r.ke) {
    const int64_t g = (E - 1 - tk) / r.m;
    tn = (int32_t)(E - g * r.m - tk);
    return r.ntail + nl + g;
  }
  return r.ntail + nl + (E - r.ke + r.m - 1) / r.m + (tk - r.kf);
}

// the launch's schedule: one record per class (serial)
// `ntail_in` (per class, may be null), `group`, `fine`: the dynamic deal's shape -- costly tail tiles, tiles per group,
// and how many single tiles per warp of the grid close the class (see deal_to_tile); only the host passes them.
ST_HD void make_runs(const ClsInfo* cls, int ncls, int64_t begin, int64_t end, int64_t tile, int nwarps, int G, ClsRun* run,
                     const int64_t* ntail_in = nullptr, int group = 1, int64_t fine = 0, int fine_pos = 70) {
  int64_t rot = 0, wused = 0;
  const int64_t W = (int64_t)G * nwarps;
  for (int ci = 0; ci < ncls; ++ci) {
    ClsRun r;
    const int64_t coff = cls[ci].offset, csize = cls[ci].size;
    const TailStrategy& S = cls[ci].S;
    r.lo = (begin > coff ? begin : coff) - coff;
    r.hi = (end < coff + csize ? end : coff + csize) - coff;
    r.mode = (S.tau != 0 && r.lo < r.hi) ? ((S.tau == 1 || S.nE == 0 || S.direct) ? 1 : 2) : 0;
    if (!r.mode) { r.lo = 0; r.hi = 0; }
    r.k0 = r.lo / tile;
    r.k1 = (r.hi + tile - 1) / tile;
    r.ch0 = r.k0 / nwarps;
    r.ch1 = (r.k1 + nwarps - 1) / nwarps;
    r.rot = (int32_t)(rot % G);
    if (r.mode) rot += r.ch1 - r.ch0;
    // the first tiles of the stream are dealt statically, one per warp, so that no warp starts with an atomic
    r.ntail = 0;
    r.m = 1;
    r.kf = r.k0;
    r.ke = r.k0;
    r.nd = r.k1 - r.k0;
    r.pad_ = 0;
    if (r.mode == 1 && ntail_in != nullptr && r.hi == csize) r.ntail = ntail_in[ci] < r.k1 - r.k0 ? ntail_in[ci] : r.k1 - r.k0;
    if (r.mode == 1 && S.nE == 0 && group > 1) {  // (classes with earlier runs: uneven tiles, never grouped)
      // Groups cut the number of tile starts (each costs ~2 us of warp time), but a group must stay below a warp's fair
      // share of the class, or the deal cannot even out: m = a warp's share of what is left of the class after `fine`
      // single tiles per warp of the grid, capped at `group` (measured: rank 6 dim 64 wants 3, rank 8 dim 40 none).
      const int64_t body = r.k1 - r.ntail - r.k0;
      const int64_t f = fine * W < body ? fine * W : body;
      int64_t m = ((body - f) * 4 / W + 3) / 4;  // floor(share + 0.75)
      if (m > group) m = group;
      if (m > 1) {
        r.m = (int32_t)m;
        r.kf = r.k0 + (body - f) * fine_pos / 100 / m * m;  // whole groups below the single tiles
        r.ke = r.kf + f;
        r.nd = r.ntail + (r.kf - r.k0) / m + (r.k1 - r.ntail - r.ke + m - 1) / m + f;
      }
    }
    r.s0 = wused;
    r.ns = 0;
    if (r.mode == 1) { r.ns = r.nd < W - wused ? r.nd : W - wused; wused += r.ns; }
    run[ci] = r;
  }
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// lane 0 only, without a branch: order the generic-proxy reads of the slot before the copy, arm the barrier, copy
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void ring_issue(int lane, uint32_t bar, uint32_t dst, const void* src, uint32_t bytes, int fence, uint64_t pol) {
  asm volatile(
      "{\n"
      " .reg .pred p, q, f;\n"
      " setp.eq.s32 p, %0, 0;\n"
      " setp.ne.and.u32 q, %4, 0, p;\n"
      " setp.ne.and.s32 f, %5, 0, p;\n"
      " @f fence.proxy.async.shared::cta;\n"
      " @p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %4;\n"
      " @q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%2], [%3], %4, [%1], %6;\n"
      "}\n" ::"r"(lane),
      "r"(bar), "r"(dst), "l"(src), "r"(bytes), "r"(fence), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

// One tile of a warp's stream, queued by the producer cursor for the consumer cursor (shared memory, per warp)
struct alignas(16) TileQ {
  DirEntry de;   // directory entry of the tile (copied asynchronously when the tile is entered)
  int64_t tk;    // tile index inside the class
  int32_t ci;    // class
  int32_t pad_;
};

// A warp's stream: producer cursor (what to copy next) and consumer cursor (what the walk reads next).  Every
// warp owns R ring slots of Bel components and one mbarrier per slot; the stream is the concatenation of the
// warp's tiles, cut into sub-chunks of Bel components.  The warp itself refills a slot as soon as the
// sub-chunk in it has been consumed (release -> issue), so the producer cursor always runs R sub-chunks ahead
// of the consumer cursor -- across tiles and classes.
// Which tiles a warp gets: classes are visited in order; inside a mode-A class either the static round robin
// (tile tk, tk + G * NW, ...) or -- `ctr` set -- DYNAMICALLY, one atomicAdd on the class's counter per tile:
// tile costs differ by up to 5x (column walk vs table), and a static deal leaves the slowest warp 50 % behind the
// median.  The claim for the tile after next is issued one tile early, so its latency is never waited for.
// Mode-B classes are always static (their chunks are per CTA).  The producer queues every tile it enters
// (class, tile index, directory entry) and the kernel's loops pop them in the same order.
// Host build (tests/emu): a copy is a memcpy at issue time, waits are no-ops, claims are the static deal; the
// cursor logic is the same code.
template <typename T>
struct RingSrc {
  // ring
  T* ring;         // this warp's R slots of Bel elements
  uint32_t ring_s, slot_bytes;  // its shared-space address, bytes per slot
  uint64_t l2pol;  // L2 policy of the stream: evict first (every component is read once; the directory, the tile sums and
                   // the tables' global sources are what should stay in L2)
  uint32_t bar0;   // shared-space address of this warp's R mbarriers
  int R, Bel;
  int lane;
  // stream description
  const ClsRun* run;
  const ClsInfo* cls;
  const T* A;      // points at packed coordinate `begin`
  const DirEntry* dir;      // tile directory (nullptr: none)
  unsigned long long* ctr;  // per-class tile counters (nullptr: static deal)
  TileQ* queue;    // this warp's QD queue entries
  int QD;
  int64_t begin, tile;
  int ncls, NW, G, warp, cta;
  int64_t step;    // G * NW: distance between two tiles of this warp inside a class (static deal)
  int ondemand;    // dynamic deal, ungrouped classes: the last `ondemand` entries per warp of the grid are claimed on demand, not one ahead
  // producer cursor
  int pci;
  int64_t ptk, pk0, pk1, plo, phi;  // tile being copied, the class's tiles, class range
  int64_t pdbase, pn;               // dynamic deal: deal index of claim 0, deal index of the current tile
  int32_t ptn;                      // tiles in the current deal entry (grouped deal)
  unsigned long long pend;          // dynamic deal: the next claim of this class (valid in lane 0)
  bool pdyn, pahead, pstatic_only;  // the current class is dealt dynamically / with claims one tile ahead / entirely by the static first deal
  const T* pcbase;                  // first component of the class
  const T* pbase;                   // first component of the tile being copied
  int plen, poff;                   // its length, offset of the next sub-chunk
  int pslot;
  bool pvalid;
  int qtail, qtail_i;               // tiles queued so far, and its queue slot
  // consumer cursor
  const T* cbase;
  int clen, cur;  // current tile: first component, length, sub-chunk held (or to be acquired)
  int cslot;
  uint32_t cpar;  // parity to wait for on slot cslot
  bool held;
  int qhead, qhead_i;
  int64_t n_issued, n_waited;  // sub-chunk counters (consistency checks of the emulation)

  ST_HD int bel() const { return Bel; }
  ST_HD int64_t jc0(int ci) const {
    int j = cta - run[ci].rot;
    return j < 0 ? j + G : j;
  }

  ST_HD unsigned long long claim(int ci) {
#ifdef __CUDA_ARCH__
    unsigned long long r = 0;
    if (lane == 0) r = atomicAdd(ctr + ci, 1ULL);
    return r;
#else
    return ctr[ci]++;  // emulation: the warps run one after the other, every claim takes the next entry
#endif
  }
  // n-th tile of the class in deal order: the costly tail first (from the end), then the rest in address order;
  // >= pk1 when the class is exhausted
  ST_HD int64_t deal_tile(int64_t n) {
    pn = n;
    return deal_to_tile(run[pci], n, ptn);
  }
  ST_HD int64_t claimed(unsigned long long raw) {
#ifdef __CUDA_ARCH__
    return deal_tile(pdbase + (int64_t)__shfl_sync(0xffffffffu, raw, 0));
#else
    return deal_tile(pdbase + (int64_t)raw);
#endif
  }
  // enter the tile ptk of class pci: source range, queue entry, directory entry
  ST_HD void set_tile() {
    int64_t w0 = ptk * tile, w1 = w0 + ptn * tile;
    if (w0 < plo) w0 = plo;
    if (w1 > phi) w1 = phi;
    pbase = pcbase + w0;
    plen = (int)(w1 - w0);
    poff = 0;
    TileQ& q = queue[qtail_i];
    q.tk = ptk;
    q.ci = pci;
    if (dir != nullptr) {
      const DirEntry* src = dir + (cls[pci].tile_base + ptk);
#ifdef __CUDA_ARCH__
      if (lane < 2) cp_async16(reinterpret_cast<char*>(&q.de) + 16 * lane, reinterpret_cast<const char*>(src) + 16 * lane);
#else
      q.de = *src;
#endif
    }
    ++qtail;
    if (++qtail_i == QD) qtail_i = 0;
  }
  // first tile of this warp in class ci; false: none
  ST_HD bool enter_class(int ci) {
    const ClsRun& r = run[ci];
    if (!r.mode) return false;
    pk0 = r.k0;
    pk1 = r.k1;
    pdyn = ctr != nullptr && r.mode == 1;
    ptn = 1;
    int64_t tk;
    if (pdyn) {
      // classes with many tiles per warp claim one tile ahead (the atomic's latency is never waited for); small
      // classes claim on demand, or the first warps to arrive would take two tiles each and leave none
      pahead = r.nd >= 4 * step;
      pdbase = r.ns;
      pstatic_only = r.ns >= r.nd;
      const int64_t gw = (int64_t)cta * NW + warp;
      if (gw >= r.s0 && gw < r.s0 + r.ns) {
        tk = deal_tile(gw - r.s0);  // this warp's statically dealt tile of the class
        if (pahead) pend = claim(ci);
      } else if (pstatic_only) {
        return false;  // every tile of the class went out with the static first deal: nothing to claim
      } else {
        const unsigned long long first = claim(ci);
        if (pahead) pend = claim(ci);
        tk = claimed(first);
      }
    } else {
      tk = (r.ch0 + jc0(ci)) * NW + warp;
      if (tk < r.k0) tk += step;
    }
    if (tk >= r.k1) return false;
    ptk = tk;
    plo = r.lo;
    phi = r.hi;
    pcbase = A + (cls[ci].offset - begin);
    set_tile();
    return true;
  }
  // position the producer on the next tile of this warp; false: the stream has ended
  ST_HD bool next_tile() {
    if (pdyn) {
      if (pahead) {
        ptk = claimed(pend);
        // the last 2 * step tiles (two per warp of the grid) are claimed on demand: a warp sitting on a claimed but
        // unstarted tile while others have run dry is what the tail of the launch is made of
        if (ptk < pk1) {
          // (a grouped class ends in single tiles -- short pieces -- and keeps claiming ahead to the end)
          if (run[pci].nd - pn > (run[pci].m > 1 ? 0 : (int64_t)ondemand * step)) pend = claim(pci);
          else pahead = false;
        }
      } else if (pstatic_only) {
        ptk = pk1;
      } else {
        ptk = claimed(claim(pci));
      }
    } else {
      ptk += step;
    }
    if (ptk < pk1) { set_tile(); return true; }
    while (++pci < ncls)
      if (enter_class(pci)) return true;
    return false;
  }
  ST_HD void start() {
    step = (int64_t)G * NW;
#ifdef __CUDA_ARCH__
    ring_s = smem_u32(ring);
    l2pol = l2_evict_first_policy();
#else
    l2pol = 0;
#endif
    slot_bytes = (uint32_t)Bel * (uint32_t)sizeof(T);
    pslot = 0;
    pvalid = false;
    qtail = qhead = 0;
    qtail_i = qhead_i = 0;
    pend = 0;
    pdbase = 0;
    ptn = 1;
    pn = 0;
    pdyn = false;
    pahead = false;
    pstatic_only = false;
    cslot = 0;
    cpar = 0;
    held = false;
    cur = 0;
    clen = 0;
    n_issued = 0;
    n_waited = 0;
    for (pci = 0; pci < ncls; ++pci)
      if (enter_class(pci)) { pvalid = true; break; }
    for (int r = 0; r < R; ++r) issue(true);
  }
  ST_HD void wait_slot() {
#ifdef __CUDA_ARCH__
    mbar_wait(bar0 + 8u * cslot, cpar);
#else
    ++n_waited;
#endif
  }
  // copy the next sub-chunk of the stream into slot pslot (which must be free)
  ST_HD void issue(bool fresh = false) {
    (void)fresh;
    if (!pvalid) return;
    int len = plen - poff;
    if (len > Bel) len = Bel;
    const uint32_t bytes = (uint32_t)(len * (int)sizeof(T)) & ~15u;  // whole 16-byte units; the rest is fetched by chunk()
#ifdef __CUDA_ARCH__
    // (the first fill of a slot needs no proxy fence: nothing has read the slot; and a fence there would wait for the
    // prologue's prefetch loads)
    ring_issue(lane, bar0 + 8u * pslot, ring_s + (uint32_t)pslot * slot_bytes, pbase + poff, bytes, fresh ? 0 : 1, l2pol);
#else
    for (uint32_t i = 0; i < bytes / sizeof(T); ++i) ring[(size_t)pslot * Bel + i] = pbase[poff + i];
    ++n_issued;
#endif
    if (++pslot == R) pslot = 0;
    poff += len;
    if (poff >= plen) pvalid = next_tile();
  }
  // consumer: is the next tile of the stream a tile of class ci?
  ST_HD bool head_is(int ci) const { return qhead != qtail && queue[qhead_i].ci == ci; }
  // consumer: take the next tile of the stream (its directory entry has landed when this returns)
  ST_HD const TileQ& pop() {
    const TileQ& q = queue[qhead_i];
    ++qhead;
    if (++qhead_i == QD) qhead_i = 0;
#ifdef __CUDA_ARCH__
    if (dir != nullptr) {
      cp_async_wait_all();
      __syncwarp();
    }
#endif
    return q;
  }
  // consumer: open the tile [base, base + len) -- must be the tile just popped
  ST_HD void open_tile(const T* base, int len) {
    cbase = base;
    clen = len;
    cur = 0;
    held = false;
  }
  ST_HD void release() {
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
    held = false;
    if (++cslot == R) { cslot = 0; cpar ^= 1u; }
    ++cur;
    issue();
  }
  ST_HD const T* chunk(int k) {
    while (cur < k) {
      if (!held) wait_slot();  // not in practice: the walk reads every sub-chunk
      release();
    }
    T* slot = ring + cslot * Bel;
    if (!held) {
      wait_slot();
      held = true;
      const int len = clen - k * Bel;
      if (len <= Bel) {  // last sub-chunk: components behind the last whole 16-byte unit of the tile
        const int nb = (int)(((uint32_t)(len * (int)sizeof(T)) & ~15u) / sizeof(T));
        if (nb < len) {
#ifdef __CUDA_ARCH__
          if (lane < len - nb) slot[nb + lane] = cbase[k * Bel + nb + lane];
          __syncwarp();
#else
          for (int l = 0; l < len - nb; ++l) slot[nb + l] = cbase[k * Bel + nb + l];
#endif
        }
      }
    }
    return slot;
  }
  ST_HD void close_tile() {
    while (cur * Bel < clen) {
      if (!held) wait_slot();
      release();
    }
  }
};

// shared-memory layout of vec_ring_kernel (host and device agree through this one function)
struct RingLayout {
  size_t priv, xr, blen, binom, cls, cdesc, ws, ctl, run, queue, bar, ring, total;
};
ST_HD RingLayout ring_layout(int esize, int64_t dim, int tbl_cap, int priv_cap, int binom_smem, int ncls, int cdesc_smem, int nw, int R, int Bel) {
  RingLayout L;
  size_t off = ((size_t)tbl_cap * esize + 15) / 16 * 16;
  L.priv = off;  // per-warp xr tables of the classes with earlier runs (priv_cap = nw * dim entries, or 0)
  off = (off + (size_t)priv_cap * esize + 15) / 16 * 16;
  L.xr = off;
  off = (off + 2 * (size_t)dim * esize + 15) / 16 * 16;
  L.blen = off;
  off = (off + (size_t)dim * sizeof(int32_t) + 15) / 16 * 16;
  L.binom = off;
  off += (size_t)binom_smem * sizeof(int64_t);
  L.cls = off;
  off += (size_t)ncls * sizeof(ClsInfo);
  L.cdesc = off;
  off += cdesc_smem ? (size_t)ncls * sizeof(ClassDesc) : 0;
  off = (off + 15) / 16 * 16;
  L.ws = off;
  off += (size_t)nw * sizeof(WarpScratch);
  L.ctl = off;
  off = (off + sizeof(TailCtrl) + 15) / 16 * 16;
  L.run = off;
  off = (off + (size_t)ncls * sizeof(ClsRun) + 15) / 16 * 16;
  L.queue = off;
  off += (size_t)nw * (R + 2) * sizeof(TileQ);
  L.bar = off;
  off += (size_t)nw * R * sizeof(uint64_t);
  off = (off + 127) / 128 * 128;
  L.ring = off;
  off += (size_t)nw * R * Bel * esize;
  L.total = off;
  return L;
}

// ------------------------------------------------------------------------------------------------------
// tile directory: the distinct values (class order, absolute) of the component at the start of every tile,
// 16 x uint16 per tile, written once per (rank, dim, tile size) by the GPU index enumerator (st_vec.cu) so that
// starting a tile costs one 32-byte load instead of a combinatorial unrank.
// ------------------------------------------------------------------------------------------------------
// entry -> earlier-run values E (ascending), their weight factor, and the last run's relabelled combination
template <typename T>
ST_HD double dir_decode(const ClassDesc& C, const TailStrategy& S, const DirEntry& d, const T* __restrict__ xs, int32_t* E, int32_t* u) {
  double w = (double)C.gamma;
  int nE = 0;
  for (int j = 0; j < C.nruns - 1; ++j) {
    for (int i = 0; i < C.run_len[j]; ++i) {
      const int32_t v = d.v[C.run_start[j] + i];
      const double xv = (double)xs[v];
      for (int m = 0; m < C.run_mult[j]; ++m) w *= xv;
      int e = nE++;
      while (e > 0 && E[e - 1] > v) { E[e] = E[e - 1]; --e; }
      E[e] = v;
    }
  }
  for (int i = 0; i < S.gt; ++i) {
    const int32_t v = d.v[S.nE + i];
    int32_t below = 0;
    for (int e = 0; e < nE; ++e) below += (E[e] < v);
    u[i] = v - below;
  }
  return w;
}

// T[q] = prod of xr over the q-th tau-combination of range(Rt), for q in [q, qe): contiguous slice of one thread
// (serial reference form; the kernel uses build_table_level below)
template <typename T>
ST_HD void build_table_slice(const PlanView& P, const TailStrategy& S, const T* __restrict__ xr, T* __restrict__ tbl,
                             int64_t q, int64_t qe) {
  if (q >= qe) return;
  int32_t cmb[ST_MAX_RANK];
  comb_unrank(P.binom, P.rank, q, S.Rt, S.tau, cmb);
  while (true) {
    T p = xr[cmb[0]];
    for (int k = 1; k < S.tau; ++k) p *= xr[cmb[k]];
    tbl[q] = p;
    if (++q >= qe) break;
    int j = S.tau - 1;
    while (cmb[j] + 1 > S.Rt - (S.tau - j)) --j;
    ++cmb[j];
    for (int k = j + 1; k < S.tau; ++k) cmb[k] = cmb[k - 1] + 1;
  }
}

// Level-by-level table build.  The t-combinations of range(Rt) in lexicographic order are grouped by their
// first value c0, and the group of c0 is xr[c0] times the SUFFIX of the (t-1)-table that holds the
// combinations whose first value exceeds c0 (a contiguous run: lexicographic order again):
//     T_t[C(Rt,t) - C(Rt-c0,t) + i] = xr[c0] * T_{t-1}[C(Rt,t-1) - C(Rt-1-c0,t-1) + i],  0 <= i < C(Rt-1-c0,t-1)
// with T_1 = xr.  One level is a set of independent, coalesced scaled copies: thread `tid` of `nthreads` takes
// the groups warp by warp.  The caller separates levels with a barrier.  src/dst must not overlap.
template <typename T>
ST_HD void build_table_level(const int64_t* __restrict__ bt, int rk, int Rt, int t, const T* __restrict__ xr, const T* __restrict__ src,
                             T* __restrict__ dst, int tid, int nthreads) {
  const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
  const int64_t nt = binom_at(bt, rk, Rt, t), nt1 = binom_at(bt, rk, Rt, t - 1);
  for (int c0 = warp; c0 + t <= Rt; c0 += nw) {
    const int64_t go = nt - binom_at(bt, rk, Rt - c0, t);
    const int64_t len = binom_at(bt, rk, Rt - 1 - c0, t - 1);
    const int64_t so = nt1 - len;
    const T xv = xr[c0];
    for (int64_t i = lane; i < len; i += 32) dst[go + i] = xv * src[so + i];
  }
}

// Suffix of the pair table (hybrid pair walk): rows k >= k0, entry (k, l) at  rs(k) - rs(k0) + l - k - 1.
// Thread `tid` of `nthreads` takes the rows warp by warp; coalesced stores.
template <typename T>
ST_HD void build_pair_suffix(int Rt, int k0, const T* __restrict__ xr, T* __restrict__ dst, int tid, int nthreads) {
  const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
  const int qs = k0 * Rt - k0 * (k0 + 1) / 2;
  for (int k = k0 + warp; k < Rt - 1; k += nw) {
    const int go = k * Rt - k * (k + 1) / 2 - qs - k - 1;
    const T xv = xr[k];
    for (int l = k + 1 + lane; l < Rt; l += 32) dst[go + l] = xv * xr[l];
  }
}

// Scratch (beyond the C(Rt, tau) entries of the table itself) of the level-by-level build: levels t < tau
// alternate between two buffers behind the table, A (tau - t odd) and B (tau - t even).
ST_HD void table_scratch(const int64_t* bt, int rk, int Rt, int tau, int64_t* nA, int64_t* nB) {
  *nA = 0;
  *nB = 0;
  for (int t = 2; t < tau; ++t) {
    const int64_t c = binom_at(bt, rk, Rt, t);
    if ((tau - t) & 1) { if (c > *nA) *nA = c; }
    else { if (c > *nB) *nB = c; }
  }
}

// source / destination of level t inside the buffer `tbl` (table first, then scratch A, then scratch B)
template <typename T>
ST_HD T* table_level_buffer(T* tbl, int64_t tbl_n, int64_t nA, int tau, int t) {
  if (t == tau) return tbl;
  return ((tau - t) & 1) ? tbl + tbl_n : tbl + tbl_n + nA;
}

}  // namespace st
