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
struct TailCtrl {
  double red[32];
  double wE;               // gamma * prod over earlier runs x[v]^m
  int32_t E[ST_MAX_RANK];  // values of the earlier runs, ascending
  int32_t cur_cls;
  int64_t cur_seg;
  long long item[2];       // work item broadcast (dynamic scheduling), double-buffered
  int32_t last;            // this CTA is the last one to finish (fused finalize)
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

// one 16-byte vector of components (streaming load: every component is read exactly once)
template <typename T>
ST_HD void ld_vec16(const T* __restrict__ p, T* v) {
#ifdef __CUDA_ARCH__
  if (sizeof(T) == 8) {
    const double2 r = __ldcs(reinterpret_cast<const double2*>(p));
    v[0] = (T)r.x;
    v[1] = (T)r.y;
  } else {
    const float4 r = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = (T)r.x;
    v[1] = (T)r.y;
    v[2 % (16 / (int)sizeof(T))] = (T)r.z;
    v[3 % (16 / (int)sizeof(T))] = (T)r.w;
  }
#else
  for (int k = 0; k < 16 / (int)sizeof(T); ++k) v[k] = p[k];
#endif
}

// ------------------------------------------------------------------------------------------------------
// walk_range: one warp's walk over segment positions [q0, q1) of one segment (q1 - q0 < 2^30).
//
// Memory and index arithmetic are DECOUPLED.  The components are fetched in BATCHES of U slots (a slot is one
// 16-byte vector per lane: 512 contiguous bytes per warp), all U loads of a batch issued back to back into
// registers; two batches are kept (one being consumed, the next in flight) so a warp always has loads
// outstanding.  Batches start at the 16-byte boundary at or below the range start and ignore block boundaries.  The odometer hands out "pieces" (block ∩ range, positions relative to q0):
//     weight(idx) = hw * tbl[toff + idx]   for idx in [pb, pe)
// A batch inside one piece takes the fast path (per 16-byte vector: one LDG.128, VEC table loads, VEC FMAs);
// otherwise the batch is consumed piece by piece with per-element predicates.  The warp-uniform odometer's
// common step -- increment the last head value -- costs a handful of instructions; carries take the general
// path over the local arrays u[] / pref[].  `tbl` is the tail table (T for tau >= 2, xr for tau == 1), `xr`
// the head-factor table, `blen` the optional table blen[u] = C(Rt-1-u, tau) of block lengths (nullptr:
// computed from the binomial table).  `u_init` is the last run's combination at q0 when the caller knows it (tile
// directory, start of a segment); otherwise it is unranked here, after the first two batches have been requested.
// The same code runs on the host (tests/emu replays it lane by lane).
// ------------------------------------------------------------------------------------------------------
template <typename T, int U>
ST_HD double walk_range(const PlanView& P, const TailStrategy& S, const T* __restrict__ tbl, const T* __restrict__ xr,
                        const int32_t* __restrict__ blen, double wE, const T* __restrict__ Aseg, int64_t q0, int64_t q1, int lane,
                        const int32_t* u_init, T (&bufA)[U][16 / sizeof(T)], T (&bufB)[U][16 / sizeof(T)], bool preloaded,
                        const T* __restrict__ next_ap, WarpScratch& ws) {
  ST_ASSUME_SHARED(tbl);
  ST_ASSUME_SHARED(xr);
  constexpr int VEC = 16 / (int)sizeof(T);  // components per 16-byte vector
  constexpr int SLOT = 32 * VEC;            // components per slot
  constexpr int BATCH = U * SLOT;
  const int64_t* __restrict__ bt = P.binom;
  const int rk = P.rank;
  const int gt = S.gt, hn = S.hn, tau = S.tau, Rt = S.Rt;
  const int tbl_n = (int)S.tbl_n;
  const int n = (int)(q1 - q0);
  const T* __restrict__ ap = Aseg + q0;
  // components between the 16-byte boundary at or below `ap` and `ap` (they belong to whoever owns them: masked out)
  const int a_pre = (int)(((uintptr_t)ap & 15u) / sizeof(T));
  const T* __restrict__ lp = ap - a_pre + lane * VEC;  // this lane's vector of slot 0 of the batch at position -a_pre
  int pos = -a_pre;                                    // position (relative to q0) of the current batch
  // bufA / bufB: two batches in registers, one being consumed, one in flight.  They belong to the caller so that
  // the stream can continue ACROSS ranges: with `next_ap` (16-byte aligned start of the caller's next range, at
  // least two whole batches long; only legal when this range has an even number of batches) the first two
  // batches of the next range are requested as soon as the buffers free up, and the next call passes
  // `preloaded`.
  // fetch the batch at position p (whole vectors while they end inside the range, single components up to n,
  // zeros beyond: only the last batch of a range is partial)
  auto load = [&](T (&buf)[U][VEC], int p) {
    const T* __restrict__ bp = lp + (p + a_pre);
    if (p + BATCH <= n) {
#pragma unroll
      for (int s = 0; s < U; ++s) ld_vec16(bp + s * SLOT, buf[s]);
    } else {
#pragma unroll
      for (int s = 0; s < U; ++s) {
        const int i0 = p + s * SLOT + lane * VEC;
        if (i0 + VEC <= n) {
          ld_vec16(bp + s * SLOT, buf[s]);
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) buf[s][k] = (i0 + k < n && i0 + k >= 0) ? bp[s * SLOT + k] : T(0);
        }
      }
    }
  };
  if (!preloaded) {
    load(bufA, pos);
    if (pos + BATCH < n) load(bufB, pos + BATCH);
  }
  int nx = 0;  // batches of the next range requested so far
  auto load_next = [&](T (&buf)[U][VEC]) {
    const T* __restrict__ bp = next_ap + nx * BATCH + lane * VEC;
#pragma unroll
    for (int s = 0; s < U; ++s) ld_vec16(bp + s * SLOT, buf[s]);
    ++nx;
  };

  // ---- odometer state at q0
  int32_t* u = ws.u;       // head combination (relabelled values), warp-uniform; slow path only
  double* pref = ws.pref;  // pref[i] = wE * prod_{j<i} xr[u[j]];                 slow path only
  if (u_init) {
    for (int i = 0; i < gt; ++i) u[i] = u_init[i];  // from the tile directory / the start of a segment
  } else {
#ifdef __CUDA_ARCH__
    comb_unrank_warp(bt, rk, q0, Rt, gt, u, lane);
#else
    comb_unrank(bt, rk, q0, Rt, gt, u);
#endif
  }
  const int tq = (int)comb_rank(bt, rk, u + hn, Rt, tau);  // table index of the first tail
  pref[0] = wE;
  for (int i = 0; i < hn; ++i) pref[i + 1] = pref[i] * (double)xr[u[i]];
  int u_last = hn ? u[hn - 1] : -1;           // == u[hn-1]
  double pref_prev = hn ? pref[hn - 1] : wE;  // == pref[hn-1]
  double hw = pref[hn];
  int pb = 0;                                  // current piece [pb, pe), positions relative to q0
  int pe = (tbl_n - tq < n) ? tbl_n - tq : n;
  int toff = tq;
  T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);  // per-lane partials of the current piece (independent FMA chains)
  double total = 0.0;                            // per-lane running total

  // next block: lexicographic successor of the head combination
  auto advance = [&]() {
    pb = pe;
    if (hn == 0) { pe = n; return; }  // defensive: a segment with hn == 0 is a single block
    if (u_last + 1 <= Rt - tau - 1) {
      ++u_last;  // fast path: no carry
    } else {
      u[hn - 1] = u_last;
      int j = hn - 1;
      while (j >= 0 && u[j] + 1 > Rt - (gt - j)) --j;
      if (j < 0) { pe = n; return; }  // defensive: cannot happen inside a segment
      ++u[j];
      for (int k = j + 1; k < hn; ++k) u[k] = u[k - 1] + 1;
      for (int k = j; k < hn - 1; ++k) pref[k + 1] = pref[k] * (double)xr[u[k]];
      u_last = u[hn - 1];
      pref_prev = pref[hn - 1];
    }
    hw = pref_prev * (double)xr[u_last];
    // first tail of the new head is (b+1, b+2, ..), b = u_last: block length C(Rt-1-b, tau), at the table's end
    const int bl = tau == 1 ? Rt - 1 - u_last : (blen ? blen[u_last] : (int)binom_at(bt, rk, Rt - 1 - u_last, tau));
    toff = tbl_n - bl - pb;
    pe = (bl < n - pb) ? pb + bl : n;
  };
  auto flush = [&]() {
    total += hw * (((double)s0 + (double)s1) + ((double)s2 + (double)s3));
    s0 = T(0);
    s1 = T(0);
    s2 = T(0);
    s3 = T(0);
  };
  // slot s of a batch, entirely inside the current piece: no predicates
  auto slot_full = [&](const T (&v)[VEC], const T* __restrict__ tp, int s) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const T t = tp[k];
      switch ((s * VEC + k) & 3) {
        case 0: s0 += v[k] * t; break;
        case 1: s1 += v[k] * t; break;
        case 2: s2 += v[k] * t; break;
        default: s3 += v[k] * t; break;
      }
    }
  };

  // consume the batch at `pos`, valid up to bend = min(pos + BATCH, n); invariant on entry: pb <= max(pos, 0) < pe
  auto consume = [&](T (&cur)[U][VEC]) {
    const int bend = (pos + BATCH < n) ? pos + BATCH : n;
    if (pb <= pos && pe >= pos + BATCH) {
      // one piece covers the whole batch
      const T* __restrict__ tp = tbl + (toff + pos + lane * VEC);
#pragma unroll
      for (int s = 0; s < U; ++s) slot_full(cur[s], tp + s * SLOT, s);
      if (pe == bend) {
        flush();
        if (pe < n) advance();
      }
      return;
    }
    while (true) {
      // the part of the current piece inside this batch: slots it covers entirely take the plain path, the
      // (at most two) slots holding its ends are predicated per component
#pragma unroll
      for (int s = 0; s < U; ++s) {
        const int sb = pos + s * SLOT;
        if (pe > sb && pb < sb + SLOT) {  // the piece overlaps this slot (warp-uniform)
          if (pb <= sb && pe >= sb + SLOT) {
            slot_full(cur[s], tbl + (toff + sb + lane * VEC), s);
          } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const int idx = sb + lane * VEC + k;
              const bool in = (idx >= pb) & (idx < pe);
              const T tv = tbl[toff + (in ? idx : pb)];
              s0 += (in ? cur[s][k] : T(0)) * tv;
            }
          }
        }
      }
      if (pe > bend) break;  // the piece continues in the next batch
      flush();
      if (pe >= n) break;    // end of the range
      advance();
      if (pb >= bend) break; // the next piece starts with the next batch
    }
  };

  while (true) {
    // tight run: while the current piece covers the two batches in registers and the two batches after them are
    // whole, stream with no per-batch bookkeeping (per 16-byte vector: one LDG.128, VEC table loads, VEC FMAs)
    if (pb <= pos && pos + 2 * BATCH <= pe && pos + 4 * BATCH <= n) {
      const T* __restrict__ tp = tbl + (toff + pos + lane * VEC);
      const T* __restrict__ gp = lp + (pos + a_pre) + 2 * BATCH;  // this lane's vector of the batch after next
      int run = (pe - pos) / (2 * BATCH);
      const int run_n = (n - pos - 2 * BATCH) / (2 * BATCH);
      if (run_n < run) run = run_n;
      pos += run * (2 * BATCH);
      for (; run > 0; --run) {
#pragma unroll
        for (int s = 0; s < U; ++s) slot_full(bufA[s], tp + s * SLOT, s);
#pragma unroll
        for (int s = 0; s < U; ++s) ld_vec16(gp + s * SLOT, bufA[s]);
#pragma unroll
        for (int s = 0; s < U; ++s) slot_full(bufB[s], tp + BATCH + s * SLOT, s);
#pragma unroll
        for (int s = 0; s < U; ++s) ld_vec16(gp + BATCH + s * SLOT, bufB[s]);
        tp += 2 * BATCH;
        gp += 2 * BATCH;
      }
      if (pe == pos) {  // the piece ended exactly here (pos < n: two more batches were loaded)
        flush();
        advance();
      }
    }
    consume(bufA);
    pos += BATCH;
    if (pos + BATCH < n) load(bufA, pos + BATCH);
    else if (next_ap != nullptr && nx < 2) load_next(bufA);
    if (pos >= n) break;
    consume(bufB);
    pos += BATCH;
    if (pos + BATCH < n) load(bufB, pos + BATCH);
    else if (next_ap != nullptr && nx < 2) load_next(bufB);
    if (pos >= n) break;
  }
  return total + hw * (((double)s0 + (double)s1) + ((double)s2 + (double)s3));
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
