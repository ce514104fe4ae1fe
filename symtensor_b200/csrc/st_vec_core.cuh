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

template <typename T>
ST_HD T ld_stream(const T* p) {
#ifdef __CUDA_ARCH__
  return __ldcs(p);
#else
  return *p;
#endif
}

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
// Shared memory: [T: tbl_cap x T][xr: dim x T][xs: dim x T (copy of x)][private xr: nwarps x dim x T][ctrl]
struct TailCtrl {
  double red[32];
  double wE;               // gamma * prod over earlier runs x[v]^m
  int32_t E[ST_MAX_RANK];  // values of the earlier runs, ascending
  int32_t cur_cls;
  int64_t cur_seg;
  long long item;          // work item broadcast (dynamic scheduling)
};

// earlier runs of segment `sidx`: values (ascending, in E) and the weight gamma * prod x[v]^m
template <typename T>
ST_HD double unrank_earlier(const PlanView& P, const ClassDesc& C, int64_t sidx, const T* __restrict__ xs,
                                                 int32_t* E) {
  double w = (double)C.gamma;
  if (C.nruns <= 1) return w;
  int32_t vals[ST_MAX_RANK];
  int64_t dig[ST_MAX_RANK];
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

// ------------------------------------------------------------------------------------------------------
// walk_range: one warp's walk over segment positions [q0, q1) of one segment.
//
// Memory and index arithmetic are DECOUPLED.  On the device the components are streamed through a per-warp
// shared-memory ring with cp.async (16-byte copies, NST stages of 1 KB, NST-1 in flight while one is
// consumed) in fixed stages that ignore block boundaries; the walk is handed one 32-wide SLOT of consecutive
// components at a time and multiplies it against the "pieces" (block ∩ range) that overlap it:
//     weight(idx) = hw * tbl[toff + idx]   for idx in [pb, pe)
// A warp-uniform odometer produces the pieces: the common step -- increment the last head value -- costs a
// handful of instructions (fast path); carries take the general path over the local arrays u[] / pref[].
// All hot state is in scalar locals (registers); only u[] and pref[] are indexed dynamically.
// `tbl` is the tail table (T for tau >= 2, xr for tau == 1), `xr` the head-factor table, `blen` the optional
// table  blen[u] = C(Rt-1-u, tau)  of block lengths (nullptr: computed from the binomial table).
// STAGED = false reads the components directly (CPU emulation in tests/emu, which replays this code lane by
// lane; on the device it serves as the reference path for tiny ranges).
// ------------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define ST_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define ST_ASSUME_SHARED(p)
#endif
#ifdef __CUDACC__
__device__ __forceinline__ void st_cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void st_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void st_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

template <typename T, int NST, bool STAGED>
ST_HD double walk_range(const PlanView& P, const TailStrategy& S, const T* __restrict__ tbl, const T* __restrict__ xr,
                        const int32_t* __restrict__ blen, double wE, const T* __restrict__ Aseg, int64_t q0, int64_t q1, int lane,
                        T* ring, const int32_t* u_init) {
  ST_ASSUME_SHARED(tbl);
  ST_ASSUME_SHARED(xr);
  const int64_t* __restrict__ bt = P.binom;
  const int rk = P.rank;
  const int gt = S.gt, hn = S.hn, tau = S.tau, Rt = S.Rt;
  const int tbl_n = (int)S.tbl_n;
  const int n = (int)(q1 - q0);
  int32_t u[ST_MAX_RANK];        // head combination (relabelled values), warp-uniform; slow path only
  double pref[ST_MAX_RANK + 1];  // pref[i] = wE * prod_{j<i} xr[u[j]];                 slow path only
  if (u_init) { for (int i = 0; i < gt; ++i) u[i] = u_init[i]; }
  else comb_unrank(bt, rk, q0, Rt, gt, u);
  const int tq = (int)comb_rank(bt, rk, u + hn, Rt, tau);  // table index of the first tail
  pref[0] = wE;
  for (int i = 0; i < hn; ++i) pref[i + 1] = pref[i] * (double)xr[u[i]];
  int u_last = hn ? u[hn - 1] : -1;           // == u[hn-1]
  double pref_prev = hn ? pref[hn - 1] : wE;  // == pref[hn-1]
  double hw = pref[hn];
  int pb = 0;                                  // current piece [pb, pe), positions relative to q0
  int pe = (tbl_n - tq < n) ? tbl_n - tq : n;
  int toff = tq;
  T s = T(0);          // per-lane partial of the current piece
  double total = 0.0;  // per-lane running total

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
  // consume the slot [base, slot_end): this lane holds component base + lane in v (0 beyond slot_end)
  auto slot = [&](int base, int slot_end, T v) {
    const int idx = base + lane;
    while (true) {
      const bool in = (idx >= pb) & (idx < pe);
      const T tv = tbl[toff + (in ? idx : pb)];
      s += (in ? v : T(0)) * tv;
      if (pe >= slot_end) break;  // the piece covers the rest of the slot
      total += hw * (double)s;
      s = T(0);
      advance();
    }
  };

  const T* __restrict__ ap = Aseg + q0;
  bool staged = false;
#ifdef __CUDA_ARCH__
  staged = STAGED;
  if (STAGED) {
    constexpr int VEC = 16 / (int)sizeof(T);  // components per 16-byte vector
    constexpr int STAGE_E = 2 * 32 * VEC;     // a stage is 1 KB: two 16-byte vectors per lane
    // consume one stage [base, slot_end) held at shared address `sp`
    auto vslot = [&](const T* __restrict__ sp, int base, int slot_end) {
      if (pb <= base && pe >= base + STAGE_E && slot_end == base + STAGE_E) {
        // warp-uniform fast path: one piece covers the whole stage; each lane multiplies two 16-byte vectors
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int off = h * 32 * VEC + lane * VEC;
          const float4 raw = *reinterpret_cast<const float4*>(sp + off);
          const T* __restrict__ tp = tbl + (toff + base + off);
          if (sizeof(T) == 8) {
            s += (T)__hiloint2double(__float_as_int(raw.y), __float_as_int(raw.x)) * tp[0];
            s += (T)__hiloint2double(__float_as_int(raw.w), __float_as_int(raw.z)) * tp[VEC - 1];
          } else {
            s += (T)raw.x * tp[0];
            s += (T)raw.y * tp[1 % VEC];
            s += (T)raw.z * tp[2 % VEC];
            s += (T)raw.w * tp[3 % VEC];
          }
        }
        return;
      }
      // general path, piece-major: lanes stride the components of each piece that overlaps the stage
      while (true) {
        const int lo = pb > base ? pb : base, hi = pe < slot_end ? pe : slot_end;
        const T* __restrict__ tp = tbl + toff;
        for (int idx = lo + lane; idx < hi; idx += 32) s += sp[idx - base] * tp[idx];
        if (pe >= slot_end) break;
        total += hw * (double)s;
        s = T(0);
        advance();
      }
    };
    int a0 = (int)(((16u - (unsigned)((uintptr_t)ap & 15u)) & 15u) / sizeof(T));  // components before 16-byte alignment
    if (a0 > n) a0 = n;
    if (a0 > 0) slot(0, a0, lane < a0 ? ld_stream(ap + lane) : T(0));
    const int nb = (n - a0) / VEC * VEC;  // body: whole 16-byte vectors
    const int nchunks = (nb + STAGE_E - 1) / STAGE_E;
    const int body_end = a0 + nb;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring) + lane * 16;
    const char* __restrict__ gp = reinterpret_cast<const char*>(ap + a0 + lane * VEC);  // this lane's next vector to fetch
    int e_next = lane * VEC;                                                             // its component index in the body
    auto fetch = [&](int stage) {
      if (e_next < nb) st_cp_async16(ring_s + stage * 1024, gp);
      if (e_next + 32 * VEC < nb) st_cp_async16(ring_s + stage * 1024 + 512, gp + 512);
      st_cp_async_commit();
      gp += 1024;
      e_next += STAGE_E;
    };
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) fetch(c);
    int st = 0;           // stage holding chunk c
    int st_in = NST - 1;  // stage receiving chunk c + NST - 1
    int cbase = a0;
    for (int c = 0; c < nchunks; ++c) {
      fetch(st_in);
      st_cp_async_wait<NST - 1>();
      __syncwarp();
      const int cend = (cbase + STAGE_E < body_end) ? cbase + STAGE_E : body_end;
      vslot(ring + st * STAGE_E, cbase, cend);
      __syncwarp();  // the stage is overwritten by the copy issued in the next iteration
      st = (st + 1 == NST) ? 0 : st + 1;
      st_in = (st_in + 1 == NST) ? 0 : st_in + 1;
      cbase += STAGE_E;
    }
    if (body_end < n) slot(body_end, n, body_end + lane < n ? ld_stream(ap + body_end + lane) : T(0));
  }
#endif
  if (!staged) {
    for (int base = 0; base < n; base += 32) {
      const int idx = base + lane;
      slot(base, (base + 32 < n) ? base + 32 : n, idx < n ? ap[idx] : T(0));
    }
  }
  return total + hw * (double)s;
}

template <typename T>
ST_HD double walk_range_direct(const PlanView& P, const TailStrategy& S, const T* tbl, const T* xr, const int32_t* blen, double wE,
                               const T* Aseg, int64_t q0, int64_t q1, int lane) {
  return walk_range<T, 2, false>(P, S, tbl, xr, blen, wE, Aseg, q0, q1, lane, nullptr, nullptr);
}

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

// T[q] = prod of xr over the q-th tau-combination of range(Rt), for q in [q, qe): contiguous slice of one thread
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

}  // namespace st
