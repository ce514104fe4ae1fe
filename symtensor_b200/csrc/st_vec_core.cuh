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
// hn values are the "head", the last tau the "tail".  Two modes per class (host cost model):
//   mode A (tau >= 2)  tail weights come from the shared-memory table T (rebuilt by the CTA when E changes);
//   mode B (tau == 1)  the tail weight of relabelled value u is x[actual(u)]^mu, computed by the lane -- no
//                      table, no CTA synchronisation; used when segments are too short to amortise a table.
// Shared memory: [T table: tbl_cap x T][xr: dim x T (xrel^mu, mode A)][xs: dim x T (copy of x)][ctrl]
struct TailCtrl {
  double red[32];
  double wE;               // gamma * prod over earlier runs x[v]^m
  int32_t E[ST_MAX_RANK];  // values of the earlier runs, ascending
  int32_t cur_cls;
  int64_t cur_seg;
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

// One warp streams segment positions [q0, q1) of a segment whose first component is Aseg[0].
// MODE_A: tail weights from `tbl` (tau-combination table) and head factors from `xr`;
// MODE_B: tau == 1, weights computed from xs / E.
template <typename T, bool MODE_A>
ST_HD void walk_piece(const PlanView& P, const TailStrategy& S, const T* __restrict__ tbl,
                                           const T* __restrict__ xr, const T* __restrict__ xs, const int32_t* E, double wE,
                                           const T* __restrict__ Aseg, int64_t q0, int64_t q1, int lane, double& total) {
  const int64_t* bt = P.binom;
  const int rk = P.rank;
  const int gt = S.gt, hn = S.hn, tau = S.tau, Rt = S.Rt;
  int32_t u[ST_MAX_RANK];
  comb_unrank(bt, rk, q0, Rt, gt, u);
  int64_t tq = comb_rank(bt, rk, u + hn, Rt, tau);  // table index of the current tail
  double pref[ST_MAX_RANK + 1];                      // pref[i] = wE * prod_{j<i} xrel[u[j]]^mu
  pref[0] = wE;
  for (int i = 0; i < hn; ++i)
    pref[i + 1] = pref[i] * (double)(MODE_A ? xr[u[i]] : xrel_pow<T>(xs, E, S.nE, S.mu, u[i]));
  int64_t pos = q0;
  const int64_t tbl_n = S.tbl_n;
  while (true) {
    const int64_t left = q1 - pos;
    int64_t cnt = tbl_n - tq;  // the block ends where the table ends
    if (cnt > left) cnt = left;
    const T* __restrict__ ap = Aseg + pos;
    T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int64_t i = lane;
    if (MODE_A) {
      // coalesced dot product  <A[block], T[slice]>
      const T* __restrict__ tp = tbl + tq;
      for (; i + 96 < cnt; i += 128) {
        const T a0 = ld_stream(ap + i), a1 = ld_stream(ap + i + 32), a2 = ld_stream(ap + i + 64), a3 = ld_stream(ap + i + 96);
        s0 += a0 * tp[i];
        s1 += a1 * tp[i + 32];
        s2 += a2 * tp[i + 64];
        s3 += a3 * tp[i + 96];
      }
      for (; i < cnt; i += 32) s0 += ld_stream(ap + i) * tp[i];
    } else {
      for (; i + 32 < cnt; i += 64) {
        const T a0 = ld_stream(ap + i), a1 = ld_stream(ap + i + 32);
        s0 += a0 * xrel_pow<T>(xs, E, S.nE, S.mu, (int32_t)(tq + i));
        s1 += a1 * xrel_pow<T>(xs, E, S.nE, S.mu, (int32_t)(tq + i + 32));
      }
      for (; i < cnt; i += 32) s0 += ld_stream(ap + i) * xrel_pow<T>(xs, E, S.nE, S.mu, (int32_t)(tq + i));
    }
    total += pref[hn] * ((double)(s0 + s1) + (double)(s2 + s3));
    pos += cnt;
    if (pos >= q1) break;
    // next head (warp-uniform): increment u[hn-1], carrying while no room is left for the rest of the run
    int j = hn - 1;
    while (j >= 0 && u[j] + 1 > Rt - (gt - j)) --j;
    if (j < 0) break;  // defensive: cannot happen inside a segment
    ++u[j];
    for (int k = j + 1; k < hn; ++k) u[k] = u[k - 1] + 1;
    for (int k = j; k < hn; ++k)
      pref[k + 1] = pref[k] * (double)(MODE_A ? xr[u[k]] : xrel_pow<T>(xs, E, S.nE, S.mu, u[k]));
    // first tail of the new head is (b+1, b+2, ..), b = u[hn-1]: table index tbl_n - C(Rt-1-b, tau)
    tq = tbl_n - binom_at(bt, rk, Rt - 1 - u[hn - 1], tau);
  }
}


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
