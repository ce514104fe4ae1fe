// Shared definitions of the symtensor_b200 CUDA library: class descriptors, the per-(rank, dim) plan and
// the combinatorial rank / unrank routines (usable on host and device).
//
// The storage order implemented here is the reference's, bit for bit:
//   class order      symtensor/utils.py:839-856, 1000-1002        (partitions, descending lexicographic)
//   order in a class symtensor/permcls_symtensor.py:288-347       (σindex_iter)
//   representative   symtensor/permcls_symtensor.py:375-381       (get_index_representative)
//   flat order       symtensor/flat_symtensor.py:39-50, 219-220   (combinations_with_replacement)
// in the closed form of SURVEY.md A.2: a class is a sequence of runs of equal multiplicity; run j is an
// increasing g_j-combination of the R_j values not used by earlier runs (relabelled by their order among
// the unused values); the position is the mixed-radix number of the runs' lexicographic ranks.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/symtensor_b200.h"

#define ST_HD __host__ __device__ __forceinline__

namespace st {

struct ClassDesc {
  int32_t nvals;                 // l: distinct index values of a component
  int32_t nruns;                 // t: runs of equal multiplicity
  int32_t mult[ST_MAX_RANK];     // multiplicity m_k of value position k (descending)
  int32_t run_start[ST_MAX_RANK];  // first value position of run j
  int32_t run_len[ST_MAX_RANK];    // g_j
  int32_t run_mult[ST_MAX_RANK];   // multiplicity shared by run j
  int64_t radix[ST_MAX_RANK];      // C(R_j, g_j), R_j = dim - (values used by earlier runs)
  int64_t size;                  // prod_j radix[j] (0 if l > dim)
  int64_t offset;                // start in the ST_LAYOUT_PERMCLS buffer
  int64_t gamma;                 // rank!/prod(m_k!)
};

// Device-visible part of a plan.  binom[n * (rank + 1) + k] = C(n, k) for 0 <= n <= dim + rank, saturated at
// INT64_MAX (only entries that index existing storage are ever compared against positions).
struct PlanView {
  int32_t rank;
  int32_t ncls;
  int64_t dim;
  int64_t total;           // padded length of the permcls buffer
  int64_t flat_size;       // C(dim + rank - 1, rank)
  const ClassDesc* cls;    // [ncls]
  const int64_t* offsets;  // [ncls + 1]
  const int64_t* binom;    // [(dim + rank + 1) * (rank + 1)]
};

ST_HD int64_t binom_at(const int64_t* tbl, int rank, int64_t n, int k) {
  // callers guarantee 0 <= k <= rank and n <= dim + rank; negative n means "no values left"
  return n < 0 ? 0 : tbl[n * (rank + 1) + k];
}

// Lexicographic rank of the increasing combination c[0..g) of range(n) among the C(n, g) combinations.
ST_HD int64_t comb_rank(const int64_t* tbl, int rank, const int32_t* c, int64_t n, int g) {
  int64_t r = binom_at(tbl, rank, n, g) - 1;
  for (int k = 0; k < g; ++k) r -= binom_at(tbl, rank, n - 1 - c[g - 1 - k], k + 1);
  return r;
}

// Inverse of comb_rank.  Element i is found by binary search: the number of combinations whose i-th element
// is smaller than v (given the previous element) is C(n-1-prev, k) - C(n-v, k), k = g - i.
ST_HD void comb_unrank(const int64_t* tbl, int rank, int64_t r, int64_t n, int g, int32_t* c) {
  int64_t prev = -1;
  for (int i = 0; i < g; ++i) {
    const int k = g - i;
    const int64_t all = binom_at(tbl, rank, n - 1 - prev, k);
    int64_t lo = prev + 1, hi = n - k;
    while (lo < hi) {
      const int64_t mid = (lo + hi + 1) >> 1;
      if (all - binom_at(tbl, rank, n - mid, k) <= r) lo = mid; else hi = mid - 1;
    }
    r -= all - binom_at(tbl, rank, n - lo, k);
    c[i] = (int32_t)lo;
    prev = lo;
  }
}

// position inside a class  ->  the l distinct values in class order
ST_HD void permcls_unrank_vals(const PlanView& P, const ClassDesc& C, int64_t pos, int32_t* vals) {
  int64_t dig[ST_MAX_RANK];
  for (int j = C.nruns - 1; j >= 0; --j) {
    const int64_t q = pos / C.radix[j];
    dig[j] = pos - q * C.radix[j];
    pos = q;
  }
  int32_t used[ST_MAX_RANK];  // values of earlier runs, ascending
  int nused = 0;
  for (int j = 0; j < C.nruns; ++j) {
    const int g = C.run_len[j], s = C.run_start[j];
    comb_unrank(P.binom, P.rank, dig[j], P.dim - nused, g, vals + s);
    for (int i = 0; i < g; ++i) {  // undo the relabelling: skip values taken by earlier runs
      int32_t v = vals[s + i];
      for (int u = 0; u < nused; ++u) v += (v >= used[u]);
      vals[s + i] = v;
    }
    for (int i = 0; i < g; ++i) {  // merge the run into `used` (insertion keeps it ascending)
      const int32_t v = vals[s + i];
      int u = nused++;
      while (u > 0 && used[u - 1] > v) { used[u] = used[u - 1]; --u; }
      used[u] = v;
    }
  }
}

// the l distinct values in class order  ->  position inside the class
ST_HD int64_t permcls_rank_vals(const PlanView& P, const ClassDesc& C, const int32_t* vals) {
  int32_t used[ST_MAX_RANK];
  int nused = 0;
  int64_t pos = 0;
  for (int j = 0; j < C.nruns; ++j) {
    const int g = C.run_len[j], s = C.run_start[j];
    int32_t rel[ST_MAX_RANK];
    for (int i = 0; i < g; ++i) {
      const int32_t v = vals[s + i];
      int32_t below = 0;
      for (int u = 0; u < nused; ++u) below += (used[u] < v);
      rel[i] = v - below;
    }
    pos = pos * C.radix[j] + comb_rank(P.binom, P.rank, rel, P.dim - nused, g);
    for (int i = 0; i < g; ++i) {
      const int32_t v = vals[s + i];
      int u = nused++;
      while (u > 0 && used[u - 1] > v) { used[u] = used[u - 1]; --u; }
      used[u] = v;
    }
  }
  return pos;
}

// Arbitrary multi-index (rank entries, any order) -> class ordinal + the values in class order.
// Groups equal values, orders groups by count (descending), ties by ascending value: the reference's
// get_index_representative.  Returns -1 if an entry is outside [0, dim).
ST_HD int classify_index(const PlanView& P, const int32_t* idx, int32_t* vals) {
  const int r = P.rank;
  int32_t s[ST_MAX_RANK];
  for (int i = 0; i < r; ++i) {
    const int32_t v = idx[i];
    if (v < 0 || v >= P.dim) return -1;
    int u = i;
    while (u > 0 && s[u - 1] > v) { s[u] = s[u - 1]; --u; }
    s[u] = v;
  }
  int32_t gv[ST_MAX_RANK], gc[ST_MAX_RANK];
  int ng = 0;
  for (int i = 0; i < r; ++i) {
    if (ng && gv[ng - 1] == s[i]) ++gc[ng - 1];
    else { gv[ng] = s[i]; gc[ng] = 1; ++ng; }
  }
  for (int i = 1; i < ng; ++i) {  // stable insertion sort by count, descending
    const int32_t v = gv[i], c = gc[i];
    int u = i;
    while (u > 0 && gc[u - 1] < c) { gv[u] = gv[u - 1]; gc[u] = gc[u - 1]; --u; }
    gv[u] = v; gc[u] = c;
  }
  for (int i = 0; i < ng; ++i) vals[i] = gv[i];
  for (int c = 0; c < P.ncls; ++c) {
    const ClassDesc& C = P.cls[c];
    if (C.nvals != ng) continue;
    bool same = true;
    for (int i = 0; i < ng; ++i) same = same && (C.mult[i] == gc[i]);
    if (same) return c;
  }
  return -1;
}

// flat layout: sorted multi-index i_0 <= ... <= i_{r-1}  <->  strict combination i_k + k of range(dim + r - 1)
ST_HD int64_t flat_rank_sorted(const PlanView& P, const int32_t* s) {
  const int r = P.rank;
  int64_t pos = P.flat_size - 1;
  for (int k = 0; k < r; ++k) pos -= binom_at(P.binom, r, P.dim - 1 + k - s[r - 1 - k], k + 1);
  return pos;
}

ST_HD void flat_unrank_sorted(const PlanView& P, int64_t pos, int32_t* s) {
  comb_unrank(P.binom, P.rank, pos, P.dim + P.rank - 1, P.rank, s);
  for (int k = 0; k < P.rank; ++k) s[k] -= k;
}

// flat position of the sorted r-tuple s (r <= plan rank), using the plan's binomial table
ST_HD int64_t flat_rank_r(const PlanView& P, const int32_t* s, int r) {
  int64_t pos = binom_at(P.binom, P.rank, P.dim + r - 1, r) - 1;
  for (int k = 0; k < r; ++k) pos -= binom_at(P.binom, P.rank, P.dim - 1 + k - s[r - 1 - k], k + 1);
  return pos;
}

ST_HD void flat_unrank_r(const PlanView& P, int64_t pos, int r, int32_t* s) {
  comb_unrank(P.binom, P.rank, pos, P.dim + r - 1, r, s);
  for (int k = 0; k < r; ++k) s[k] -= k;
}

// Successor of a component in its class's storage order (the sigma-index order, lexicographic in the runs' combinations, the
// last run varying fastest): `vals` (class order) is advanced in place; false at the end of the class.  An odometer step --
// a few dozen instructions instead of the divisions and binary searches of a full unrank -- for kernels that walk runs of
// consecutive coordinates.  Run j is an increasing combination of the values not used by runs 0 .. j-1.
// skip: the last `skip` values of the last run are treated as exhausted, i.e. the result is the first component of the next
// ROW (the stretch of consecutive components that differ only in those values).
ST_HD bool permcls_advance(const PlanView& P, const ClassDesc& C, int32_t* vals, int skip) {
  const int32_t d = (int32_t)P.dim;
  for (int j = C.nruns - 1; j >= 0; --j) {
    const int s = C.run_start[j], g = C.run_len[j];
    for (int i = g - 1 - (j == C.nruns - 1 ? skip : 0); i >= 0; --i) {
      const int32_t a = vals[s + i];
      int32_t avail = d - 1 - a;  // values above a that earlier runs have not taken
      for (int u = 0; u < s; ++u) avail -= (vals[u] > a);
      if (avail < g - i) continue;  // position i cannot move: no room for it and the positions after it
      // advance position i, then the positions after it take the next free values
      int32_t v = a;
      for (int q = i; q < g; ++q) {
        bool taken;
        do {
          ++v;
          taken = false;
          for (int u = 0; u < s; ++u) taken = taken || (vals[u] == v);
        } while (taken);
        vals[s + q] = v;
      }
      // later runs restart at their first combination: the smallest values not used before them
      for (int jj = j + 1; jj < C.nruns; ++jj) {
        const int s2 = C.run_start[jj], g2 = C.run_len[jj];
        int32_t w = -1;
        for (int q = 0; q < g2; ++q) {
          bool taken;
          do {
            ++w;
            taken = false;
            for (int u = 0; u < s2; ++u) taken = taken || (vals[u] == w);
          } while (taken);
          vals[s2 + q] = w;
        }
      }
      return true;
    }
  }
  return false;
}
ST_HD bool permcls_next_vals(const PlanView& P, const ClassDesc& C, int32_t* vals) { return permcls_advance(P, C, vals, 0); }

// class containing the packed coordinate c (offsets ascending); padding belongs to the class before it
ST_HD int class_of_coord(const PlanView& P, int64_t c) {
  int lo = 0, hi = P.ncls - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.offsets[mid] <= c) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---- row walk: a warp-uniform cursor over consecutive packed coordinates ----------------------------------------------
// A ROW is the stretch of consecutive components of a class that differ only in the last tau = min(g, 3) values of the last
// run (g values): those run through the increasing tau-combinations of the m values above the run's previous value that the
// earlier runs have not taken, in lexicographic order -- C(m, tau) components, ~300 on average at rank 8 dim 40.  A warp that
// serves 32 consecutive coordinates walks the (few) rows that cover them with uniform odometer steps (permcls_advance) and
// every lane decodes its own component from its offset in its row: no per-component unrank of the whole multi-index.  What
// a lane latches: the row's values in class order packed in bytes (hence dim <= 255), the previous value b, m and its offset.
constexpr int kRowTauMax = 3;
constexpr int kRowBinomStride = 256;  // C(n, 2) and C(n, 3) for n < 256: the lane-side table [2][256]

struct RowCursor {
  int32_t ci;        // class of `cur` (ncls: past the end)
  int32_t off;       // position of `cur` inside its row
  int64_t cur;       // first coordinate not handed out yet
  int64_t cls_end;   // end of class ci's components; [cls_end, next_off) is alignment padding
  int64_t next_off;  // first coordinate of class ci + 1
  int32_t vals[ST_MAX_RANK];  // the row's component values in class order (the last tau: any component of the row)
  unsigned long long valsp;   // ... packed one per byte (0xff fill), refreshed whenever the row changes
  int32_t b, m, len;          // of the row at the cursor (rowcursor_refresh): previous value, free values above it, C(m, tau) components
};

struct RowLatch {
  unsigned long long valsp;  // values in class order, one per byte (the tail bytes are the lane's to fill)
  int32_t b, m, o, ci;
  int32_t state;  // 0: not served (coordinates behind a class's padding), 1: component, 2: alignment padding
};

// per-class constants of the lane-side decode, packed: bits 0..31 the value position of every index entry (4 bits each,
// entries beyond the rank -> 7, whose byte is 0xff whenever the rank is below 8), 32..35 tau, 36..39 the number of values,
// 40..43 the number of values of earlier runs
ST_HD unsigned long long row_class_info(const ClassDesc& C, int rank) {
  unsigned long long e2v = 0;
  int e = 0;
  for (int v = 0; v < C.nvals; ++v)
    for (int m = 0; m < C.mult[v]; ++m) e2v |= (unsigned long long)v << (4 * e++);
  for (; e < 8; ++e) e2v |= 7ull << (4 * e);
  const int g = C.nruns ? C.run_len[C.nruns - 1] : 0;
  const int tau = g < kRowTauMax ? g : kRowTauMax;
  (void)rank;
  return e2v | ((unsigned long long)tau << 32) | ((unsigned long long)C.nvals << 36) | ((unsigned long long)(C.nruns ? C.run_start[C.nruns - 1] : 0) << 40);
}

ST_HD void rowcursor_enter_class(const PlanView& P, RowCursor& rc, int ci) {
  rc.ci = ci;
  rc.off = 0;
  if (ci >= P.ncls) { rc.cur = rc.cls_end = rc.next_off = INT64_MAX; return; }
  const ClassDesc& C = P.cls[ci];
  rc.cur = C.offset;
  rc.cls_end = C.offset + C.size;
  rc.next_off = P.offsets[ci + 1];
  for (int i = 0; i < C.nvals; ++i) rc.vals[i] = i;  // first component: every run takes the smallest values left
  rc.valsp = ~0ull; rc.b = -1; rc.m = 0; rc.len = 0;  // (rowcursor_refresh: by the caller, once the cursor is final)
}

// the row at the cursor: previous value b of the last run (-1 if the tail is the whole run), the number m of free values
// above it; returns the row's length C(m, tau)
ST_HD int32_t rowcursor_row(const PlanView& P, const RowCursor& rc, int32_t* b_out, int32_t* m_out) {
  const ClassDesc& C = P.cls[rc.ci];
  const int s = C.run_start[C.nruns - 1], g = C.run_len[C.nruns - 1];
  const int tau = g < kRowTauMax ? g : kRowTauMax;
  const int32_t b = g > tau ? rc.vals[C.nvals - tau - 1] : -1;
  int32_t m = (int32_t)P.dim - 1 - b;
  for (int u = 0; u < s; ++u) m -= (rc.vals[u] > b);
  *b_out = b;
  *m_out = m;
  return (int32_t)binom_at(P.binom, P.rank, m, tau);
}

ST_HD void rowcursor_pack(const PlanView& P, RowCursor& rc) {
  const ClassDesc& C = P.cls[rc.ci];
  unsigned long long p = ~0ull;
  for (int i = 0; i < C.nvals; ++i) p = (p & ~(0xffull << (8 * i))) | ((unsigned long long)rc.vals[i] << (8 * i));
  rc.valsp = p;
}

// b, m, len and the packed values of the row at the cursor: once per row, so that a batch inside a long row costs the walk a
// dozen instructions
ST_HD void rowcursor_refresh(const PlanView& P, RowCursor& rc) {
  if (rc.cur >= rc.cls_end) { rc.b = -1; rc.m = 0; rc.len = 0; return; }
  int32_t b, m;
  const int32_t len = rowcursor_row(P, rc, &b, &m);
  rc.b = b;
  rc.m = m;
  rc.len = len;
  rowcursor_pack(P, rc);
}

ST_HD void rowcursor_seek(const PlanView& P, RowCursor& rc, int64_t c) {
  const int ci = class_of_coord(P, c);
  rowcursor_enter_class(P, rc, ci);
  rc.cur = c;
  if (c >= rc.cls_end) { rowcursor_refresh(P, rc); return; }
  const ClassDesc& C = P.cls[ci];
  permcls_unrank_vals(P, C, c - C.offset, rc.vals);
  // position inside the row: the lexicographic rank of the tail among the combinations of the free values above b
  const int s = C.run_start[C.nruns - 1], g = C.run_len[C.nruns - 1];
  const int tau = g < kRowTauMax ? g : kRowTauMax;
  int32_t b, m;
  int32_t r = rowcursor_row(P, rc, &b, &m) - 1;
  for (int k = 0; k < tau; ++k) {
    const int32_t v = rc.vals[C.nvals - 1 - k];
    int32_t y = v - b - 1;  // its index among the free values above b
    for (int u = 0; u < s; ++u) y -= (rc.vals[u] > b && rc.vals[u] < v);
    r -= (int32_t)binom_at(P.binom, P.rank, m - 1 - y, k + 1);
  }
  rc.off = r;
  rowcursor_refresh(P, rc);
}

// Serve the coordinates below `batch_end`: every lane of a warp calls this with the same cursor (so the walk is uniform) and
// its own coordinate c, and leaves with the latch of the row that holds c.  The cursor ends at batch_end (inside a row if one
// straddles it) or at the start of a later class.
ST_HD void rowcursor_serve(const PlanView& P, RowCursor& rc, int64_t c, int64_t batch_end, RowLatch& L) {
  // (every field initialised: with the latch left undefined for the lanes no row serves, nvcc 12.9 merged L.valsp with the
  // cursor's register and every lane saw the values of the LAST row of the call -- found with st_debug_rowwalk_device)
  L.state = 0;
  L.valsp = 0; L.b = 0; L.m = 0; L.o = 0; L.ci = 0;
  while (rc.cur < batch_end) {
    if (rc.cur >= rc.cls_end) {  // alignment padding, then the next class
      if (c >= rc.cur && c < rc.next_off) L.state = 2;
      rowcursor_enter_class(P, rc, rc.ci + 1);
      if (rc.ci < P.ncls) rowcursor_refresh(P, rc);
      continue;
    }
    const int32_t len = rc.len - rc.off;  // what is left of the row
    const int64_t left = batch_end - rc.cur;
    const int32_t take = left < (int64_t)len ? (int32_t)left : len;
    if (c >= rc.cur && c < rc.cur + take) {
      L.valsp = rc.valsp; L.b = rc.b; L.m = rc.m; L.o = rc.off + (int32_t)(c - rc.cur); L.ci = rc.ci; L.state = 1;
    }
    rc.cur += take;
    if (take < len) {
      rc.off += take;  // the row straddles the batch
    } else {
      rc.off = 0;
      if (rc.cur < rc.cls_end) {
        const ClassDesc& C = P.cls[rc.ci];
        const int g = C.run_len[C.nruns - 1];
        permcls_advance(P, C, rc.vals, g < kRowTauMax ? g : kRowTauMax);
        rowcursor_refresh(P, rc);
      }
    }
  }
}

// lane side: the sorted multi-index (8 slots, entries beyond the rank are 0xff) of the o-th component of a row.
// info = row_class_info of the row's class; B23[(k - 2) * kRowBinomStride + n] = C(n, k) for k = 2, 3.
ST_HD void row_component(unsigned long long valsp, int32_t b, int32_t m, int32_t o, unsigned long long info, const int32_t* B23, int32_t* K) {
  const int tau = (int)((info >> 32) & 15), nvals = (int)((info >> 36) & 15), s = (int)((info >> 40) & 15);
  // the earlier runs' values: the tail steps over those above b.  One or two of them (most components) are kept in two
  // registers, ascending; more go through a count to a fixed point
  int32_t e0 = 0x7fffffff, e1 = 0x7fffffff;
  if (s == 1) {
    e0 = (int32_t)(valsp & 0xffull);
  } else if (s == 2) {
    const int32_t w0 = (int32_t)(valsp & 0xffull), w1 = (int32_t)((valsp >> 8) & 0xffull);
    e0 = w0 < w1 ? w0 : w1;
    e1 = w0 < w1 ? w1 : w0;
  }
  if (e0 <= b) e0 = 0x7fffffff;  // (then e1, if any, is the only one above b -- or none)
  if (e1 <= b) e1 = 0x7fffffff;
  int32_t used[8];
  if (s > 2) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int32_t w = (int32_t)((valsp >> (8 * u)) & 0xffull);
      used[u] = (u < s && w > b) ? w : 0x7fffffff;
    }
  }
  int32_t prev = -1, r = o;
  uint32_t tailw = 0;
#pragma unroll
  for (int i = 0; i < kRowTauMax; ++i) {
    if (i >= tau) break;
    const int k = tau - i;
    int32_t y;
    if (k == 1) {
      y = prev + 1 + r;
    } else if (k == 2) {
      // pairs (y, .) of the N values above prev: z = y - prev - 1 is the largest z with z (2N - z - 1) / 2 <= r
      const int32_t N = m - 1 - prev, t2 = 2 * N - 1;
#ifdef __CUDA_ARCH__
      int32_t z = (int32_t)(((float)t2 - sqrtf((float)(t2 * t2 - 8 * r))) * 0.5f);
#else
      int32_t z = (int32_t)(((double)t2 - sqrt((double)(t2 * t2 - 8 * r))) * 0.5);
#endif
      z = z < 0 ? 0 : (z > N - 2 ? N - 2 : z);
      while (z > 0 && z * (t2 - z) / 2 > r) --z;
      while (z < N - 2 && (z + 1) * (t2 - z - 1) / 2 <= r) ++z;
      r -= z * (t2 - z) / 2;
      y = prev + 1 + z;
    } else {
      const int32_t* Bk = B23 + (k - 2) * kRowBinomStride;
      const int32_t all = Bk[m - 1 - prev];
      int32_t lo = prev + 1, hi = m - k;
      while (lo < hi) {
        const int32_t mid = (lo + hi + 1) >> 1;
        if (all - Bk[m - mid] <= r) lo = mid; else hi = mid - 1;
      }
      r -= all - Bk[m - lo];
      y = lo;
    }
    prev = y;
    // the y-th free value above b
    int32_t v = b + 1 + y;
    if (s > 2) {
      for (;;) {
        int32_t cnt = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) cnt += (used[u] <= v);
        if (b + 1 + y + cnt == v) break;
        v = b + 1 + y + cnt;
      }
    } else {
      v += (e0 <= v);
      v += (e1 <= v);
    }
    tailw |= (uint32_t)v << (8 * i);
  }
  {  // the tail's bytes replace the bytes [nvals - tau, nvals) of the packed values (one 64-bit shift instead of one per value)
    const int sh = 8 * (nvals - tau);
    const unsigned long long tm = (unsigned long long)((1u << (8 * tau)) - 1u) << sh;
    valsp = (valsp & ~tm) | ((unsigned long long)tailw << sh);
  }
  // class order -> one value per index entry (the entry -> value-position nibbles of `info` are a byte-permute selector) -> sorted
  // (odd-even merge sort network for 8 keys)
#ifdef __CUDA_ARCH__
  {
    const uint32_t vlo = (uint32_t)valsp, vhi = (uint32_t)(valsp >> 32);
    const uint32_t w0 = __byte_perm(vlo, vhi, (uint32_t)info & 0xffffu), w1 = __byte_perm(vlo, vhi, ((uint32_t)info >> 16) & 0xffffu);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      K[e] = (int32_t)((w0 >> (8 * e)) & 0xffu);
      K[4 + e] = (int32_t)((w1 >> (8 * e)) & 0xffu);
    }
  }
#else
#pragma unroll
  for (int e = 0; e < 8; ++e) K[e] = (int32_t)((valsp >> (8 * ((info >> (4 * e)) & 7))) & 0xffull);
#endif
#define ST_CE(i, j) { const int32_t lo_ = K[i] < K[j] ? K[i] : K[j], hi_ = K[i] < K[j] ? K[j] : K[i]; K[i] = lo_; K[j] = hi_; }
  ST_CE(0, 1) ST_CE(2, 3) ST_CE(4, 5) ST_CE(6, 7)
  ST_CE(0, 2) ST_CE(1, 3) ST_CE(4, 6) ST_CE(5, 7)
  ST_CE(1, 2) ST_CE(5, 6)
  ST_CE(0, 4) ST_CE(1, 5) ST_CE(2, 6) ST_CE(3, 7)
  ST_CE(2, 4) ST_CE(3, 5)
  ST_CE(1, 2) ST_CE(3, 4) ST_CE(5, 6)
#undef ST_CE
}

// packed coordinate of the permcls layout -> sorted multi-index (false for alignment padding)
ST_HD bool permcls_coord_sorted(const PlanView& P, int64_t c, int32_t* K) {
  const int ci = class_of_coord(P, c);
  const ClassDesc& C = P.cls[ci];
  const int64_t pos = c - C.offset;
  if (pos >= C.size) return false;
  int32_t vals[ST_MAX_RANK];
  permcls_unrank_vals(P, C, pos, vals);
  int n = 0;
  for (int v = 0; v < C.nvals; ++v)
    for (int m = 0; m < C.mult[v]; ++m) {
      const int32_t x = vals[v];
      int u = n++;
      while (u > 0 && K[u - 1] > x) { K[u] = K[u - 1]; --u; }
      K[u] = x;
    }
  return true;
}

}  // namespace st

// ---- host-side plan cache (st_plan.cu) ---------------------------------------------------------------
namespace st {

struct HostPlan {
  int rank = 0;
  int64_t dim = 0;
  int ncls = 0;
  ClassDesc* h_cls = nullptr;   // [ncls]
  int64_t* h_offsets = nullptr; // [ncls + 1]
  int64_t* h_binom = nullptr;
  int64_t binom_rows = 0;
  int64_t flat_size = 0;
  bool flat_overflow = false;
  bool size_overflow = false;
  PlanView host_view() const;
};

// Returns the cached host plan (never freed), or nullptr with the error message set.
const HostPlan* get_host_plan(int rank, int64_t dim);
// Returns a view whose pointers live on the current CUDA device; status via return code.
int get_device_plan(int rank, int64_t dim, PlanView* out);

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_cuda(cudaError_t e, const char* what);
// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (device, kernel) -- the attribute is per device (st_vec.cu)
int set_max_dynamic_smem(const void* func, int bytes);
int sm_count();  // of the current device (st_vec.cu)

}  // namespace st
