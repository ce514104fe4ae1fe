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

// class containing the packed coordinate c (offsets ascending); padding belongs to the class before it
ST_HD int class_of_coord(const PlanView& P, int64_t c) {
  int lo = 0, hi = P.ncls - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.offsets[mid] <= c) lo = mid; else hi = mid - 1;
  }
  return lo;
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

}  // namespace st
