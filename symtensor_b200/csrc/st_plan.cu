// Class tables and the per-(rank, dim) plan cache; host-side C-ABI entry points.
//
// Replaces (exact integers throughout):
//   utils._perm_classes / _all_index_counts          symtensor/utils.py:839-856, 1000-1002
//   utils._get_permclass_size                         symtensor/utils.py:925-933
//   utils.get_permclass_multiplicity / multinom       symtensor/utils.py:207-223, 760-776
//   SymmetricTensor.indep_size                        symtensor/base.py:833-844
//   PosRegistry / _convert_dense_index (single index) symtensor/permcls_symtensor.py:422-479
//   flat index_of_multicombination                    symtensor/flat_symtensor.py:39-50
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "st_common.cuh"

namespace st {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return ST_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return ST_ERR_CUDA;
}

static const int64_t kI64Max = INT64_MAX;

static int64_t sat_mul(int64_t a, int64_t b, bool* ovf) {
  const __int128 p = (__int128)a * (__int128)b;
  if (p > (__int128)kI64Max) { *ovf = true; return kI64Max; }
  return (int64_t)p;
}

static void gen_partitions(int n, int maxpart, std::vector<int>& cur, std::vector<std::vector<int>>& out) {
  if (n == 0) { out.push_back(cur); return; }
  for (int first = (n < maxpart ? n : maxpart); first >= 1; --first) {
    cur.push_back(first);
    gen_partitions(n - first, first, cur, out);
    cur.pop_back();
  }
}

PlanView HostPlan::host_view() const {
  PlanView v;
  v.rank = rank;
  v.ncls = ncls;
  v.dim = dim;
  v.total = h_offsets[ncls];
  v.flat_size = flat_size;
  v.cls = h_cls;
  v.offsets = h_offsets;
  v.binom = h_binom;
  return v;
}

static HostPlan* build_host_plan(int rank, int64_t dim) {
  HostPlan* hp = new HostPlan();
  hp->rank = rank;
  hp->dim = dim;
  // binomial table C(n, k), n in [0, dim + rank], k in [0, rank], saturated
  const int K1 = rank + 1;
  hp->binom_rows = dim + rank + 1;
  hp->h_binom = new int64_t[hp->binom_rows * K1];
  for (int64_t n = 0; n < hp->binom_rows; ++n) {
    for (int k = 0; k < K1; ++k) {
      int64_t v;
      if (k == 0) v = 1;
      else if (n == 0) v = 0;
      else {
        const int64_t a = hp->h_binom[(n - 1) * K1 + k - 1], b = hp->h_binom[(n - 1) * K1 + k];
        v = (a > kI64Max - b) ? kI64Max : a + b;
      }
      hp->h_binom[n * K1 + k] = v;
    }
  }
  if (rank == 0) hp->flat_size = 1;  // also for dim == 0 (symtensor/tests/test_utils.py:82)
  else hp->flat_size = hp->h_binom[(dim + rank - 1) * K1 + rank];
  hp->flat_overflow = (hp->flat_size == kI64Max);

  std::vector<std::vector<int>> parts;
  std::vector<int> cur;
  gen_partitions(rank, rank, cur, parts);
  hp->ncls = (int)parts.size();
  hp->h_cls = new ClassDesc[hp->ncls];
  hp->h_offsets = new int64_t[hp->ncls + 1];
  memset(hp->h_cls, 0, sizeof(ClassDesc) * hp->ncls);
  int64_t off = 0;
  for (int c = 0; c < hp->ncls; ++c) {
    ClassDesc& C = hp->h_cls[c];
    const std::vector<int>& p = parts[c];
    C.nvals = (int)p.size();
    for (int i = 0; i < C.nvals; ++i) C.mult[i] = p[i];
    int nr = 0;
    for (int i = 0; i < C.nvals; ++i) {
      if (i == 0 || p[i] != p[i - 1]) { C.run_start[nr] = i; C.run_len[nr] = 1; C.run_mult[nr] = p[i]; ++nr; }
      else ++C.run_len[nr - 1];
    }
    C.nruns = nr;
    bool ovf = false;
    int64_t size = 1, remaining = dim;
    for (int j = 0; j < nr; ++j) {
      const int g = C.run_len[j];
      const int64_t rdx = remaining < g ? 0 : hp->h_binom[remaining * K1 + g];
      if (rdx == kI64Max) ovf = true;
      C.radix[j] = rdx;
      size = sat_mul(size, rdx, &ovf);
      remaining -= g;
    }
    if (C.nvals > dim) size = 0;
    C.size = size;
    if (ovf && size != 0) hp->size_overflow = true;
    // gamma = rank! / prod(m_k!) built as a product of binomials C(rem, m_k)
    int64_t gamma = 1;
    int rem = rank;
    for (int i = 0; i < C.nvals; ++i) {
      // C(rem, m) with rem <= ST_MAX_RANK: small exact values
      int64_t b = 1;
      for (int t = 1; t <= p[i]; ++t) b = b * (rem - p[i] + t) / t;
      gamma *= b;
      rem -= p[i];
    }
    C.gamma = gamma;
    C.offset = off;
    hp->h_offsets[c] = off;
    const int64_t padded = size > kI64Max - ST_CLASS_ALIGN ? kI64Max
                                                            : (size + ST_CLASS_ALIGN - 1) / ST_CLASS_ALIGN * ST_CLASS_ALIGN;
    if (off > kI64Max - padded) { hp->size_overflow = true; off = kI64Max; }
    else off += padded;
  }
  hp->h_offsets[hp->ncls] = off;
  return hp;
}

static std::mutex g_mu;
static std::map<std::pair<int, int64_t>, HostPlan*> g_host_plans;

struct DevPlan {
  PlanView view;
};
static std::map<std::tuple<int, int, int64_t>, DevPlan*> g_dev_plans;

const HostPlan* get_host_plan(int rank, int64_t dim) {
  if (rank < 0 || rank > ST_MAX_RANK) { set_error("rank %d outside [0, %d]", rank, ST_MAX_RANK); return nullptr; }
  // the binomial table has (dim + rank + 1) x (rank + 1) int64 entries: keep it below 256 MiB
  if (dim < 0 || (dim + rank + 1) > ((int64_t)1 << 25) / (rank + 1)) {
    set_error("dim %lld unsupported for rank %d (binomial table too large)", (long long)dim, rank);
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_pair(rank, dim);
  auto it = g_host_plans.find(key);
  if (it != g_host_plans.end()) return it->second;
  HostPlan* hp = build_host_plan(rank, dim);
  g_host_plans[key] = hp;
  return hp;
}

int get_device_plan(int rank, int64_t dim, PlanView* out) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return ST_ERR_INVALID;
  if (hp->size_overflow) { set_error("tensor of rank %d dim %lld does not fit int64 positions", rank, (long long)dim); return ST_ERR_OVERFLOW; }
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(dev, rank, dim);
  auto it = g_dev_plans.find(key);
  if (it != g_dev_plans.end()) { *out = it->second->view; return ST_OK; }
  // one allocation: [ClassDesc x ncls][offsets x (ncls+1)][binom]
  const size_t b_cls = sizeof(ClassDesc) * hp->ncls;
  const size_t b_off = sizeof(int64_t) * (hp->ncls + 1);
  const size_t b_bin = sizeof(int64_t) * hp->binom_rows * (rank + 1);
  char* d = nullptr;
  rc = check_cuda(cudaMalloc(&d, b_cls + b_off + b_bin), "cudaMalloc(plan)");
  if (rc) return rc;
  // synchronous copies: plans are created once per (device, rank, dim)
  rc = check_cuda(cudaMemcpy(d, hp->h_cls, b_cls, cudaMemcpyHostToDevice), "cudaMemcpy(plan classes)");
  if (!rc) rc = check_cuda(cudaMemcpy(d + b_cls, hp->h_offsets, b_off, cudaMemcpyHostToDevice), "cudaMemcpy(plan offsets)");
  if (!rc) rc = check_cuda(cudaMemcpy(d + b_cls + b_off, hp->h_binom, b_bin, cudaMemcpyHostToDevice), "cudaMemcpy(plan binom)");
  if (rc) { cudaFree(d); return rc; }
  DevPlan* dp = new DevPlan();
  dp->view = hp->host_view();
  dp->view.cls = reinterpret_cast<const ClassDesc*>(d);
  dp->view.offsets = reinterpret_cast<const int64_t*>(d + b_cls);
  dp->view.binom = reinterpret_cast<const int64_t*>(d + b_cls + b_off);
  g_dev_plans[key] = dp;
  *out = dp->view;
  return ST_OK;
}

}  // namespace st

using namespace st;

extern "C" {

int st_version(void) { return 100; }

const char* st_last_error(void) { return st::g_err; }

int64_t st_launch_count(void) { return st::g_launches.load(); }

int st_num_classes(int rank) {
  const HostPlan* hp = get_host_plan(rank, 1);
  return hp ? hp->ncls : -1;
}

int st_class_table(int rank, int64_t dim, int32_t* parts, int32_t* nparts, int64_t* sizes, int64_t* mults,
                   int64_t* offsets) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return ST_ERR_INVALID;
  for (int c = 0; c < hp->ncls; ++c) {
    const ClassDesc& C = hp->h_cls[c];
    if (parts) for (int i = 0; i < rank; ++i) parts[c * rank + i] = i < C.nvals ? C.mult[i] : 0;
    if (nparts) nparts[c] = C.nvals;
    if (sizes) sizes[c] = C.size;
    if (mults) mults[c] = C.gamma;
    if (offsets) offsets[c] = hp->h_offsets[c];
  }
  if (offsets) offsets[hp->ncls] = hp->h_offsets[hp->ncls];
  if (hp->size_overflow) { set_error("class sizes of rank %d dim %lld exceed int64", rank, (long long)dim); return ST_ERR_OVERFLOW; }
  return ST_OK;
}

int st_indep_size(int rank, int64_t dim, int64_t* out) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp || !out) { if (hp) set_error("null output"); return ST_ERR_INVALID; }
  *out = hp->flat_size;
  if (hp->flat_overflow) { set_error("C(dim+rank-1, rank) exceeds int64"); return ST_ERR_OVERFLOW; }
  return ST_OK;
}

int st_host_permcls_rank(int rank, int64_t dim, const int32_t* idx, int32_t* cls, int64_t* pos) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp || !idx || !cls || !pos) { if (hp) set_error("null pointer"); return ST_ERR_INVALID; }
  if (hp->size_overflow) { set_error("positions exceed int64"); return ST_ERR_OVERFLOW; }
  const PlanView P = hp->host_view();
  int32_t vals[ST_MAX_RANK];
  const int c = classify_index(P, idx, vals);
  if (c < 0) { set_error("index entry outside [0, dim)"); return ST_ERR_INVALID; }
  *cls = c;
  *pos = permcls_rank_vals(P, P.cls[c], vals);
  return ST_OK;
}

// debug / test hook (host only): `count` consecutive components of class `cls` from position `pos`, the first by a full
// unrank, the others by odometer steps (permcls_next_vals: what the outer kernel walks its runs with); returns how many were
// written (fewer at the end of the class)
int64_t st_debug_permcls_successors(int rank, int64_t dim, int32_t cls, int64_t pos, int64_t count, int32_t* idx) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp || !idx || cls < 0 || cls >= hp->ncls) return -1;
  const PlanView P = hp->host_view();
  const ClassDesc& C = P.cls[cls];
  if (pos < 0 || pos >= C.size) return 0;
  int32_t vals[ST_MAX_RANK];
  permcls_unrank_vals(P, C, pos, vals);
  int64_t n = 0;
  for (;;) {
    int o = 0;
    for (int i = 0; i < C.nvals; ++i) for (int m = 0; m < C.mult[i]; ++m) idx[n * rank + o++] = vals[i];
    ++n;
    if (n >= count || !permcls_next_vals(P, C, vals)) break;
  }
  return n;
}

// debug / test hook (host only): the row walk the multiply.outer kernel serves its coordinates with, replayed on the CPU --
// spans of `span` coordinates (a seek each), batches of 32 lanes, every lane with its own copy of the batch's cursor.  Writes
// the sorted multi-index of every coordinate of [begin, end) (rank ints each; -1 for padding) and returns end - begin.
int64_t st_debug_rowwalk(int rank, int64_t dim, int64_t begin, int64_t end, int64_t span, int32_t* idx) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp || !idx || dim > 255 || rank > 8 || span < 1 || begin < 0 || end < begin) return -1;
  const PlanView P = hp->host_view();
  if (end > P.total) return -1;
  std::vector<int32_t> B23(2 * kRowBinomStride);
  for (int k = 2; k <= 3; ++k)
    for (int n = 0; n < kRowBinomStride; ++n) B23[(k - 2) * kRowBinomStride + n] = k == 2 ? n * (n - 1) / 2 : n * (n - 1) * (n - 2) / 6;
  for (int64_t s0 = begin; s0 < end; s0 += span) {
    const int64_t s1 = std::min(end, s0 + span);
    RowCursor rc;
    rowcursor_seek(P, rc, s0);
    for (int64_t b = s0; b < s1; b += 32) {
      const int64_t be = std::min(s1, b + 32);
      RowCursor after = rc;
      for (int lane = 0; lane < 32; ++lane) {
        const int64_t c = b + lane;
        RowCursor mine = rc;
        RowLatch L;
        rowcursor_serve(P, mine, c, be, L);
        if (lane == 0) after = mine;
        if (c >= be) continue;
        int32_t* o = idx + (c - begin) * rank;
        if (L.state != 1) { for (int i = 0; i < rank; ++i) o[i] = -1; continue; }
        int32_t K[8];
        row_component(L.valsp, L.b, L.m, L.o, row_class_info(P.cls[L.ci], rank), B23.data(), K);
        for (int i = 0; i < rank; ++i) o[i] = K[i];
      }
      rc = after;
    }
  }
  return end - begin;
}

int st_host_permcls_unrank(int rank, int64_t dim, int32_t cls, int64_t pos, int32_t* idx) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp || !idx) { if (hp) set_error("null pointer"); return ST_ERR_INVALID; }
  if (hp->size_overflow) { set_error("positions exceed int64"); return ST_ERR_OVERFLOW; }
  if (cls < 0 || cls >= hp->ncls) { set_error("class ordinal %d outside [0, %d)", cls, hp->ncls); return ST_ERR_INVALID; }
  const PlanView P = hp->host_view();
  const ClassDesc& C = P.cls[cls];
  if (pos < 0 || pos >= C.size) { set_error("position %lld outside [0, %lld)", (long long)pos, (long long)C.size); return ST_ERR_INVALID; }
  int32_t vals[ST_MAX_RANK];
  permcls_unrank_vals(P, C, pos, vals);
  int o = 0;
  for (int i = 0; i < C.nvals; ++i) for (int m = 0; m < C.mult[i]; ++m) idx[o++] = vals[i];
  return ST_OK;
}

int st_host_flat_rank(int rank, int64_t dim, const int32_t* idx, int64_t* pos) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp || !idx || !pos) { if (hp) set_error("null pointer"); return ST_ERR_INVALID; }
  if (hp->flat_overflow) { set_error("positions exceed int64"); return ST_ERR_OVERFLOW; }
  int32_t s[ST_MAX_RANK];
  for (int i = 0; i < rank; ++i) {
    const int32_t v = idx[i];
    if (v < 0 || v >= dim) { set_error("index entry outside [0, dim)"); return ST_ERR_INVALID; }
    int u = i;
    while (u > 0 && s[u - 1] > v) { s[u] = s[u - 1]; --u; }
    s[u] = v;
  }
  *pos = flat_rank_sorted(hp->host_view(), s);
  return ST_OK;
}

int st_host_flat_unrank(int rank, int64_t dim, int64_t pos, int32_t* idx) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp || !idx) { if (hp) set_error("null pointer"); return ST_ERR_INVALID; }
  if (hp->flat_overflow) { set_error("positions exceed int64"); return ST_ERR_OVERFLOW; }
  if (pos < 0 || pos >= hp->flat_size) { set_error("position outside [0, size)"); return ST_ERR_INVALID; }
  flat_unrank_sorted(hp->host_view(), pos, idx);
  return ST_OK;
}

}  // extern "C"
