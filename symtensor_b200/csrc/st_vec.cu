// contract_all_indices_with_vector on the packed layouts (symtensor/symalg.py:505-527).
//
// The reference computes  s = sum_{i1..ir} A[i1..ir] x[i1]...x[ir]  as r rounds of
// [todense -> np.tensordot(., x, 1) -> r!-symmetrize -> repack].  In packed space the same number is
//     s = sum_classes gamma_c * sum_p A_c[p] * prod_j x[v_j(p)]^{m_j}                       
// i.e. ONE streaming pass over the packed buffer: every stored value is read from HBM exactly once, so the
// kernel is HBM-bound (algorithmic bytes = sizeof(T) per packed component).
//
// Two kernels:
//  * vec_tail_kernel ("tail table", the production path for ST_LAYOUT_PERMCLS).  Storage order is
//    lexicographic in the distinct values, so for a fixed "head" (all values but the last tau values of
//    the last run) the tail components are CONTIGUOUS in memory and their weights are a contiguous slice of
//    a head-independent table  T[q] = prod_{u in q-th tau-combination} xrel[u]^mu  kept in shared memory.
//    A warp walks the heads of its range with a warp-uniform odometer and streams each block as a
//    coalesced dot product  hw * <A[block], T[slice]>  (per element: one LDG, one LDS, one FMA).
//  * vec_generic_kernel (one full unrank per element; any layout; also the flat-layout path and the
//    cross-check used by the tests).
// Both write one fp64 partial per CTA; vec_finalize_kernel adds them in a fixed order (deterministic).
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "st_common.cuh"
#include "st_vec_core.cuh"

namespace st {

static const int kMaxCtas = 148 * 8;
static const int kMaxPartials = 4096;   // fp64 partial sums per launch (one per work item / per CTA)
static const int kTailThreads = 512;    // 16 warps per CTA, one CTA per SM (the tail table fills shared memory)
static const int kStageBytes = 1024;    // two 16-byte cp.async per lane
// tuning knobs (st_set_tuning): work items per CTA, ring depth
static int g_items_per_cta = 8;
static int g_ring_stages = 4;

template <typename T>
struct VecArgs {
  PlanView P;
  const TailStrategy* strat;
  const T* A;        // points at packed coordinate `begin`
  const T* x;
  int64_t begin, end;
  double* partials;  // [gridDim.x]
  int64_t n_items;
  int64_t item_elems;  // packed coordinates per work item (multiple of ST_CLASS_ALIGN)
  unsigned long long* counter;  // [2]: next work item, finished CTAs (library-owned, per stream, self-resetting)
  int32_t tbl_cap;     // table entries that fit the dynamic shared memory
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum -> *dst; `red` has one slot per warp
__device__ __forceinline__ void block_store_partial(double v, double* red, double* dst) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double s = lane < nw ? red[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) *dst = s;
  }
}

__global__ void vec_finalize_kernel(const double* __restrict__ partials, int n, double* out64, float* out32) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = warp_sum(s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) {
      if (out64) *out64 = t;
      if (out32) *out32 = (float)t;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// generic kernel: one thread per packed coordinate, full unrank
// ------------------------------------------------------------------------------------------------------
template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256) vec_generic_kernel(VecArgs<T> a) {
  __shared__ double red[32];
  const PlanView& P = a.P;
  double acc = 0.0;
  for (int64_t c = a.begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < a.end; c += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)ld_stream(a.A + (c - a.begin));
    if (LAYOUT == ST_LAYOUT_PERMCLS) {
      const int ci = class_of_coord(P, c);
      const ClassDesc& C = P.cls[ci];
      const int64_t pos = c - C.offset;
      if (pos >= C.size) continue;  // alignment padding
      int32_t vals[ST_MAX_RANK];
      permcls_unrank_vals(P, C, pos, vals);
      double w = (double)C.gamma;
      for (int k = 0; k < C.nvals; ++k) {
        const double xv = (double)a.x[vals[k]];
        for (int m = 0; m < C.mult[k]; ++m) w *= xv;
      }
      acc += v * w;
    } else {
      int32_t s[ST_MAX_RANK];
      flat_unrank_sorted(P, c, s);
      // weight = r!/prod(n_v!) * prod x[i_k]: the j-th repeat of a value contributes x/j, position k a factor k+1
      double w = 1.0;
      int rep = 0;
      for (int k = 0; k < P.rank; ++k) {
        rep = (k > 0 && s[k] == s[k - 1]) ? rep + 1 : 1;
        w *= (double)a.x[s[k]] * (double)(k + 1) / (double)rep;
      }
      acc += v * w;
    }
  }
  block_store_partial(acc, red, a.partials + blockIdx.x);
}

// One warp streams segment positions [q0, q1): unrank the start with the whole warp, then the staged walk.
template <typename T, int NST>
__device__ __forceinline__ double walk_range_staged(const PlanView& P, const TailStrategy& S, const T* tbl, const T* xr,
                                                    const int32_t* blen, double wE, const T* __restrict__ Aseg, int64_t q0, int64_t q1,
                                                    int lane, T* ring) {
  int32_t u0[ST_MAX_RANK];
  comb_unrank_warp(P.binom, P.rank, q0, S.Rt, S.gt, u0, lane);
  return walk_range<T, NST, true>(P, S, tbl, xr, blen, wE, Aseg, q0, q1, lane, ring, u0);
}

// Shared memory: [T / private xr tables][xr: dim][xs: dim][blen: dim x i32][ring: nwarps x NST x 512 B][ctrl]
template <typename T, int NST>
__global__ void __launch_bounds__(kTailThreads, 1) vec_tail_kernel(VecArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PlanView& P = a.P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarps = kTailThreads / 32;
  T* tbl = reinterpret_cast<T*>(smem_raw);
  T* priv = tbl + (size_t)warp * P.dim;  // this warp's private xr table (tau == 1 classes); aliases the shared table
  size_t off = ((size_t)a.tbl_cap * sizeof(T) + 15) / 16 * 16;
  T* xr = reinterpret_cast<T*>(smem_raw + off);
  T* xs = xr + P.dim;
  off = (off + 2 * (size_t)P.dim * sizeof(T) + 15) / 16 * 16;
  int32_t* blen = reinterpret_cast<int32_t*>(smem_raw + off);
  off = (off + (size_t)P.dim * sizeof(int32_t) + 15) / 16 * 16;
  T* ring = reinterpret_cast<T*>(smem_raw + off + (size_t)warp * NST * kStageBytes);
  off += (size_t)nwarps * NST * kStageBytes;
  TailCtrl* ctl = reinterpret_cast<TailCtrl*>(smem_raw + off);
  if (threadIdx.x == 0) { ctl->cur_cls = -1; ctl->cur_seg = -1; }
  for (int i = threadIdx.x; i < P.dim; i += kTailThreads) xs[i] = a.x[i];
  int priv_cls = -1;       // (class, segment) the private table was built for -- warp-uniform
  int64_t priv_seg = -1;
  double priv_wE = 0.0;

  while (true) {
    // ---- dynamic scheduling: CTAs claim work items from a global counter (slow regions cannot pile up on one CTA)
    __syncthreads();
    if (threadIdx.x == 0) ctl->item = (long long)atomicAdd(a.counter, 1ULL);
    __syncthreads();
    const int64_t item = ctl->item;
    if (item >= a.n_items) break;
    double total = 0.0;
    const int64_t c0 = a.begin + item * a.item_elems;
    const int64_t c1 = (c0 + a.item_elems < a.end) ? c0 + a.item_elems : a.end;
    int64_t coord = c0;
    while (coord < c1) {
      const int ci = class_of_coord(P, coord);
      const ClassDesc& C = P.cls[ci];
      int64_t pos = coord - C.offset;
      if (pos >= C.size) { coord = P.offsets[ci + 1]; continue; }  // padding up to the next class
      const int64_t pend = (C.size < c1 - C.offset) ? C.size : c1 - C.offset;
      const TailStrategy S = a.strat[ci];
      const T* Acls = a.A + (C.offset - a.begin);
      if (S.tau == 1) {
        // ---- private tables: warps split the class range; each warp walks its segments on its own
        if (ctl->cur_cls != -1) {  // CTA-uniform: the private tables overwrite the shared table
          __syncthreads();
          if (threadIdx.x == 0) { ctl->cur_cls = -1; ctl->cur_seg = -1; }
          __syncthreads();
        }
        const int64_t len = pend - pos;
        int64_t per = (len + nwarps - 1) / nwarps;
        per = (per + 31) / 32 * 32;
        int64_t w0 = pos + (int64_t)warp * per;
        const int64_t w1 = (w0 + per < pend) ? w0 + per : pend;
        while (w0 < w1) {
          const int64_t sidx = w0 / S.seg;
          const int64_t sbase = sidx * S.seg;
          const int64_t q1 = (S.seg < w1 - sbase) ? S.seg : w1 - sbase;
          if (priv_cls != ci || priv_seg != sidx) {
            int32_t E[ST_MAX_RANK];
            priv_wE = unrank_earlier<T>(P, C, sidx, xs, E);
            __syncwarp();
            for (int uu = lane; uu < S.Rt; uu += 32) priv[uu] = xrel_pow<T>(xs, E, S.nE, S.mu, uu);
            __syncwarp();
            priv_cls = ci;
            priv_seg = sidx;
          }
          total += walk_range_staged<T, NST>(P, S, priv, priv, nullptr, priv_wE, Acls + sbase, w0 - sbase, q1, lane, ring);
          w0 = sbase + q1;
        }
        pos = pend;
      } else {
        // ---- shared table: the CTA builds T once per segment
        while (pos < pend) {
          const int64_t sidx = pos / S.seg;
          const int64_t sbase = sidx * S.seg;
          const int64_t q0 = pos - sbase;
          const int64_t q1 = (S.seg < pend - sbase) ? S.seg : pend - sbase;
          if (ctl->cur_cls != ci || ctl->cur_seg != sidx) {  // CTA-uniform
            __syncthreads();  // everybody is done with the previous tables
            if (threadIdx.x == 0) {
              ctl->wE = unrank_earlier<T>(P, C, sidx, xs, ctl->E);
              ctl->cur_cls = ci;
              ctl->cur_seg = sidx;
            }
            priv_cls = -1;  // the shared table overwrites the private ones
            __syncthreads();
            for (int uu = threadIdx.x; uu < S.Rt; uu += kTailThreads) {
              xr[uu] = xrel_pow<T>(xs, ctl->E, S.nE, S.mu, uu);
              blen[uu] = (int32_t)binom_at(P.binom, P.rank, S.Rt - 1 - uu, S.tau);
            }
            __syncthreads();
            {  // T[q] over the tau-combinations of range(Rt); each thread fills a contiguous slice
              const int64_t per = (S.tbl_n + kTailThreads - 1) / kTailThreads;
              const int64_t q = (int64_t)threadIdx.x * per;
              build_table_slice<T>(P, S, xr, tbl, q, (q + per < S.tbl_n) ? q + per : S.tbl_n);
            }
            __syncthreads();
          }
          // warps split the piece [q0, q1) evenly (multiples of 32 components)
          const int64_t len = q1 - q0;
          int64_t per = (len + nwarps - 1) / nwarps;
          per = (per + 31) / 32 * 32;
          const int64_t w0 = q0 + (int64_t)warp * per;
          const int64_t w1 = (w0 + per < q1) ? w0 + per : q1;
          if (w0 < w1) total += walk_range_staged<T, NST>(P, S, tbl, xr, blen, ctl->wE, Acls + sbase, w0, w1, lane, ring);
          pos = sbase + q1;
        }
      }
      coord = C.offset + pos;
    }
    // one partial per ITEM (not per CTA): the final sum does not depend on which CTA processed which item
    __syncthreads();
    block_store_partial(total, ctl->red, a.partials + item);
  }
  // self-cleaning counters: the last CTA to finish resets them for the next launch on this stream
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(a.counter + 1, 1ULL);
    if (ticket == gridDim.x - 1) {
      a.counter[0] = 0ULL;
      a.counter[1] = 0ULL;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static int g_variant = 0;
int g_force_tau = 0;  // test hook: force the tail length (0 = cost model)

struct StratKey {
  int dev, rank, esize, nst;
  int64_t dim;
  bool operator<(const StratKey& o) const {
    return std::tie(dev, rank, esize, nst, dim) < std::tie(o.dev, o.rank, o.esize, o.nst, o.dim);
  }
};
struct StratEntry {
  TailStrategy* d_strat;
  int32_t tbl_cap;
  size_t smem_bytes;
  bool supported;
};
static std::mutex g_smu;
static std::map<StratKey, StratEntry> g_strats;

static double dbinom(const HostPlan* hp, int64_t n, int k) {
  if (n < 0 || k < 0 || k > hp->rank) return 0.0;
  return (double)hp->h_binom[n * (hp->rank + 1) + k];
}

// Choose tau per class with a small cost model (estimated warp-instructions per component):
//   streaming      0.2 (mode A: LDG + LDS + FMA)  /  0.3 + 0.03 nE (mode B: relabel + pow on the fly)
//   per block      ~30-40 warp-uniform instructions for the head odometer
//   table builds   ~0.5 per entry, amortised over a segment (multi-run classes) or over a CTA's share of the
//                  class (single-run classes, table built once per CTA)
// Pure host function (also used by the CPU emulation harness in tests/emu).
bool compute_tail_strategy(const HostPlan* hp, int esize, int nwarps, int nst, std::vector<TailStrategy>& st, int32_t* tbl_cap, size_t* smem_bytes) {
  const int rank = hp->rank;
  const int64_t dim = hp->dim;
  const size_t smem_budget = 224 * 1024;
  // everything but the table: xr, xs, blen, the cp.async rings, control block, alignment slack
  const size_t fixed = (size_t)2 * dim * esize + (size_t)dim * 4 + (size_t)nwarps * nst * kStageBytes + sizeof(TailCtrl) + 128;
  if (fixed + 32 * esize > smem_budget) return false;  // x itself does not fit shared memory
  const int64_t cap = (int64_t)((smem_budget - fixed) / esize);
  if (cap < (int64_t)nwarps * dim) return false;  // the per-warp private tables (which alias the table) must fit
  st.assign(hp->ncls, TailStrategy());
  int64_t tbl_max = std::max<int64_t>(32, (int64_t)nwarps * dim);
  for (int c = 0; c < hp->ncls; ++c) {
    const ClassDesc& C = hp->h_cls[c];
    TailStrategy& S = st[c];
    memset(&S, 0, sizeof(S));
    if (C.nvals == 0 || C.size == 0) { S.tau = 1; S.gt = 1; S.Rt = 1; S.mu = 1; S.tbl_n = 1; S.seg = 1; continue; }
    const int t = C.nruns - 1;
    S.gt = C.run_len[t];
    S.nE = C.nvals - S.gt;
    S.Rt = (int32_t)(dim - S.nE);
    S.mu = C.run_mult[t];
    S.seg = C.radix[t];
    double best = 1e300;
    int best_tau = 1;
    for (int tau = 1; tau <= S.gt; ++tau) {
      const double tn = dbinom(hp, S.Rt, tau);
      if (tau > 1 && tn > (double)cap) break;
      const double nheads = dbinom(hp, S.Rt - tau, S.gt - tau);
      const double avg_block = (double)S.seg / (nheads > 0 ? nheads : 1);
      double cost;
      if (tau == 1) {
        cost = 0.2 + 40.0 / avg_block + (S.nE ? (400.0 + 0.5 * S.Rt) / (double)S.seg : 0.0);
      } else {
        const double amort = (S.nE == 0) ? std::max(1.0, (double)C.size / 296.0) : (double)S.seg;
        cost = 0.2 + 30.0 / avg_block + (0.5 * tn + 2000.0) / amort;
      }
      if (cost < best) { best = cost; best_tau = tau; }
    }
    if (g_force_tau > 0) best_tau = std::min(g_force_tau, (int)S.gt);
    if (best_tau > 1 && dbinom(hp, S.Rt, best_tau) > (double)cap) best_tau = 1;
    S.tau = best_tau;
    S.hn = S.gt - S.tau;
    S.tbl_n = hp->h_binom[(int64_t)S.Rt * (rank + 1) + S.tau];
    if (S.tau > 1) tbl_max = std::max(tbl_max, S.tbl_n);
  }
  *tbl_cap = (int32_t)((tbl_max + 31) / 32 * 32);
  {
    size_t off = ((size_t)*tbl_cap * esize + 15) / 16 * 16;
    off = (off + 2 * (size_t)dim * esize + 15) / 16 * 16;
    off = (off + (size_t)dim * 4 + 15) / 16 * 16;
    off += (size_t)nwarps * nst * kStageBytes;
    *smem_bytes = off + sizeof(TailCtrl);
  }
  return true;
}

static int get_strategy(int rank, int64_t dim, int esize, int nst, StratEntry* out) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return ST_ERR_INVALID;
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_smu);
  StratKey key{dev, rank, esize, nst, dim};
  auto it = g_strats.find(key);
  if (it != g_strats.end()) { *out = it->second; return ST_OK; }
  StratEntry e;
  e.d_strat = nullptr;
  e.tbl_cap = 32;
  e.smem_bytes = 0;
  std::vector<TailStrategy> st;
  e.supported = compute_tail_strategy(hp, esize, kTailThreads / 32, nst, st, &e.tbl_cap, &e.smem_bytes);
  if (e.supported) {
    rc = check_cuda(cudaMalloc(&e.d_strat, sizeof(TailStrategy) * hp->ncls), "cudaMalloc(strategy)");
    if (rc) return rc;
    rc = check_cuda(cudaMemcpy(e.d_strat, st.data(), sizeof(TailStrategy) * hp->ncls, cudaMemcpyHostToDevice), "cudaMemcpy(strategy)");
    if (rc) return rc;
  }
  g_strats[key] = e;
  *out = e;
  return ST_OK;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// one {next item, finished CTAs} counter pair per (device, stream); zeroed once, reset by the kernel itself
static std::mutex g_cmu;
static std::map<std::pair<int, cudaStream_t>, unsigned long long*> g_counters;
static int get_counter(cudaStream_t stream, unsigned long long** out) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_cmu);
  auto key = std::make_pair(dev, stream);
  auto it = g_counters.find(key);
  if (it != g_counters.end()) { *out = it->second; return ST_OK; }
  unsigned long long* d = nullptr;
  rc = check_cuda(cudaMalloc(&d, 2 * sizeof(unsigned long long)), "cudaMalloc(counter)");
  if (rc) return rc;
  rc = check_cuda(cudaMemset(d, 0, 2 * sizeof(unsigned long long)), "cudaMemset(counter)");
  if (rc) return rc;
  g_counters[key] = d;
  *out = d;
  return ST_OK;
}

template <typename T, int NST>
static int launch_tail_t(VecArgs<T>& a, const StratEntry& se, int64_t len, int* grid_out, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(vec_tail_kernel<T, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                        "cudaFuncSetAttribute");
    if (rc) return rc;
    attr_set = true;
  }
  // One resident CTA per SM claims work items dynamically; the item size gives every CTA about
  // `g_items_per_cta` items: large enough to amortise the per-warp unrank, small enough to balance the tail.
  const int64_t quantum = 32 * (kTailThreads / 32);
  int64_t grid = std::min<int64_t>((int64_t)sm_count(), kMaxCtas);
  grid = std::max<int64_t>(1, std::min<int64_t>(grid, (len + quantum - 1) / quantum));
  int64_t nitems = std::min<int64_t>(grid * g_items_per_cta, kMaxPartials);
  int64_t item = (len + nitems - 1) / nitems;
  item = (item + quantum - 1) / quantum * quantum;
  a.item_elems = item;
  a.n_items = (len + item - 1) / item;
  grid = std::min<int64_t>(grid, a.n_items);
  {
    int rc = get_counter(stream, &a.counter);
    if (rc) return rc;
  }
  vec_tail_kernel<T, NST><<<(int)grid, kTailThreads, se.smem_bytes, stream>>>(a);
  *grid_out = (int)a.n_items;  // number of partials written
  return ST_OK;
}

template <typename T>
static int launch_tail(VecArgs<T>& a, const StratEntry& se, int nst, int64_t len, int* grid_out, cudaStream_t stream) {
  switch (nst) {
    case 2: return launch_tail_t<T, 2>(a, se, len, grid_out, stream);
    case 3: return launch_tail_t<T, 3>(a, se, len, grid_out, stream);
    case 6: return launch_tail_t<T, 6>(a, se, len, grid_out, stream);
    case 8: return launch_tail_t<T, 8>(a, se, len, grid_out, stream);
    default: return launch_tail_t<T, 4>(a, se, len, grid_out, stream);
  }
}

// Launch the main pass over [begin, end): writes `*grid_out` fp64 partials to `partials`.
template <typename T>
static int vec_partials(int layout, int rank, int64_t dim, const T* d_packed, int64_t begin, int64_t end, const T* d_x,
                        double* partials, int* grid_out, cudaStream_t stream) {
  if (layout != ST_LAYOUT_PERMCLS && layout != ST_LAYOUT_FLAT) { set_error("unknown layout %d", layout); return ST_ERR_INVALID; }
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  const int64_t total = layout == ST_LAYOUT_PERMCLS ? P.total : P.flat_size;
  if (begin < 0 || end < begin || end > total) { set_error("range [%lld, %lld) outside [0, %lld]", (long long)begin, (long long)end, (long long)total); return ST_ERR_INVALID; }
  if (begin % ST_CLASS_ALIGN) { set_error("begin must be a multiple of %d", ST_CLASS_ALIGN); return ST_ERR_INVALID; }
  if (!partials || (end > begin && (!d_packed || (dim > 0 && !d_x)))) { set_error("null pointer"); return ST_ERR_INVALID; }
  VecArgs<T> a;
  a.P = P;
  a.strat = nullptr;
  a.A = d_packed;
  a.x = d_x;
  a.begin = begin;
  a.end = end;
  a.partials = partials;
  a.n_items = 1;
  a.item_elems = end - begin;
  a.tbl_cap = 0;
  a.counter = nullptr;
  *grid_out = 1;
  if (end == begin) return check_cuda(cudaMemsetAsync(partials, 0, sizeof(double), stream), "cudaMemsetAsync");
  StratEntry se;
  se.supported = false;
  int nst = 6;
  if (layout == ST_LAYOUT_PERMCLS && rank > 0 && g_variant != 1) {
    // the tables come first: take the tail lengths the cost model picks with the smallest ring, then the deepest
    // ring (up to g_ring_stages; 4 stages = 3 KB in flight per warp) that still leaves room for those tables
    StratEntry base;
    rc = get_strategy(rank, dim, (int)sizeof(T), 2, &base);
    if (rc) return rc;
    se = base;
    nst = 2;
    if (base.supported) {
      const int cand[4] = {8, 6, 4, 3};
      for (int i = 0; i < 4; ++i) {
        if (cand[i] > g_ring_stages) continue;
        StratEntry e2;
        rc = get_strategy(rank, dim, (int)sizeof(T), cand[i], &e2);
        if (rc) return rc;
        if (e2.supported && e2.tbl_cap == base.tbl_cap) { se = e2; nst = cand[i]; break; }
      }
    }
    if (!se.supported && g_variant == 2) { set_error("tail-table kernel unavailable for dim %lld", (long long)dim); return ST_ERR_UNSUPPORTED; }
  }
  if (se.supported) {
    a.strat = se.d_strat;
    a.tbl_cap = se.tbl_cap;
    rc = launch_tail<T>(a, se, nst, end - begin, grid_out, stream);
    if (rc) return rc;
    count_launch();
    return check_cuda(cudaGetLastError(), "vec_tail_kernel");
  }
  const int64_t n = end - begin;
  int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  grid = std::min(grid, kMaxCtas);
  if (layout == ST_LAYOUT_PERMCLS) vec_generic_kernel<T, ST_LAYOUT_PERMCLS><<<grid, 256, 0, stream>>>(a);
  else vec_generic_kernel<T, ST_LAYOUT_FLAT><<<grid, 256, 0, stream>>>(a);
  count_launch();
  *grid_out = grid;
  return check_cuda(cudaGetLastError(), "vec_generic_kernel");
}

template <typename T>
static int vec_finalize(const double* partials, int n, T* d_out, cudaStream_t stream) {
  if (sizeof(T) == 8) vec_finalize_kernel<<<1, 256, 0, stream>>>(partials, n, reinterpret_cast<double*>(d_out), nullptr);
  else vec_finalize_kernel<<<1, 256, 0, stream>>>(partials, n, nullptr, reinterpret_cast<float*>(d_out));
  count_launch();
  return check_cuda(cudaGetLastError(), "vec_finalize_kernel");
}

template <typename T>
static int contract_vec(int layout, int rank, int64_t dim, const T* d_packed, int64_t begin, int64_t end, const T* d_x,
                        T* d_out, void* d_ws, cudaStream_t stream) {
  if (!d_out || !d_ws) { set_error("null pointer"); return ST_ERR_INVALID; }
  double* partials = reinterpret_cast<double*>(d_ws);
  int grid = 1;
  int rc = vec_partials<T>(layout, rank, dim, d_packed, begin, end, d_x, partials, &grid, stream);
  if (rc) return rc;
  return vec_finalize<T>(partials, grid, d_out, stream);
}

// ---- host-buffer entry: chunked, double-buffered host->device streaming overlapped with the kernel ----
struct HostStage {
  int dev = -1;
  size_t chunk_bytes = 0;
  char* d_buf[2] = {nullptr, nullptr};
  char* d_x = nullptr;
  size_t x_bytes = 0;
  double* d_ws = nullptr;
  size_t ws_slots = 0;
  char* d_out = nullptr;
  char* h_out = nullptr;  // pinned
  cudaStream_t s_copy = nullptr, s_comp = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, computed[2] = {nullptr, nullptr};
};
static std::mutex g_hmu;
static std::map<int, HostStage*> g_stages;
static const size_t kChunkBytes = (size_t)64 << 20;

static int get_stage(size_t x_bytes, size_t ws_slots, HostStage** out) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  HostStage*& st = g_stages[dev];
  if (!st) {
    st = new HostStage();
    st->dev = dev;
    st->chunk_bytes = kChunkBytes;
    for (int b = 0; b < 2 && !rc; ++b) rc = check_cuda(cudaMalloc(&st->d_buf[b], kChunkBytes), "cudaMalloc(stage)");
    if (!rc) rc = check_cuda(cudaMalloc(&st->d_out, 16), "cudaMalloc(out)");
    if (!rc) rc = check_cuda(cudaMallocHost(&st->h_out, 16), "cudaMallocHost(out)");
    if (!rc) rc = check_cuda(cudaStreamCreateWithFlags(&st->s_copy, cudaStreamNonBlocking), "cudaStreamCreate");
    if (!rc) rc = check_cuda(cudaStreamCreateWithFlags(&st->s_comp, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int b = 0; b < 2 && !rc; ++b) {
      rc = check_cuda(cudaEventCreateWithFlags(&st->copied[b], cudaEventDisableTiming), "cudaEventCreate");
      if (!rc) rc = check_cuda(cudaEventCreateWithFlags(&st->computed[b], cudaEventDisableTiming), "cudaEventCreate");
    }
    if (rc) return rc;
  }
  if (st->x_bytes < x_bytes) {
    if (st->d_x) cudaFree(st->d_x);
    st->x_bytes = std::max<size_t>(x_bytes, 4096);
    rc = check_cuda(cudaMalloc(&st->d_x, st->x_bytes), "cudaMalloc(x)");
    if (rc) return rc;
  }
  if (st->ws_slots < ws_slots) {
    if (st->d_ws) cudaFree(st->d_ws);
    st->ws_slots = ws_slots;
    rc = check_cuda(cudaMalloc(&st->d_ws, ws_slots * sizeof(double)), "cudaMalloc(ws)");
    if (rc) return rc;
  }
  *out = st;
  return ST_OK;
}

template <typename T>
static int contract_vec_host(int layout, int rank, int64_t dim, const T* h_packed, int64_t total, const T* h_x, T* h_out) {
  if (!h_out || (total > 0 && !h_packed) || (dim > 0 && !h_x)) { set_error("null pointer"); return ST_ERR_INVALID; }
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return ST_ERR_INVALID;
  const int64_t expect = layout == ST_LAYOUT_PERMCLS ? hp->h_offsets[hp->ncls] : hp->flat_size;
  if (total != expect) { set_error("packed length %lld, expected %lld", (long long)total, (long long)expect); return ST_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(g_hmu);
  const int64_t chunk_elems = (int64_t)(kChunkBytes / sizeof(T)) / 32768 * 32768;
  const int64_t n_chunks = std::max<int64_t>(1, (total + chunk_elems - 1) / chunk_elems);
  HostStage* st = nullptr;
  int rc = get_stage((size_t)dim * sizeof(T), (size_t)n_chunks * kMaxPartials, &st);
  if (rc) return rc;
  rc = check_cuda(cudaMemsetAsync(st->d_ws, 0, (size_t)n_chunks * kMaxPartials * sizeof(double), st->s_comp), "cudaMemsetAsync(ws)");
  if (rc) return rc;
  if (dim > 0) {
    rc = check_cuda(cudaMemcpyAsync(st->d_x, h_x, (size_t)dim * sizeof(T), cudaMemcpyHostToDevice, st->s_comp), "cudaMemcpyAsync(x)");
    if (rc) return rc;
  }
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int b = (int)(c & 1);
    const int64_t begin = c * chunk_elems;
    const int64_t end = std::min(total, begin + chunk_elems);
    if (c >= 2) { rc = check_cuda(cudaStreamWaitEvent(st->s_copy, st->computed[b], 0), "cudaStreamWaitEvent"); if (rc) return rc; }
    if (end > begin) {
      rc = check_cuda(cudaMemcpyAsync(st->d_buf[b], h_packed + begin, (size_t)(end - begin) * sizeof(T), cudaMemcpyHostToDevice, st->s_copy),
                      "cudaMemcpyAsync(chunk)");
      if (rc) return rc;
    }
    rc = check_cuda(cudaEventRecord(st->copied[b], st->s_copy), "cudaEventRecord");
    if (!rc) rc = check_cuda(cudaStreamWaitEvent(st->s_comp, st->copied[b], 0), "cudaStreamWaitEvent");
    if (rc) return rc;
    int grid = 1;
    rc = vec_partials<T>(layout, rank, dim, reinterpret_cast<const T*>(st->d_buf[b]), begin, end, reinterpret_cast<const T*>(st->d_x),
                         st->d_ws + c * kMaxPartials, &grid, st->s_comp);
    if (rc) return rc;
    rc = check_cuda(cudaEventRecord(st->computed[b], st->s_comp), "cudaEventRecord");
    if (rc) return rc;
  }
  rc = vec_finalize<T>(st->d_ws, (int)(n_chunks * kMaxPartials), reinterpret_cast<T*>(st->d_out), st->s_comp);
  if (rc) return rc;
  rc = check_cuda(cudaMemcpyAsync(st->h_out, st->d_out, sizeof(T), cudaMemcpyDeviceToHost, st->s_comp), "cudaMemcpyAsync(out)");
  if (rc) return rc;
  rc = check_cuda(cudaStreamSynchronize(st->s_comp), "cudaStreamSynchronize");
  if (rc) return rc;
  memcpy(h_out, st->h_out, sizeof(T));
  return ST_OK;
}

}  // namespace st

using namespace st;

extern "C" {

int64_t st_contract_vec_workspace_bytes(void) { return (int64_t)sizeof(double) * kMaxPartials; }

int st_set_tuning(const char* key, int64_t value) {
  if (!key) { set_error("null key"); return ST_ERR_INVALID; }
  const std::string k(key);
  if (k == "vec_ring_stages" && (value == 2 || value == 3 || value == 4 || value == 6 || value == 8)) { g_ring_stages = (int)value; return ST_OK; }
  if (k == "vec_items_per_cta" && value >= 1 && value <= 64) { g_items_per_cta = (int)value; return ST_OK; }
  if (k == "vec_force_tau" && value >= 0 && value <= ST_MAX_RANK) {
    g_force_tau = (int)value;
    std::lock_guard<std::mutex> lk(g_smu);
    g_strats.clear();  // strategies are rebuilt on next use (device tables of old entries are leaked: test hook)
    return ST_OK;
  }
  set_error("unknown tuning key '%s' or value %lld out of range", key, (long long)value);
  return ST_ERR_INVALID;
}

int st_set_vec_variant(int variant) {
  if (variant < 0 || variant > 2) { set_error("variant must be 0, 1 or 2"); return ST_ERR_INVALID; }
  g_variant = variant;
  return ST_OK;
}

int st_contract_vec_f64(int layout, int rank, int64_t dim, const double* d_packed, int64_t begin, int64_t end,
                        const double* d_x, double* d_out, void* d_workspace, void* stream) {
  return contract_vec<double>(layout, rank, dim, d_packed, begin, end, d_x, d_out, d_workspace, (cudaStream_t)stream);
}

int st_contract_vec_f32(int layout, int rank, int64_t dim, const float* d_packed, int64_t begin, int64_t end,
                        const float* d_x, float* d_out, void* d_workspace, void* stream) {
  return contract_vec<float>(layout, rank, dim, d_packed, begin, end, d_x, d_out, d_workspace, (cudaStream_t)stream);
}

int st_contract_vec_host_f64(int layout, int rank, int64_t dim, const double* h_packed, int64_t total, const double* h_x,
                             double* h_out) {
  return contract_vec_host<double>(layout, rank, dim, h_packed, total, h_x, h_out);
}

int st_contract_vec_host_f32(int layout, int rank, int64_t dim, const float* h_packed, int64_t total, const float* h_x,
                             float* h_out) {
  return contract_vec_host<float>(layout, rank, dim, h_packed, total, h_x, h_out);
}

}  // extern "C"
