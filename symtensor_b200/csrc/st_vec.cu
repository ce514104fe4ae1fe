// contract_all_indices_with_vector on the packed layouts (symtensor/symalg.py:505-527).
//
// The reference computes  s = sum_{i1..ir} A[i1..ir] x[i1]...x[ir]  as r rounds of
// [todense -> np.tensordot(., x, 1) -> r!-symmetrize -> repack].  In packed space the same number is
//     s = sum_classes gamma_c * sum_p A_c[p] * prod_j x[v_j(p)]^{m_j}                       
// i.e. ONE streaming pass over the packed buffer: every stored value is read from HBM exactly once, so the
// kernel is HBM-bound (algorithmic bytes = sizeof(T) per packed component).
//
// Two kernels:
//  * vec_ring_kernel (the production path for ST_LAYOUT_PERMCLS).  Storage order is lexicographic in the distinct
//    values, so for a fixed "head" (all values but the last tau values of the last run) the tail components
//    are CONTIGUOUS in memory and their weights are a contiguous slice of a head-independent table
//    T[q] = prod_{u in q-th tau-combination} xrel[u]^mu  kept in shared memory (for tau == 2 at large dimensions: the
//    suffix of that table that fits, the long first rows being formed on the fly).  A warp walks the heads of
//    its tiles with a warp-uniform odometer; the components arrive through per-warp rings of cp.async.bulk
//    copies, tiles are dealt to the warps dynamically, and every tile's sum has its own slot, so the result does
//    not depend on the deal.
//  * vec_generic_kernel (one full unrank per element; any layout; also the flat-layout path and the
//    cross-check used by the tests).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "st_common.cuh"
#include "st_vec_core.cuh"

namespace st {

static const int kMaxCtas = 148 * 8;
static const int kMaxPartials = 4096;   // fp64 partial sums per launch (one per warp of the grid / per CTA)
// vec_ring_kernel's workspace: [0] the launch's sum, [8, 8 + warps) per-warp sums (small classes), then one 32-byte
// slot per tile of the whole tensor (kTilePartOff + 4 * directory index); the tile size grows so that the tiles fit.
// Every store to it is a whole, aligned 32-byte sector: 8-byte stores scattered over time leave partially written
// sectors behind, and reading those back (and fencing them) cost the last CTA ~15 us
static const int kWarpPartOff = 8;       // per-warp sums start here (64-byte aligned: a CTA's sums are whole sectors)
static const int kTilePartOff = kMaxPartials + 16;  // per-tile sums: one 32-byte sector each (the sum, then zeros)
static const int64_t kWsSlots = (int64_t)1 << 19;  // 4 MB of fp64 slots
static const int kMaxCounters = 2 + 256;  // ticket pair + one tile counter per class
static const int kBinomSmemMax = 40 * 1024;  // the binomial table is copied to shared memory when it is at most this big

template <typename T>
struct VecArgs {
  PlanView P;
  const TailStrategy* strat;
  const T* A;        // points at packed coordinate `begin`
  const T* x;
  int64_t begin, end;
  double* partials;  // [warps of the grid] (tail kernel) / [gridDim.x] (generic kernel)
  T* out;            // fused finalize: the last CTA to finish adds the partials in index order (nullptr: caller finalizes)
  int64_t tile_elems;  // components per tile (multiple of ST_CLASS_ALIGN)
  const DirEntry* dir;         // tile directory (nullptr: unrank every tile start)
  const int64_t* tile_base;    // [ncls + 1] first directory entry of each class
  const DirEntry* sdir;        // per-component directory of the small classes (nullptr: unrank)
  const int64_t* sbase;        // [ncls + 1] first sdir entry of each class
  unsigned long long* counter;  // finished CTAs (library-owned, per stream, self-resetting)
  int32_t tbl_cap;     // table entries that fit the dynamic shared memory
  int32_t binom_smem;  // entries of the binomial table to copy to shared memory (0: use the global copy)
  int32_t cdesc_smem;  // copy the class descriptors to shared memory
  double* sum_out;     // vec_ring_kernel: where the launch's sum goes (fp64)
  int32_t dynamic;     // vec_ring_kernel: deal the tiles of mode-A classes dynamically (per-class counters behind `counter`)
  int32_t ondemand;    // vec_ring_kernel: rounds at the end of a class whose entries are claimed on demand
  unsigned long long* tl;  // debug timeline (nullptr: off): [cta][16] phase stamps, then [warp of the grid] finish stamps
  int32_t priv_cap;    // vec_ring_kernel: entries of the per-warp xr tables (nwarps * dim, or 0)
  int32_t ring_slots;  // vec_ring_kernel: slots per warp (R)
  int32_t ring_elems;  // vec_ring_kernel: components per slot (Bel; Bel * sizeof(T) is a multiple of 16)
  int32_t pdl;         // vec_ring_kernel: launched as a programmatic dependent of the previous launch on the stream (ST_VEC_OVERLAP)
};

// The launch's class records and tile schedule as kernel parameters (rank <= 8: at most 22 classes): computed on
// the host in microseconds, they would cost every CTA several microseconds of serial 64-bit divisions.
static const int kMaxSchedCls = 24;
struct RingSched {
  int32_t n;  // 0: not provided (too many classes) -- the kernel builds them itself
  int32_t pad_;
  ClsInfo cls[kMaxSchedCls];
  ClsRun run[kMaxSchedCls];
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
    const double v = (double)__ldcs(a.A + (c - a.begin));
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

// GPU index enumerator, bulk form: the values of the component at the start of every tile
__global__ void vec_dir_kernel(PlanView P, int64_t tile, const int64_t* __restrict__ tile_base, int64_t ntiles, DirEntry* __restrict__ dir) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  int lo = 0, hi = P.ncls - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_base[mid] <= t) lo = mid; else hi = mid - 1;
  }
  const ClassDesc& C = P.cls[lo];
  const int64_t pos = (t - tile_base[lo]) * tile;
  DirEntry e;
  for (int i = 0; i < ST_MAX_RANK; ++i) e.v[i] = 0;
  if (pos < C.size) {
    int32_t vals[ST_MAX_RANK];
    permcls_unrank_vals(P, C, pos, vals);
    for (int i = 0; i < C.nvals; ++i) e.v[i] = (uint16_t)vals[i];
  }
  dir[t] = e;
}

// CTA-wide build of the shared tail table (xr must be in place and visible); ends with a barrier
template <typename T>
__device__ __forceinline__ void build_shared_table(const PlanView& P, const TailStrategy& S, const T* xr, T* tbl, int nthreads) {
  int64_t nA, nB;
  table_scratch(P.binom, P.rank, S.Rt, S.tau, &nA, &nB);
  for (int t = 2; t <= S.tau; ++t) {
    const T* src = t == 2 ? xr : table_level_buffer<T>(tbl, S.tbl_n, nA, S.tau, t - 1);
    T* dst = table_level_buffer<T>(tbl, S.tbl_n, nA, S.tau, t);
    build_table_level<T>(P.binom, P.rank, S.Rt, t, xr, src, dst, threadIdx.x, nthreads);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
// ring kernel (production path): per-warp TMA rings
// ------------------------------------------------------------------------------------------------------
// Same schedule and tables as the tail-table walk described above, but the components never pass through
// registers on their way in: every warp owns a RING of R slots of B bytes in shared memory and keeps it full
// with 1-D bulk copies (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier per slot.  The copies are
// issued by the warp itself -- a slot is refilled the moment its sub-chunk has been consumed -- so there is no
// producer warp, no "empty" barrier, and the bytes in flight per SM (NW x R x B, ~48-96 KB) do not depend on
// how fast the arithmetic runs: the index arithmetic (odometer steps, table rebuilds, class changes) overlaps
// with the memory stream instead of stalling it.  A warp's stream is the concatenation of its tiles, so the
// ring also prefetches ACROSS tiles and classes (the first copies are issued before the tables are built).
// Final reduction of vec_ring_kernel (last CTA): this thread's share of the concatenated slot ranges in ctl.  Its own
// function, not inlined: inside the kernel the register allocator serialised the loads (one L2 round trip per
// slot, ~10 us); here 16 slots are requested before the first is added.
__device__ __noinline__ double reduce_slots(const TailCtrl* ctl, int tid, int nthreads, bool first_pass) {
  // the ranges as one concatenated index space; this thread's indices tid, tid + nthreads, ... only grow, so the
  // range of an index is tracked incrementally.  Three separate phases per batch -- addresses, loads, adds -- with
  // UNCONDITIONAL loads (an index past the end is clamped to the last slot and its value dropped), so that all the
  // loads of a batch are in flight together: with predicated loads inside the address loop the compiler kept only a
  // few in flight and the reduction of ~8.7 K slots took 6.5 us (measured with timeline stamps).
  constexpr int NB = 20;
  const int n = ctl->red_n;
  const int64_t nslots = ctl->red_start[n];
  double s = 0.0;
  if (nslots <= 0) return s;
  int r = 0;
  int64_t r_lo = 0, r_hi = ctl->red_start[1];
  const double* r_ptr = ctl->red_ptr[0];
  int stride = first_pass ? 1 : 4;  // per-warp sums are dense, tile sums one per 32-byte sector
  for (int64_t base = tid; base < nslots; base += NB * (int64_t)nthreads) {
    const double* p[NB];
    unsigned ok = 0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      int64_t idx = base + j * (int64_t)nthreads;
      if (idx < nslots) ok |= 1u << j; else idx = nslots - 1;
      while (idx >= r_hi) {
        ++r;
        r_lo = r_hi;
        r_hi = ctl->red_start[r + 1];
        r_ptr = ctl->red_ptr[r];
        stride = 4;
      }
      p[j] = r_ptr + stride * (idx - r_lo);
    }
    double v[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) v[j] = __ldcg(p[j]);
#pragma unroll
    for (int j = 0; j < NB; ++j) s += ((ok >> j) & 1u) ? v[j] : 0.0;
  }
  return s;
}

template <typename T>
__global__ void __launch_bounds__(512, 1) vec_ring_kernel(const __grid_constant__ VecArgs<T> a, const __grid_constant__ RingSched sched) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PlanView P = a.P;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nthreads = blockDim.x, nwarps = blockDim.x >> 5;
  const int G = gridDim.x;
  const RingLayout L = ring_layout((int)sizeof(T), P.dim, a.tbl_cap, a.priv_cap, a.binom_smem, P.ncls, a.cdesc_smem, nwarps, a.ring_slots, a.ring_elems);
  T* tbl = reinterpret_cast<T*>(smem_raw);
  T* priv = reinterpret_cast<T*>(smem_raw + L.priv) + (size_t)warp * P.dim;  // this warp's private xr table (classes with earlier runs)
  T* xr = reinterpret_cast<T*>(smem_raw + L.xr);
  T* xs = xr + P.dim;
  int32_t* blen = reinterpret_cast<int32_t*>(smem_raw + L.blen);
  int64_t* binom_s = reinterpret_cast<int64_t*>(smem_raw + L.binom);
  ClsInfo* cls_s = reinterpret_cast<ClsInfo*>(smem_raw + L.cls);
  ClassDesc* cdesc_s = reinterpret_cast<ClassDesc*>(smem_raw + L.cdesc);
  WarpScratch& ws = reinterpret_cast<WarpScratch*>(smem_raw + L.ws)[warp];
  TailCtrl* ctl = reinterpret_cast<TailCtrl*>(smem_raw + L.ctl);
  ClsRun* run = reinterpret_cast<ClsRun*>(smem_raw + L.run);
  const int64_t tile = a.tile_elems;
  auto stamp = [&](int slot) {
    if (a.tl != nullptr && threadIdx.x == 0) {
      unsigned long long ts;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
      a.tl[(size_t)blockIdx.x * 16 + slot] = ts;
    }
  };
  stamp(0);
  // Programmatic dependent launch (ST_VEC_OVERLAP): this launch may have started while the previous launch on the
  // stream -- another vector contraction -- was still draining its tail.  Everything a launch writes before `pdl_sync`
  // (tile slots, claim counters) lives in the half of the workspace / the counter set of ITS parity, which the launch
  // before the previous one used; that one is complete, because a CTA lets the next launch go (launch_dependents) only
  // after it has seen the previous launch complete (wait).  So at most two launches are ever in flight.
  bool pdl_done = false;
  auto pdl_sync = [&]() {
    if (!pdl_done) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      pdl_done = true;
    }
  };

  // ---- 1. class records and the launch's tile schedule (from the kernel parameters when the host could put them
  // there), barriers; the stream starts at once
  if (threadIdx.x == 0) { ctl->cur_cls = -1; ctl->cur_seg = -1; }
  const uint32_t bar0 = smem_u32(smem_raw + L.bar) + 8u * (uint32_t)(warp * a.ring_slots);
  if (lane == 0) for (int r = 0; r < a.ring_slots; ++r) mbar_init(bar0 + 8u * r, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (sched.n > 0) {
    for (int c = threadIdx.x; c < P.ncls; c += nthreads) {
      cls_s[c] = sched.cls[c];
      run[c] = sched.run[c];
    }
    __syncthreads();
  } else {
    for (int c = threadIdx.x; c < P.ncls; c += nthreads) {
      cls_s[c].offset = a.P.cls[c].offset;
      cls_s[c].size = a.P.cls[c].size;
      cls_s[c].tile_base = a.tile_base ? a.tile_base[c] : 0;
      cls_s[c].sbase = a.sbase ? a.sbase[c] : 0;
      cls_s[c].S = a.strat[c];
    }
    __syncthreads();
    if (threadIdx.x == 0) make_runs(cls_s, P.ncls, a.begin, a.end, tile, nwarps, G, run);
    __syncthreads();
  }
  RingSrc<T> src;
  src.ring = reinterpret_cast<T*>(smem_raw + L.ring) + (size_t)warp * a.ring_slots * a.ring_elems;
  src.bar0 = bar0;
  src.R = a.ring_slots;
  src.Bel = a.ring_elems;
  src.lane = lane;
  src.run = run;
  src.cls = cls_s;
  src.A = a.A;
  src.dir = a.dir;
  src.ctr = a.dynamic ? a.counter + 2 : nullptr;
  src.QD = a.ring_slots + 2;
  src.queue = reinterpret_cast<TileQ*>(smem_raw + L.queue) + (size_t)warp * (a.ring_slots + 2);
  src.begin = a.begin;
  src.tile = tile;
  src.ondemand = a.ondemand;
  src.ncls = P.ncls;
  src.NW = nwarps;
  src.G = G;
  src.warp = warp;
  src.cta = blockIdx.x;
  // x is requested now and stored after the stream has started (its latency overlaps the first claims and copies)
  const T x_pre = (int64_t)threadIdx.x < P.dim ? a.x[threadIdx.x] : T(0);
  constexpr int kPre = 3;
  int64_t b_pre[kPre];
  int32_t c_pre[kPre];
  const int nw32 = a.cdesc_smem ? (int)(sizeof(ClassDesc) / 4) * P.ncls : 0;
#pragma unroll
  for (int j = 0; j < kPre; ++j) {
    const int i = threadIdx.x + j * nthreads;
    b_pre[j] = i < a.binom_smem ? a.P.binom[i] : 0;
    c_pre[j] = i < nw32 ? reinterpret_cast<const int32_t*>(a.P.cls)[i] : 0;
  }
  stamp(1);
  src.start();
  stamp(2);

  // ---- 2. x, binomials, class descriptors
  if ((int64_t)threadIdx.x < P.dim) xs[threadIdx.x] = x_pre;
  for (int i = threadIdx.x + nthreads; i < P.dim; i += nthreads) xs[i] = a.x[i];
#pragma unroll
  for (int j = 0; j < kPre; ++j) {
    const int i = threadIdx.x + j * nthreads;
    if (i < a.binom_smem) binom_s[i] = b_pre[j];
    if (i < nw32) reinterpret_cast<int32_t*>(cdesc_s)[i] = c_pre[j];
  }
  if (a.binom_smem) {
    for (int i = threadIdx.x + kPre * nthreads; i < a.binom_smem; i += nthreads) binom_s[i] = a.P.binom[i];
    P.binom = binom_s;
  }
  if (a.cdesc_smem) {
    for (int i = threadIdx.x + kPre * nthreads; i < nw32; i += nthreads) reinterpret_cast<int32_t*>(cdesc_s)[i] = reinterpret_cast<const int32_t*>(a.P.cls)[i];
    P.cls = cdesc_s;
  }
  __syncthreads();
  // the first class with a class-wide table (no earlier runs): its table is built now and stays for the whole launch
  // (the classes with per-warp tables do not touch it), so no warp ever waits for another at a class change
  for (int ci = 0; ci < P.ncls; ++ci) {
    const TailStrategy& S = cls_s[ci].S;
    if (run[ci].mode != 1 || S.tau < 2 || S.nE != 0) continue;
    if (threadIdx.x == 0) {
      ctl->wE = unrank_earlier<T>(P, P.cls[ci], 0, xs, ctl->E, ws);
      ctl->cur_cls = ci;
      ctl->cur_seg = 0;
    }
    for (int uu = threadIdx.x; uu < S.Rt; uu += nthreads) {
      xr[uu] = xrel_pow<T>(xs, ctl->E, 0, S.mu, uu);
      blen[uu] = (int32_t)binom_at(P.binom, P.rank, S.Rt - 1 - uu, S.tau);
    }
    __syncthreads();
    if (!S.direct) {
      build_shared_table<T>(P, S, xr, tbl, nthreads);
    } else if (S.k0 < S.Rt - 1) {
      build_pair_suffix<T>(S.Rt, S.k0, xr, tbl, threadIdx.x, nthreads);
      __syncthreads();
    }
    break;
  }
  stamp(3);
  int priv_cls = -1;  // (class, segment) the private table was built for -- warp-uniform
  int64_t priv_seg = -1;
  double priv_wE = 0.0;
  const int64_t W = (int64_t)G * nwarps;                   // warps of the grid
  const int64_t gw = (int64_t)blockIdx.x * nwarps + warp;  // this warp
  double total = 0.0;

  // ---- 3. tail-table classes, in stream order.  Every tile's sum goes to its own slot of the workspace (the deal
  // may be dynamic: the final reduction must not depend on who walked which tile).
  unsigned long long dbg_tiles = 0, dbg_maxdur = 0, dbg_last = 0, dbg_maxtile = 0;  // debug timeline only
  double* tile_part = a.partials + kTilePartOff;
  auto store_tile = [&](int64_t slot, double v) {
    v = warp_sum(v);
    if (lane < 4) tile_part[slot * 4 + lane] = lane == 0 ? v : 0.0;  // one aligned 32-byte sector
  };
  for (int ci = 0; ci < P.ncls; ++ci) {
    const ClsRun rr = run[ci];
    if (!rr.mode) continue;
    if (ci < 5 && ci != 1) stamp(5 + ci);  // (slot 6 is the end of the last CTA's reduce_slots)
    const TailStrategy S = cls_s[ci].S;
    const ClassDesc& C = P.cls[ci];
    const int64_t coff = cls_s[ci].offset, csize = cls_s[ci].size;
    const T* Acls = a.A + (coff - a.begin);
    const int64_t lo = rr.lo, hi = rr.hi, k0 = rr.k0, k1 = rr.k1, ch0 = rr.ch0, ch1 = rr.ch1;
    const int64_t dbase = cls_s[ci].tile_base;
    if (rr.mode == 1) {
      // ---- mode A: tiles are per warp; tables are per class (no earlier runs) or per warp (tau == 1 / direct)
      const bool shared_tbl = S.tau >= 2 && S.nE == 0;
      if (shared_tbl && ctl->cur_cls != ci) {  // CTA-uniform; the first such class was set up in the prologue
        __syncthreads();  // everybody is done with the previous tables
        if (threadIdx.x == 0) {
          ctl->wE = unrank_earlier<T>(P, C, 0, xs, ctl->E, ws);
          ctl->cur_cls = ci;
          ctl->cur_seg = 0;
        }
        for (int uu = threadIdx.x; uu < S.Rt; uu += nthreads) {
          xr[uu] = xrel_pow<T>(xs, ctl->E, 0, S.mu, uu);
          blen[uu] = (int32_t)binom_at(P.binom, P.rank, S.Rt - 1 - uu, S.tau);
        }
        __syncthreads();
        if (!S.direct) {
          build_shared_table<T>(P, S, xr, tbl, nthreads);
        } else if (S.k0 < S.Rt - 1) {
          build_pair_suffix<T>(S.Rt, S.k0, xr, tbl, threadIdx.x, nthreads);
          __syncthreads();
        }
      }
      const double wE0 = ctl->wE;
      int32_t* E = ws.E;
      int32_t* u0 = ws.u0;
      while (src.head_is(ci)) {
        unsigned long long t_begin = 0;
        if (a.tl != nullptr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
        const TileQ& tq = src.pop();
        const int64_t tk = tq.tk;
        int32_t tn;
        const int64_t dn = tile_to_deal(rr, tk, tn);  // deal entry (its slot) and tile count (grouped deal)
        int64_t w0 = tk * tile;
        int64_t w1 = w0 + tn * tile;
        const bool have_dir = a.dir != nullptr && w0 >= lo;  // the tile starts inside the launch range
        if (w0 < lo) w0 = lo;
        if (w1 > hi) w1 = hi;
        const int64_t tw0 = w0;
        src.open_tile(Acls + w0, (int)(w1 - w0));
        const DirEntry& de = tq.de;
        double tsum = 0.0;
        if (shared_tbl) {
          if (have_dir) {  // single-run class: the entry is the combination itself
            __syncwarp();
            if (lane < S.gt) ws.u[lane] = de.v[lane];
            __syncwarp();
          }
          tsum = walk_tile_any<T>(P, S, tbl, xr, blen, wE0, w0, (int)(w1 - w0), 0, lane, have_dir ? ws.u : nullptr, src, ws);
        } else {
          const bool gap = S.direct && S.nE != 0 && S.mu == 1;
          bool first = true;
          while (w0 < w1) {  // private tables: segment by segment
            const int64_t sidx = S.nE ? w0 / S.seg : 0;
            const int64_t sbase = sidx * S.seg;
            const int64_t q1 = (S.seg < w1 - sbase) ? S.seg : w1 - sbase;
            const bool from_dir = first && have_dir;
            if (from_dir) {
              const double wE = dir_decode<T>(C, S, de, xs, E, u0);
              if (priv_cls != ci || priv_seg != sidx) priv_wE = wE;
            } else if (w0 == sbase) {
              for (int i = 0; i < S.gt; ++i) u0[i] = i;  // a segment starts with the first combination
            }
            if (priv_cls != ci || priv_seg != sidx) {
              if (!from_dir) priv_wE = unrank_earlier<T>(P, C, sidx, xs, E, ws);
              __syncwarp();
              if (!gap) {  // (gap classes relabel x on the fly from ws.E)
                for (int uu = lane; uu < S.Rt; uu += 32) priv[uu] = xrel_pow<T>(xs, E, S.nE, S.mu, uu);
                __syncwarp();
              }
              priv_cls = ci;
              priv_seg = sidx;
            }
            tsum += walk_tile_any<T>(P, S, gap ? xs : priv, gap ? xs : priv, nullptr, priv_wE, w0 - sbase, (int)(sbase + q1 - w0), (int)(w0 - tw0), lane,
                                     (from_dir || w0 == sbase) ? u0 : nullptr, src, ws);
            w0 = sbase + q1;
            first = false;
          }
        }
        src.close_tile();
        if (warp == 0) pdl_sync();  // (by now the previous launch is long complete: no stall)
        // a statically dealt entry belongs to this warp whatever happens: its sum needs no slot of its own (the last CTA
        // reads one 32-byte sector per slot, ~0.8 ns each on its one SM) and goes into the warp's sum
        if (a.dynamic && dn < rr.ns) total += tsum;
        else store_tile(dbase + k0 + dn, tsum);
        if (a.tl != nullptr) {
          unsigned long long t_end;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
          ++dbg_tiles;
          if (t_end - t_begin > dbg_maxdur) { dbg_maxdur = t_end - t_begin; dbg_maxtile = (unsigned long long)(dbase + tk); }
          dbg_last = t_begin;
        }
      }
    } else {
      // ---- mode B: chunks of nwarps tiles per CTA (static deal); the CTA rebuilds T once per segment
      const int64_t chunk = tile * nwarps;
      for (int64_t jc = src.jc0(ci); ch0 + jc < ch1; jc += G) {
        const int64_t tk = (ch0 + jc) * nwarps + warp;  // this warp's tile of the chunk
        const bool exists = tk >= k0 && tk < k1;
        int64_t tw0 = tk * tile, tw1 = tw0 + tile;
        if (tw0 < lo) tw0 = lo;
        if (tw1 > hi) tw1 = hi;
        const TileQ* tq = nullptr;
        if (exists) {
          tq = &src.pop();  // the producer cursor queued exactly this tile
          src.open_tile(Acls + tw0, (int)(tw1 - tw0));
        }
        double tsum = 0.0;
        int64_t pos = (ch0 + jc) * chunk;
        int64_t pend = pos + chunk;
        if (pos < lo) pos = lo;
        if (pend > hi) pend = hi;
        while (pos < pend) {
          const int64_t sidx = pos / S.seg;
          const int64_t sbase = sidx * S.seg;
          const int64_t q0 = pos - sbase;
          const int64_t q1 = (S.seg < pend - sbase) ? S.seg : pend - sbase;
          if (ctl->cur_cls != ci || ctl->cur_seg != sidx) {  // CTA-uniform
            __syncthreads();  // everybody is done with the previous tables
            if (threadIdx.x == 0) {
              const int64_t tk0 = (sbase + tile - 1) / tile;
              if (a.dir != nullptr && tk0 * tile < sbase + S.seg && tk0 * tile < csize) {
                ws.de = a.dir[dbase + tk0];
                ctl->wE = dir_decode<T>(C, S, ws.de, xs, ctl->E, ws.u0);
              } else {
                ctl->wE = unrank_earlier<T>(P, C, sidx, xs, ctl->E, ws);
              }
              ctl->cur_cls = ci;
              ctl->cur_seg = sidx;
            }
            __syncthreads();
            for (int uu = threadIdx.x; uu < S.Rt; uu += nthreads) {
              xr[uu] = xrel_pow<T>(xs, ctl->E, S.nE, S.mu, uu);
              blen[uu] = (int32_t)binom_at(P.binom, P.rank, S.Rt - 1 - uu, S.tau);
            }
            __syncthreads();
            build_shared_table<T>(P, S, xr, tbl, nthreads);
          }
          if (exists) {
            int64_t w0 = tk * tile - sbase, w1 = w0 + tile;
            const bool at_tile = w0 >= q0;
            if (w0 < q0) w0 = q0;
            if (w1 > q1) w1 = q1;
            if (w0 < w1) {
              int32_t* u0 = ws.u0;
              const int32_t* ui = nullptr;
              if (w0 == 0) {
                for (int i = 0; i < S.gt; ++i) u0[i] = i;
                ui = u0;
              } else if (at_tile && a.dir != nullptr) {
                dir_decode<T>(C, S, tq->de, xs, ws.E, u0);
                ui = u0;
              }
              tsum += walk_tile_any<T>(P, S, tbl, xr, blen, ctl->wE, w0, (int)(w1 - w0), (int)(sbase + w0 - tw0), lane, ui, src, ws);
            }
          }
          pos = sbase + q1;
        }
        if (exists) {
          src.close_tile();
          store_tile(dbase + tk, tsum);
        }
        if (warp == 0) pdl_sync();
      }
    }
    if (a.tl != nullptr && lane == 0 && ci < 8) {
      unsigned long long ts;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
      a.tl[(size_t)148 * 16 + 4096 + (size_t)gw * 8 + ci] = ts;
    }
  }

  // ---- 4. SMALL classes (tau == 0 in the strategy): one component per thread from the per-component directory; last,
  // where they fill the gaps of the launch's tail instead of holding up the stream at its start
  {
    int64_t sm_base = 0;
    const int64_t nthr = W * 32, tid = gw * 32 + lane;
    for (int ci = 0; ci < P.ncls; ++ci) {
      if (cls_s[ci].S.tau != 0) continue;
      const int64_t coff = cls_s[ci].offset, csize = cls_s[ci].size;
      const int64_t lo = (a.begin > coff ? a.begin : coff) - coff;
      const int64_t hi = (a.end < coff + csize ? a.end : coff + csize) - coff;
      if (lo >= hi) continue;
      const ClassDesc& C = P.cls[ci];
      const T* Acls = a.A + (coff - a.begin);
      const int nvals = C.nvals;
      const double gamma = (double)C.gamma;
      for (int64_t p = lo + (tid - sm_base % nthr + nthr) % nthr; p < hi; p += nthr) {
        const double v = (double)__ldcs(Acls + p);
        double w = gamma;
        if (a.sdir != nullptr) {
          const uint4* q = reinterpret_cast<const uint4*>(a.sdir + (cls_s[ci].sbase + p));
          const uint4 e0 = __ldg(q), e1 = __ldg(q + 1);
          const uint32_t words[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
          for (int k = 0; k < ST_MAX_RANK; ++k) {
            if (k < nvals) {
              const int val = (int)((words[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
              const double xv = (double)xs[val];
              const int m = C.mult[k];
              for (int mm = 0; mm < m; ++mm) w *= xv;
            }
          }
        } else {
          int32_t vals[ST_MAX_RANK];
          permcls_unrank_vals(P, C, p, vals);
          for (int k = 0; k < nvals; ++k) {
            const double xv = (double)xs[vals[k]];
            for (int m = 0; m < C.mult[k]; ++m) w *= xv;
          }
        }
        total += v * w;
      }
      sm_base += hi - lo;
    }
  }

  stamp(4);
  if (a.tl != nullptr && lane == 0) {
    unsigned long long ts;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
    a.tl[(size_t)148 * 16 + (size_t)blockIdx.x * nwarps + warp] = ts;
    unsigned long long* d = a.tl + (size_t)148 * 16 + 4096 + (size_t)gw * 8;
    d[5] = dbg_tiles;
    d[6] = dbg_maxdur;
    d[7] = dbg_last;
    d[2] = dbg_maxtile;
  }
  // one partial per warp of the grid, added in index order by the last CTA to finish (deterministic)
  if (warp == 0) pdl_sync();
  total = warp_sum(total);
  if (lane == 0) ctl->red[warp] = total;
  __syncthreads();
  if (warp == 0) {  // the CTA's per-warp sums as whole sectors (slots beyond nwarps up to a multiple of 4: zeros)
    const int nw4 = (nwarps + 3) & ~3;
    if (lane < nw4) a.partials[kWarpPartOff + (int64_t)blockIdx.x * nw4 + lane] = lane < nwarps ? ctl->red[lane] : 0.0;
  }
  __syncthreads();
  stamp(10);
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(a.counter, 1ULL);
    const bool last = ticket == (unsigned long long)(G - 1);
    if (last) {
      a.counter[0] = 0ULL;
      __threadfence();
    }
    ctl->last = last ? 1 : 0;
  }
  __syncthreads();
  stamp(11);
  if (ctl->last) {
    // the last CTA adds the per-warp sums (small classes) and the per-tile sums in index order, and resets the counters
    double s = 0.0;
    // All slot ranges (per-warp sums, then the tiles of every walked class) form one concatenated index space with a
    // fixed thread -> slot map; every thread requests its (up to 32) slots before it adds any: one round trip to L2.
    int ci_next = 0;
    bool first_pass = true;
    while (first_pass || ci_next < P.ncls) {
      __syncthreads();
      if (threadIdx.x == 0) {
        int n = 0;
        int64_t pos = 0;
        if (first_pass) { ctl->red_start[0] = 0; ctl->red_ptr[0] = a.partials + kWarpPartOff; pos = (int64_t)G * ((nwarps + 3) & ~3); n = 1; }
        int c = ci_next;
        for (; c < P.ncls && n < ST_MAX_RED; ++c) {
          if (!run[c].mode) continue;
          ctl->red_start[n] = pos;
          const int64_t skip = (a.dynamic && run[c].mode == 1) ? run[c].ns : 0;  // statically dealt entries: in the per-warp sums
          ctl->red_ptr[n] = tile_part + 4 * (cls_s[c].tile_base + run[c].k0 + skip);
          pos += run[c].nd - skip;  // one slot per deal entry (mode B: per tile)
          ++n;
        }
        ctl->red_start[n] = pos;
        ctl->red_n = n;
        ctl->item_next = c;
      }
      __syncthreads();
      if (first_pass) stamp(13);
      s += reduce_slots(ctl, threadIdx.x, nthreads, first_pass);
      if (first_pass) stamp(6);
      ci_next = ctl->item_next;
      first_pass = false;
    }
    stamp(14);
    for (int c = threadIdx.x; c < P.ncls; c += nthreads) a.counter[2 + c] = 0ULL;
    stamp(12);
    s = warp_sum(s);
    if (lane == 0) ctl->red[warp] = s;
    __syncthreads();
    if (warp == 0) {
      double t = lane < nwarps ? ctl->red[lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) {
        *a.sum_out = t;
        if (a.out != nullptr) *a.out = (T)t;
      }
    }
  }
  stamp(15);
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static const size_t kClsInfoBytes = sizeof(ClsInfo);
static const size_t kRingSmemBudget = 227 * 1024 - 64;  // dynamic shared memory of vec_ring_kernel
int g_variant = 0;
static int g_use_dir = 1;  // tuning knob "vec_use_dir": 0 unranks every tile start in the kernel
int g_force_tau = 0;  // test hook: force the tail length (0 = cost model)
unsigned long long* g_timeline = nullptr;  // debug: device buffer of kTimelineSlots stamps (st_set_tuning vec_timeline 1)
static const size_t kTimelineSlots = 148 * 16 + 4096 + 4096 * 8;
int g_ring_direct = 1;                 // ring kernel: allow the table-free pair walk (2: with vec_force_tau 2, force it)
int64_t g_ring_table_max = 72 * 1024;  // ring kernel: largest tail table in bytes before the pair walk takes over
extern int64_t g_mat_onfly_rows;
extern int g_mat_dmma, g_mat_pipe, g_conv_rows, g_outer_fast, g_outer_rows, g_gram_umma, g_sym22;  // st_ops.cu
extern int64_t g_sym22_min_dim;
namespace s22 { extern int g_kch, g_debug, g_tile_rgroup; extern int64_t g_batch_tiles; void clear_tile_cache(); }  // st_sym22.cu
int64_t g_short_segment = 1024;  // classes whose segments are shorter than this take the per-component phase (tuning knob)
int64_t g_small_class = 128 * 1024;  // classes up to this many components take the per-component phase (tuning knob)

struct StratKey {
  int dev, rank, esize, ring;
  int64_t dim, tile;
  bool operator<(const StratKey& o) const { return std::tie(dev, rank, esize, ring, dim, tile) < std::tie(o.dev, o.rank, o.esize, o.ring, o.dim, o.tile); }
};
struct StratEntry {
  TailStrategy* d_strat;
  DirEntry* d_dir;        // tile directory (nullptr when disabled)
  int64_t* d_tile_base;   // [ncls + 1]
  DirEntry* d_sdir;       // per-component directory of the small classes
  int64_t* d_sbase;       // [ncls + 1]
  int32_t cdesc_smem;
  int32_t tbl_cap;
  int32_t binom_smem;
  size_t smem_bytes;
  bool supported;
  // ring kernel geometry
  int32_t nwarps, ring_slots, ring_elems, priv_cap;
  int64_t tile_elems;
  // host copies for the launch-time schedule
  std::vector<TailStrategy> h_strat;
  std::vector<int64_t> h_tile_base, h_sbase;
  std::vector<int64_t> h_ntail;  // per class: trailing tiles made of many tiny blocks (dealt first)
};
static std::mutex g_smu;
// ring kernel tuning knobs (st_set_tuning)
static int g_ring_dynamic = 1;         // deal the tiles of mode-A classes dynamically
static int g_ring_warps = 16;          // consumer warps per CTA
static int g_ring_slots = 2;           // ring slots per warp
static int g_ring_bytes = 4096;        // preferred bytes per slot (shrunk to 1536 / 1024 when the table needs the room)
static int g_ring_bytes_max = 4096;    // slots grow up to this when shared memory is left over
static int g_ring_tile_bytes = 49152;    // bytes per tile (directory granularity; rounded to whole slots)
static int64_t g_short_launch_bytes = 160ll << 20;  // launches over a SLICE of at most this many bytes use tiles of ...
static int64_t g_short_launch_slots = 4;            // ... this many ring slots (tuning keys "vec_short_launch_bytes" / "_slots")
static int g_ring_small_tiles = 1;       // shrink the tiles of tensors too small to give every warp a full-size tile
static int g_ring_group = 4;             // dynamic deal: tiles per group at the start of a class (1: every tile on its own)
static int g_ring_ondemand = 2;          // dynamic deal: rounds at the end of an ungrouped class claimed on demand
static int g_ring_fine_pos = 70;         // dynamic deal: where the single tiles sit inside a grouped class (percent of its length)
static int g_ring_fine = 2;              // dynamic deal: single tiles per warp of the grid that close a grouped class
static std::map<StratKey, StratEntry> g_strats;

static double dbinom(const HostPlan* hp, int64_t n, int k) {
  if (n < 0 || k < 0 || k > hp->rank) return 0.0;
  return (double)hp->h_binom[n * (hp->rank + 1) + k];
}

// Choose tau per class with a small cost model (estimated warp-instructions per component):
//   streaming      ~0.1 (one LDG.128, VEC table loads and FMAs per 16-byte vector)
//   per block      ~50 warp-uniform instructions (odometer step + one predicated pass over the batch)
//   table builds   ~0.5 per entry, amortised over a segment (multi-run classes) or over a CTA's share of the
//                  class (single-run classes, table built once per CTA)
// Pure host function (also used by the CPU emulation harness in tests/emu).
// `reserve` > 0 selects the ring kernel's model: that many bytes are kept for the rings (the optional shared-memory
// copies of the binomial table and the class descriptors are then decided by the caller from what is left) and
// a block costs `block_cost` warp-instructions.
bool compute_tail_strategy(const HostPlan* hp, int esize, int nwarps, std::vector<TailStrategy>& st, int32_t* tbl_cap,
                           int32_t* binom_smem, int32_t* cdesc_smem, size_t* smem_bytes, size_t reserve, double block_cost) {
  const int rank = hp->rank;
  const int64_t dim = hp->dim;
  const size_t smem_budget = reserve ? kRingSmemBudget : 224 * 1024;
  const size_t binom_bytes = (size_t)hp->binom_rows * (rank + 1) * sizeof(int64_t);
  *binom_smem = (!reserve && binom_bytes <= (size_t)kBinomSmemMax) ? (int32_t)(hp->binom_rows * (rank + 1)) : 0;
  *cdesc_smem = (!reserve && (size_t)hp->ncls * sizeof(ClassDesc) <= 24 * 1024) ? 1 : 0;
  // everything but the table: xr, xs, blen, the binomial table, class records, per-warp scratch, control block, slack
  const size_t fixed = (size_t)2 * dim * esize + (size_t)dim * 4 + (size_t)*binom_smem * sizeof(int64_t) + (size_t)hp->ncls * kClsInfoBytes +
                       (*cdesc_smem ? (size_t)hp->ncls * sizeof(ClassDesc) : 0) + (size_t)nwarps * sizeof(WarpScratch) + sizeof(TailCtrl) + 128 +
                       reserve;
  if (fixed + 32 * esize > smem_budget) return false;  // x itself does not fit shared memory
  const int64_t cap = (int64_t)((smem_budget - fixed) / esize);
  if (cap < 32) return false;
  st.assign(hp->ncls, TailStrategy());
  int64_t tbl_max = 32;
  for (int c = 0; c < hp->ncls; ++c) {
    const ClassDesc& C = hp->h_cls[c];
    TailStrategy& S = st[c];
    memset(&S, 0, sizeof(S));
    if (C.nvals == 0 || C.size == 0) { S.tau = 0; S.gt = 1; S.Rt = 1; S.mu = 1; S.tbl_n = 1; S.seg = 1; continue; }
    if (C.size <= g_small_class && g_variant != 2) { S.tau = 0; S.gt = 1; S.Rt = 1; S.mu = 1; S.tbl_n = 1; S.seg = C.size; continue; }
    // classes cut into very short segments (one per assignment of the earlier runs: a table rebuild and an unrank
    // each) are cheaper one component per thread from the per-component directory, up to 8 M components
    if (reserve > 0 && g_variant != 2 && C.nruns > 1 && C.radix[C.nruns - 1] < g_short_segment && C.size <= ((int64_t)8 << 20)) {
      S.tau = 0; S.gt = 1; S.Rt = 1; S.mu = 1; S.tbl_n = 1; S.seg = C.size; continue;
    }
    const int t = C.nruns - 1;
    S.gt = C.run_len[t];
    S.nE = C.nvals - S.gt;
    S.Rt = (int32_t)(dim - S.nE);
    S.mu = C.run_mult[t];
    S.seg = C.radix[t];
    double best = 1e300;
    int best_tau = 1;
    bool best_direct = false;
    // ring kernel: a table may take at most g_ring_table_max bytes of shared memory (the rings need the rest);
    // beyond that the pair weights are computed on the fly (tau == 2, no table)
    const bool can_direct = reserve > 0 && g_ring_direct && S.gt >= 2 && S.Rt <= 40000;
    for (int tau = 1; tau <= S.gt; ++tau) {
      int64_t nA = 0, nB = 0;
      table_scratch(hp->h_binom, rank, S.Rt, tau, &nA, &nB);
      const double tn = dbinom(hp, S.Rt, tau);
      if (tau > 1 && tn + (double)nA + (double)nB > (double)cap) break;
      if (tau > 1 && can_direct && (tn + (double)nA + (double)nB) * esize > (double)g_ring_table_max) break;
      const double nheads = dbinom(hp, S.Rt - tau, S.gt - tau);
      const double avg_block = (double)S.seg / (nheads > 0 ? nheads : 1);
      double cost;
      if (tau == 1) {
        cost = 0.1 + block_cost / avg_block + (S.nE ? (400.0 + 0.5 * S.Rt) / (double)S.seg : 0.0);
      } else {
        // level-by-level build: ~0.3 warp-instructions per entry plus the barriers, once per class (single-run
        // classes) or per segment and CTA (mode B: a chunk sees about two segments)
        const double amort = (S.nE == 0) ? std::max(1.0, (double)C.size / 148.0) : 0.5 * (double)S.seg;
        cost = 0.1 + block_cost / avg_block + (0.3 * (tn + (double)nA + (double)nB) + 3000.0) / amort;
      }
      if (cost < best) { best = cost; best_tau = tau; }
    }
    if (can_direct) {
      const double nheads = dbinom(hp, S.Rt - 2, S.gt - 2);
      const double avg_block = (double)S.seg / (nheads > 0 ? nheads : 1);
      const double cost = 0.4 + block_cost / avg_block + (S.nE ? (400.0 + 0.5 * S.Rt) / (double)S.seg : 0.0);
      if (cost < best) { best = cost; best_tau = 2; best_direct = true; }
    }
    if (g_force_tau > 0) { best_tau = std::min(g_force_tau, (int)S.gt); best_direct = false; }
    if (g_force_tau == 2 && g_ring_direct == 2 && can_direct) best_direct = true;  // test hook: force the direct walk
    int64_t nA = 0, nB = 0;
    table_scratch(hp->h_binom, rank, S.Rt, best_tau, &nA, &nB);
    if (best_direct) nA = nB = 0;
    if (!best_direct && best_tau > 1 && dbinom(hp, S.Rt, best_tau) + (double)nA + (double)nB > (double)cap) { best_tau = 1; nA = nB = 0; }
    S.tau = best_tau;
    S.direct = best_direct ? 1 : 0;
    S.k0 = 0;
    int64_t direct_tbl = 0;
    if (best_direct) {
      // suffix table: as many rows as the table budget (and the shared memory left) hold; classes with earlier
      // runs have per-warp xr tables and no pair table at all
      S.k0 = S.Rt;
      if (S.nE == 0) {
        const int64_t budget = std::min<int64_t>(g_ring_table_max / esize, cap);
        int64_t m = 1;
        while (m < S.Rt && (m + 1) * m / 2 <= budget) ++m;   // C(m, 2) entries for the last m values
        if (m >= 2) { S.k0 = (int32_t)std::max<int64_t>(0, S.Rt - m); direct_tbl = (int64_t)(S.Rt - S.k0) * (S.Rt - S.k0 - 1) / 2; }
      }
    }
    S.hn = S.gt - S.tau;
    S.tbl_n = hp->h_binom[(int64_t)S.Rt * (rank + 1) + S.tau];
    if (S.tau > 1 && !S.direct) tbl_max = std::max(tbl_max, S.tbl_n + nA + nB);
    if (S.direct) tbl_max = std::max(tbl_max, direct_tbl);
  }
  *tbl_cap = (int32_t)((tbl_max + 31) / 32 * 32);
  {
    size_t off = ((size_t)*tbl_cap * esize + 15) / 16 * 16;
    off = (off + 2 * (size_t)dim * esize + 15) / 16 * 16;
    off = (off + (size_t)dim * 4 + 15) / 16 * 16;
    off += (size_t)*binom_smem * sizeof(int64_t);
    off += (size_t)hp->ncls * kClsInfoBytes;
    off += *cdesc_smem ? (size_t)hp->ncls * sizeof(ClassDesc) : 0;
    off = (off + 15) / 16 * 16;
    off += (size_t)nwarps * sizeof(WarpScratch);
    *smem_bytes = off + sizeof(TailCtrl);
  }
  return true;
}

static const int64_t kMaxDirTiles = (int64_t)1 << 23;  // 256 MB of directory at most

// ring kernel: strategy + ring geometry + tile size.  The slot size is the preferred one unless a class would lose
// its best tail length to the rings' shared memory: then smaller slots are tried (small copies cost bandwidth,
// a shorter tail costs much more).
int sm_count();
static bool compute_ring_strategy(const HostPlan* hp, int esize, std::vector<TailStrategy>& st, StratEntry& e, int64_t nsub_cap = 0) {
  const int NW = g_ring_warps, R = g_ring_slots;
  // shared memory kept away from the tables: the rings at their preferred slot size, their control structures and --
  // only when some class needs them -- the per-warp relabelled copies of x
  const size_t ctrl = (size_t)NW * R * 8 + (size_t)NW * (R + 2) * sizeof(TileQ) + (size_t)hp->ncls * sizeof(ClsRun) + 512;
  const size_t priv_bytes = (size_t)NW * hp->dim * esize + 16;
  auto needs_priv = [&](const std::vector<TailStrategy>& s) {
    for (int c = 0; c < hp->ncls; ++c) {
      if (s[c].tau == 0) continue;
      if (s[c].tau == 1) return true;                                  // tau == 1 always walks per-warp tables
      if (s[c].nE != 0 && s[c].direct && s[c].mu != 1) return true;  // (mu == 1: relabelled on the fly)
    }
    return false;
  };
  int32_t b0 = 0, c0 = 0;
  size_t sm0 = 0;
  bool ok = false;
  for (int bytes : {g_ring_bytes, 1536, 1024, 512}) {
    if (bytes > g_ring_bytes) continue;
    const size_t reserve = (size_t)NW * R * bytes + ctrl;
    ok = compute_tail_strategy(hp, esize, NW, st, &e.tbl_cap, &b0, &c0, &sm0, reserve, 40.0);
    if (ok && needs_priv(st)) ok = compute_tail_strategy(hp, esize, NW, st, &e.tbl_cap, &b0, &c0, &sm0, reserve + priv_bytes, 40.0);
    if (ok) break;
  }
  if (!ok) return false;
  e.nwarps = NW;
  e.ring_slots = R;
  e.binom_smem = 0;
  e.cdesc_smem = 0;
  e.priv_cap = needs_priv(st) ? (int32_t)(NW * hp->dim) : 0;
  const size_t base = ring_layout(esize, hp->dim, e.tbl_cap, e.priv_cap, 0, hp->ncls, 0, NW, R, 0).total;
  if (base + (size_t)NW * R * 512 > kRingSmemBudget) return false;
  size_t left = kRingSmemBudget - base;
  auto slot_bytes = [&](size_t avail, int cap_bytes) { return (int)std::min<size_t>((size_t)cap_bytes, avail / ((size_t)NW * R) / 512 * 512); };
  int bytes = slot_bytes(left, g_ring_bytes);
  left -= (size_t)NW * R * bytes;
  const size_t cdesc_bytes = (size_t)hp->ncls * sizeof(ClassDesc) + 32;
  if (cdesc_bytes <= left && cdesc_bytes <= 24 * 1024) { e.cdesc_smem = 1; left -= cdesc_bytes; }
  const size_t binom_bytes = (size_t)hp->binom_rows * (hp->rank + 1) * sizeof(int64_t);
  if (binom_bytes <= left && binom_bytes <= (size_t)kBinomSmemMax) { e.binom_smem = (int32_t)(hp->binom_rows * (hp->rank + 1)); left -= binom_bytes; }
  if (g_ring_bytes_max > bytes) bytes += slot_bytes(left, g_ring_bytes_max - bytes);
  e.ring_elems = bytes / esize;
  int64_t nsub = std::max(1, g_ring_tile_bytes / bytes);
  if (g_ring_small_tiles) {
    // small tensors: a tile per warp of the grid before tiles of the full size (rank 4 dim 50, 2.3 MB, ran on 3 CTAs --
    // 38 tiles of 48 KB, the costly last one alone 63 us -- and took 111 us; with one-slot tiles it spreads over 36 CTAs)
    int64_t total = 0;
    for (int c = 0; c < hp->ncls; ++c) total += hp->h_cls[c].size;
    const int64_t W = (int64_t)sm_count() * NW;
    const int64_t per_warp = (total * esize + W * bytes - 1) / (W * bytes);  // slots per warp, rounded up
    nsub = std::min<int64_t>(nsub, std::max<int64_t>(1, per_warp));
  }
  // short launches over a slice of the tensor (strong scaling: 1/8 of BASELINE config 2 is 69 MB, ~10 us of HBM time): the
  // launch lasts as long as its costliest tile, and one 48 KB tile of a class walked with gaps takes a warp 55 us -- small tiles
  if (nsub_cap > 0) nsub = std::min<int64_t>(nsub, nsub_cap);
  {
    // one workspace slot per tile of the whole tensor: grow the tiles until they fit
    auto total_tiles = [&](int64_t te) { int64_t n = 0; for (int c = 0; c < hp->ncls; ++c) n += (hp->h_cls[c].size + te - 1) / te; return n; };
    while (4 * total_tiles(nsub * e.ring_elems) > kWsSlots - kTilePartOff) nsub *= 2;
  }
  e.tile_elems = nsub * e.ring_elems;
  e.smem_bytes = ring_layout(esize, hp->dim, e.tbl_cap, e.priv_cap, e.binom_smem, hp->ncls, e.cdesc_smem, NW, R, e.ring_elems).total;
  return e.smem_bytes <= kRingSmemBudget;
}

// strategy of vec_ring_kernel for (device, rank, dim, element size): tail lengths, ring geometry, tile size, directories
static int get_strategy(int rank, int64_t dim, int esize, StratEntry* out, int64_t nsub_cap = 0) {
  const bool ring = true;
  int64_t tile = 0;
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return ST_ERR_INVALID;
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_smu);
  StratKey key{dev, rank, esize, 1, dim, nsub_cap};
  auto it = g_strats.find(key);
  if (it != g_strats.end()) { *out = it->second; return ST_OK; }
  StratEntry e;
  e.d_strat = nullptr;
  e.d_dir = nullptr;
  e.d_tile_base = nullptr;
  e.d_sdir = nullptr;
  e.d_sbase = nullptr;
  e.cdesc_smem = 0;
  e.tbl_cap = 32;
  e.binom_smem = 0;
  e.smem_bytes = 0;
  e.nwarps = g_ring_warps;
  e.ring_slots = 0;
  e.ring_elems = 0;
  e.priv_cap = 0;
  e.tile_elems = tile;
  std::vector<TailStrategy> st;
  if (ring) {
    e.supported = compute_ring_strategy(hp, esize, st, e, nsub_cap);
    tile = e.tile_elems;
    if (getenv("ST_VEC_DEBUG")) {
      fprintf(stderr, "[st] ring strategy rank %d dim %lld esize %d: supported %d warps %d slots %d slot_bytes %d tile_elems %lld tbl_cap %d binom_smem %d cdesc_smem %d smem %zu tau:",
              rank, (long long)dim, esize, (int)e.supported, e.nwarps, e.ring_slots, e.ring_elems * esize, (long long)e.tile_elems, e.tbl_cap, e.binom_smem,
              e.cdesc_smem, e.smem_bytes);
      for (size_t c = 0; c < st.size(); ++c) { fprintf(stderr, " %d", st[c].tau); if (st[c].direct) fprintf(stderr, "d(k0=%d)", st[c].k0); }
      fprintf(stderr, "\n");
    }
  }
  if (e.supported) {
    rc = check_cuda(cudaMalloc(&e.d_strat, sizeof(TailStrategy) * hp->ncls), "cudaMalloc(strategy)");
    if (rc) return rc;
    rc = check_cuda(cudaMemcpy(e.d_strat, st.data(), sizeof(TailStrategy) * hp->ncls, cudaMemcpyHostToDevice), "cudaMemcpy(strategy)");
    if (rc) return rc;
    // tile directory: built once per (device, rank, dim, tile size) by the GPU enumerator
    std::vector<int64_t> tb(hp->ncls + 1, 0);
    for (int c = 0; c < hp->ncls; ++c) tb[c + 1] = tb[c] + (hp->h_cls[c].size + tile - 1) / tile;
    const int64_t ntiles = tb[hp->ncls];
    e.h_strat = st;
    e.h_tile_base = tb;
    e.h_ntail.assign(hp->ncls, 0);
    if (ring) {
      // a tile that starts where the blocks are shorter than 512 components holds many pieces and costs several times
      // the usual: such tiles sit at the end of a single-run class (the head values only grow)
      const PlanView hv = hp->host_view();
      for (int c = 0; c < hp->ncls; ++c) {
        const TailStrategy& S = st[c];
        if (S.tau == 0 || S.nE != 0 || S.hn == 0) continue;
        const int64_t nt = tb[c + 1] - tb[c];
        int64_t n = 0;
        while (n < nt && n < 1024 && n < nt / 4) {
          int32_t u[ST_MAX_RANK];
          comb_unrank(hv.binom, hv.rank, (nt - 1 - n) * tile, S.Rt, S.gt, u);
          const int64_t bl = binom_at(hv.binom, hv.rank, S.Rt - 1 - u[S.hn - 1], S.tau);
          if (bl >= 512) break;
          ++n;
        }
        e.h_ntail[c] = n;
      }
    }
    if (g_use_dir && dim <= 65535 && ntiles > 0 && ntiles <= kMaxDirTiles) {
      PlanView P;
      rc = get_device_plan(rank, dim, &P);
      if (rc) return rc;
      rc = check_cuda(cudaMalloc(&e.d_tile_base, sizeof(int64_t) * (hp->ncls + 1)), "cudaMalloc(tile_base)");
      if (rc) return rc;
      rc = check_cuda(cudaMemcpy(e.d_tile_base, tb.data(), sizeof(int64_t) * (hp->ncls + 1), cudaMemcpyHostToDevice), "cudaMemcpy(tile_base)");
      if (rc) return rc;
      rc = check_cuda(cudaMalloc(&e.d_dir, sizeof(DirEntry) * ntiles), "cudaMalloc(tile directory)");
      if (rc) return rc;
      vec_dir_kernel<<<(unsigned)((ntiles + 127) / 128), 128>>>(P, tile, e.d_tile_base, ntiles, e.d_dir);
      count_launch();
      // small classes: one entry per component ("tile" of one component)
      std::vector<int64_t> sb(hp->ncls + 1, 0);
      for (int c = 0; c < hp->ncls; ++c) sb[c + 1] = sb[c] + (st[c].tau == 0 ? hp->h_cls[c].size : 0);
      const int64_t nsmall = sb[hp->ncls];
      e.h_sbase = sb;
      if (nsmall > 0) {
        rc = check_cuda(cudaMalloc(&e.d_sbase, sizeof(int64_t) * (hp->ncls + 1)), "cudaMalloc(sbase)");
        if (rc) return rc;
        rc = check_cuda(cudaMemcpy(e.d_sbase, sb.data(), sizeof(int64_t) * (hp->ncls + 1), cudaMemcpyHostToDevice), "cudaMemcpy(sbase)");
        if (rc) return rc;
        rc = check_cuda(cudaMalloc(&e.d_sdir, sizeof(DirEntry) * nsmall), "cudaMalloc(small-class directory)");
        if (rc) return rc;
        vec_dir_kernel<<<(unsigned)((nsmall + 127) / 128), 128>>>(P, 1, e.d_sbase, nsmall, e.d_sdir);
        count_launch();
      }
      rc = check_cuda(cudaDeviceSynchronize(), "vec_dir_kernel");
      if (rc) return rc;
    }
  }
  g_strats[key] = e;
  *out = e;
  return ST_OK;
}

int sm_count() {
  static std::mutex mu;
  static std::map<int, int> counts;  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  auto it = counts.find(dev);
  if (it != counts.end()) return it->second;
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  counts[dev] = n;
  return n;
}

// Two sets of {finished CTAs, -, one tile counter per class} per (device, stream), zeroed once and reset by the kernel
// itself; launches alternate between the sets (and between the halves of an ST_VEC_OVERLAP workspace), so that a launch
// overlapping the tail of the previous one never touches the previous one's counters or slots.
static std::mutex g_cmu;
struct CounterSet { unsigned long long* d; int parity; };
static std::map<std::pair<int, cudaStream_t>, CounterSet> g_counters;
static int get_counter(cudaStream_t stream, unsigned long long** out, int* parity_out) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_cmu);
  auto key = std::make_pair(dev, stream);
  auto it = g_counters.find(key);
  if (it == g_counters.end()) {
    unsigned long long* d = nullptr;
    rc = check_cuda(cudaMalloc(&d, 2 * kMaxCounters * sizeof(unsigned long long)), "cudaMalloc(counter)");
    if (rc) return rc;
    rc = check_cuda(cudaMemset(d, 0, 2 * kMaxCounters * sizeof(unsigned long long)), "cudaMemset(counter)");
    if (rc) return rc;
    it = g_counters.insert(std::make_pair(key, CounterSet{d, 0})).first;
  }
  it->second.parity ^= 1;
  *parity_out = it->second.parity;
  *out = it->second.d + (size_t)it->second.parity * kMaxCounters;
  return ST_OK;
}

// per-device: the dynamic shared memory opt-in of a kernel and the SM count (a process may drive several devices)
static std::mutex g_dmu;
static std::map<std::pair<int, const void*>, bool> g_attr_set;
int set_max_dynamic_smem(const void* func, int bytes) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_dmu);
  auto key = std::make_pair(dev, func);
  if (g_attr_set.count(key)) return ST_OK;
  rc = check_cuda(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute");
  if (rc) return rc;
  g_attr_set[key] = true;
  return ST_OK;
}

template <typename T>
static int launch_ring(VecArgs<T>& a, const StratEntry& se, const HostPlan* hp, int64_t len, int* grid_out, cudaStream_t stream, int flags) {
  {
    int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(vec_ring_kernel<T>), 227 * 1024);
    if (rc) return rc;
  }
  // One resident CTA per SM; tiles are dealt to the warps by the kernel (see RingSrc).
  const int nwarps = se.nwarps;
  const int64_t ntiles = (len + se.tile_elems - 1) / se.tile_elems;
  int64_t grid = std::min<int64_t>((int64_t)sm_count(), kMaxCtas);
  grid = std::min<int64_t>(grid, (kMaxPartials - kWarpPartOff) / ((nwarps + 3) & ~3));
  grid = std::max<int64_t>(1, std::min<int64_t>(grid, (ntiles + nwarps - 1) / nwarps));
  {
    int parity = 0;
    int rc = get_counter(stream, &a.counter, &parity);
    if (rc) return rc;
    // ST_VEC_OVERLAP: the workspace is two 4 MiB halves, one per parity
    if ((flags & ST_VEC_OVERLAP) && parity) {
      if (a.sum_out == a.partials) a.sum_out = a.partials + kWsSlots;
      a.partials += kWsSlots;
    }
  }
  a.pdl = (flags & ST_VEC_OVERLAP) ? 1 : 0;
  a.dynamic = (g_ring_dynamic && hp->ncls <= kMaxCounters - 2) ? 1 : 0;
  a.ondemand = g_ring_ondemand;
  RingSched sched;
  sched.n = 0;
  sched.pad_ = 0;
  if (hp->ncls <= kMaxSchedCls) {
    for (int c = 0; c < hp->ncls; ++c) {
      sched.cls[c].offset = hp->h_cls[c].offset;
      sched.cls[c].size = hp->h_cls[c].size;
      sched.cls[c].tile_base = se.h_tile_base.empty() ? 0 : se.h_tile_base[c];
      sched.cls[c].sbase = se.h_sbase.empty() ? 0 : se.h_sbase[c];
      sched.cls[c].S = se.h_strat[c];
    }
    // dynamic deal: the costly tail first when the launch covers the end of the class, then groups of tiles, single
    // tiles last (the static first deal follows the same order)
    make_runs(sched.cls, hp->ncls, a.begin, a.end, se.tile_elems, nwarps, (int)grid, sched.run, a.dynamic ? se.h_ntail.data() : nullptr,
              a.dynamic ? g_ring_group : 1, g_ring_fine, g_ring_fine_pos);
    sched.n = hp->ncls;
  }
  if (flags & ST_VEC_OVERLAP) {
    // programmatic dependent launch: the launch may begin once every CTA of the previous launch on the stream has let it
    // go (griddepcontrol.launch_dependents in vec_ring_kernel), i.e. while that launch drains its tail
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)(nwarps * 32));
    cfg.dynamicSmemBytes = se.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int rc = check_cuda(cudaLaunchKernelEx(&cfg, vec_ring_kernel<T>, a, sched), "cudaLaunchKernelEx(vec_ring_kernel)");
    if (rc) return rc;
  } else {
    vec_ring_kernel<T><<<(int)grid, nwarps * 32, se.smem_bytes, stream>>>(a, sched);
  }
  *grid_out = 1;  // the kernel leaves the launch's sum in partials[0]
  return ST_OK;
}

// Launch the main pass over [begin, end): writes `*grid_out` fp64 partials to `partials`.  With `d_out` the
// tail kernel also reduces them (`*fused` = true); otherwise (or on the generic path) the caller finalizes.
template <typename T>
static int vec_partials(int layout, int rank, int64_t dim, const T* d_packed, int64_t begin, int64_t end, const T* d_x,
                        double* partials, int* grid_out, T* d_out, bool* fused, cudaStream_t stream, double* ring_ws = nullptr, int flags = 0) {
  // `ring_ws`: separate workspace (kWsSlots) for vec_ring_kernel, whose only output is then partials[0]
  if (fused) *fused = false;
  if (layout != ST_LAYOUT_PERMCLS && layout != ST_LAYOUT_FLAT) { set_error("unknown layout %d", layout); return ST_ERR_INVALID; }
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  const int64_t total = layout == ST_LAYOUT_PERMCLS ? P.total : P.flat_size;
  if (begin < 0 || end < begin || end > total) { set_error("range [%lld, %lld) outside [0, %lld]", (long long)begin, (long long)end, (long long)total); return ST_ERR_INVALID; }
  if (begin % ST_CLASS_ALIGN) { set_error("begin must be a multiple of %d", ST_CLASS_ALIGN); return ST_ERR_INVALID; }
  if (!partials || (end > begin && (!d_packed || (dim > 0 && !d_x)))) { set_error("null pointer"); return ST_ERR_INVALID; }
  if (end > begin && ((uintptr_t)d_packed % 16)) { set_error("the packed buffer must be 16-byte aligned"); return ST_ERR_INVALID; }
  VecArgs<T> a;
  a.P = P;
  a.strat = nullptr;
  a.A = d_packed;
  a.x = d_x;
  a.begin = begin;
  a.end = end;
  a.partials = partials;
  a.out = nullptr;
  a.tile_elems = 0;
  a.dir = nullptr;
  a.tile_base = nullptr;
  a.sdir = nullptr;
  a.sbase = nullptr;
  a.cdesc_smem = 0;
  a.tbl_cap = 0;
  a.binom_smem = 0;
  a.counter = nullptr;
  a.ring_slots = 0;
  a.ring_elems = 0;
  a.priv_cap = 0;
  a.dynamic = 0;
  a.ondemand = 2;
  a.sum_out = partials;
  a.pdl = 0;
  a.tl = g_timeline;
  *grid_out = 1;
  if (end == begin) return check_cuda(cudaMemsetAsync(partials, 0, sizeof(double), stream), "cudaMemsetAsync");
  StratEntry se;
  se.supported = false;
  if (layout == ST_LAYOUT_PERMCLS && rank > 0 && g_variant != 1) {
    const bool short_launch = (end - begin) * (int64_t)sizeof(T) <= g_short_launch_bytes && (begin > 0 || end < total);
    rc = get_strategy(rank, dim, (int)sizeof(T), &se, short_launch ? g_short_launch_slots : 0);
    if (rc) return rc;
    if (!se.supported && g_variant == 2) { set_error("tail-table kernel unavailable for dim %lld", (long long)dim); return ST_ERR_UNSUPPORTED; }
  }
  if (se.supported) {
    a.strat = se.d_strat;
    a.tbl_cap = se.tbl_cap;
    a.binom_smem = se.binom_smem;
    a.dir = se.d_dir;
    a.tile_base = se.d_tile_base;
    a.sdir = se.d_sdir;
    a.sbase = se.d_sbase;
    a.cdesc_smem = se.cdesc_smem;
    a.out = d_out;
    a.ring_slots = se.ring_slots;
    a.ring_elems = se.ring_elems;
    a.priv_cap = se.priv_cap;
    a.tile_elems = se.tile_elems;
    if (ring_ws) a.partials = ring_ws;
    rc = launch_ring<T>(a, se, get_host_plan(rank, dim), end - begin, grid_out, stream, flags);
    if (rc) return rc;
    count_launch();
    if (fused) *fused = d_out != nullptr;
    return check_cuda(cudaGetLastError(), "vec_ring_kernel");
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
                        T* d_out, void* d_ws, cudaStream_t stream, int flags = 0) {
  if (!d_out || !d_ws) { set_error("null pointer"); return ST_ERR_INVALID; }
  if (flags & ~ST_VEC_OVERLAP) { set_error("unknown flags 0x%x", flags); return ST_ERR_INVALID; }
  double* partials = reinterpret_cast<double*>(d_ws);
  int grid = 1;
  bool fused = false;
  int rc = vec_partials<T>(layout, rank, dim, d_packed, begin, end, d_x, partials, &grid, d_out, &fused, stream, nullptr, flags);
  if (rc) return rc;
  if (fused) return ST_OK;
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
  int rc = get_stage((size_t)dim * sizeof(T), (size_t)kWsSlots + (size_t)n_chunks * kMaxPartials, &st);
  if (rc) return rc;
  double* sums = st->d_ws + kWsSlots;  // per chunk: kMaxPartials slots (the ring kernel uses the first one only)
  rc = check_cuda(cudaMemsetAsync(sums, 0, (size_t)n_chunks * kMaxPartials * sizeof(double), st->s_comp), "cudaMemsetAsync(ws)");
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
                         sums + c * kMaxPartials, &grid, nullptr, nullptr, st->s_comp, st->d_ws);
    if (rc) return rc;
    rc = check_cuda(cudaEventRecord(st->computed[b], st->s_comp), "cudaEventRecord");
    if (rc) return rc;
  }
  rc = vec_finalize<T>(sums, (int)(n_chunks * kMaxPartials), reinterpret_cast<T*>(st->d_out), st->s_comp);
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

int64_t st_contract_vec_workspace_bytes(void) { return (int64_t)sizeof(double) * kWsSlots; }

int st_set_tuning(const char* key, int64_t value) {
  if (!key) { set_error("null key"); return ST_ERR_INVALID; }
  const std::string k(key);
  if (k == "mat_dmma" && (value == 0 || value == 1)) { g_mat_dmma = (int)value; return ST_OK; }
  if (k == "mat_pipe" && (value == 0 || value == 1)) { g_mat_pipe = (int)value; return ST_OK; }
  if (k == "mat_onfly_rows" && value >= 0) { g_mat_onfly_rows = value; return ST_OK; }
  if (k == "conv_rows" && (value == 0 || value == 1)) { g_conv_rows = (int)value; return ST_OK; }
  if (k == "outer_fast" && (value == 0 || value == 1)) { g_outer_fast = (int)value; return ST_OK; }
  if (k == "outer_rows" && (value == 0 || value == 1)) { g_outer_rows = (int)value; return ST_OK; }
  if (k == "gram_umma" && (value == 0 || value == 1)) { g_gram_umma = (int)value; return ST_OK; }
  if (k == "sym22" && (value == 0 || value == 1)) { g_sym22 = (int)value; return ST_OK; }
  if (k == "sym22_min_dim" && value >= 1) { g_sym22_min_dim = value; return ST_OK; }
  if (k == "sym22_kch" && (value == 16 || value == 32)) { s22::g_kch = (int)value; return ST_OK; }
  if (k == "sym22_batch_tiles" && value >= 1) { s22::g_batch_tiles = value; return ST_OK; }
  if (k == "sym22_rgroup" && value >= 1 && value <= 64) { s22::g_tile_rgroup = (int)value; s22::clear_tile_cache(); return ST_OK; }
  if (k == "sym22_debug" && value >= 0 && value < 128) { s22::g_debug = (int)value; return ST_OK; }
  if (k == "vec_short_segment" && value >= 0) {
    g_short_segment = value;
    std::lock_guard<std::mutex> lk(g_smu);
    g_strats.clear();
    return ST_OK;
  }
  if (k == "vec_small_class" && value >= 0) {
    g_small_class = value;
    std::lock_guard<std::mutex> lk(g_smu);
    g_strats.clear();
    return ST_OK;
  }
  if (k == "vec_use_dir" && (value == 0 || value == 1)) {
    g_use_dir = (int)value;
    std::lock_guard<std::mutex> lk(g_smu);
    g_strats.clear();  // rebuilt on next use (old device tables are leaked: test hook)
    return ST_OK;
  }
  {
    if (k == "vec_ring_table_max" && value >= 0) {
      g_ring_table_max = value;
      std::lock_guard<std::mutex> lk(g_smu);
      g_strats.clear();
      return ST_OK;
    }
    if (k == "vec_timeline" && (value == 0 || value == 1)) {
      if (value && !g_timeline) {
        int rc = check_cuda(cudaMalloc(&g_timeline, kTimelineSlots * sizeof(unsigned long long)), "cudaMalloc(timeline)");
        if (rc) return rc;
      }
      if (!value && g_timeline) { cudaFree(g_timeline); g_timeline = nullptr; }
      if (g_timeline) cudaMemset(g_timeline, 0, kTimelineSlots * sizeof(unsigned long long));
      return ST_OK;
    }
    if (k == "vec_ring_dynamic" && (value == 0 || value == 1)) { g_ring_dynamic = (int)value; return ST_OK; }
    if (k == "vec_ring_direct" && value >= 0 && value <= 2) {
      g_ring_direct = (int)value;
      std::lock_guard<std::mutex> lk(g_smu);
      g_strats.clear();
      return ST_OK;
    }
    if (k == "vec_ring_group" && value >= 1 && value <= 64) { g_ring_group = (int)value; return ST_OK; }
    if (k == "vec_ring_fine" && value >= 0 && value <= 64) { g_ring_fine = (int)value; return ST_OK; }
    if (k == "vec_short_launch_bytes" && value >= 0) { g_short_launch_bytes = value; return ST_OK; }
    if (k == "vec_short_launch_slots" && value >= 1 && value <= 64) { g_short_launch_slots = value; return ST_OK; }
    if (k == "vec_ring_small_tiles" && (value == 0 || value == 1)) {
      g_ring_small_tiles = (int)value;
      std::lock_guard<std::mutex> lk(g_smu);
      g_strats.clear();
      return ST_OK;
    }
    if (k == "vec_ring_fine_pos" && value >= 0 && value <= 100) { g_ring_fine_pos = (int)value; return ST_OK; }
    if (k == "vec_ring_ondemand" && value >= 0 && value <= 64) { g_ring_ondemand = (int)value; return ST_OK; }
    int* knob = k == "vec_ring_warps" ? &g_ring_warps : k == "vec_ring_slots" ? &g_ring_slots : k == "vec_ring_bytes" ? &g_ring_bytes :
                k == "vec_ring_bytes_max" ? &g_ring_bytes_max : k == "vec_ring_tile_bytes" ? &g_ring_tile_bytes : nullptr;
    if (knob) {
      const bool ok = (knob == &g_ring_warps) ? (value >= 1 && value <= 16) : (knob == &g_ring_slots) ? (value >= 2 && value <= 8) :
                      (knob == &g_ring_tile_bytes) ? (value >= 512 && value <= (1 << 24)) : (value >= 512 && value <= 65536 && value % 512 == 0);
      if (!ok) { set_error("value %lld out of range for '%s'", (long long)value, key); return ST_ERR_INVALID; }
      *knob = (int)value;
      std::lock_guard<std::mutex> lk(g_smu);
      g_strats.clear();  // rebuilt on next use (old device tables are leaked: tuning hook)
      return ST_OK;
    }
  }
  if (k == "vec_force_tau" && value >= 0 && value <= ST_MAX_RANK) {
    g_force_tau = (int)value;
    std::lock_guard<std::mutex> lk(g_smu);
    g_strats.clear();  // strategies are rebuilt on next use (device tables of old entries are leaked: test hook)
    return ST_OK;
  }
  set_error("unknown tuning key '%s' or value %lld out of range", key, (long long)value);
  return ST_ERR_INVALID;
}

int st_debug_vec_timeline(unsigned long long* h_out, int64_t n) {
  if (!g_timeline || !h_out || n < 0 || (size_t)n > kTimelineSlots) { set_error("timeline off or bad size"); return ST_ERR_INVALID; }
  int rc = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  if (rc) return rc;
  return check_cuda(cudaMemcpy(h_out, g_timeline, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost), "cudaMemcpy(timeline)");
}

int st_set_vec_variant(int variant) {
  if (variant < 0 || variant > 2) { set_error("variant must be 0, 1 or 2"); return ST_ERR_INVALID; }
  g_variant = variant;
  std::lock_guard<std::mutex> lk(g_smu);
  g_strats.clear();  // strategies depend on the variant (old device tables are leaked: test hook)
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

int st_contract_vec_ex_f64(int layout, int rank, int64_t dim, const double* d_packed, int64_t begin, int64_t end,
                           const double* d_x, double* d_out, void* d_workspace, int flags, void* stream) {
  return contract_vec<double>(layout, rank, dim, d_packed, begin, end, d_x, d_out, d_workspace, (cudaStream_t)stream, flags);
}

int st_contract_vec_ex_f32(int layout, int rank, int64_t dim, const float* d_packed, int64_t begin, int64_t end,
                           const float* d_x, float* d_out, void* d_workspace, int flags, void* stream) {
  return contract_vec<float>(layout, rank, dim, d_packed, begin, end, d_x, d_out, d_workspace, (cudaStream_t)stream, flags);
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
