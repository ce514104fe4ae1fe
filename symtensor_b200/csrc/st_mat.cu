// contract_all_indices_with_matrix, fp64, dims up to 64: one step of the mode chain as a PERSISTENT producer / consumer
// pipeline (round 2).  Reference semantics: symtensor/symalg.py:476-496 (the chain itself: st_ops.cu, contract_mat).
//
//   T_{k+1}[(J, j); I] = sum_a W[a, j] T_k[J; sort(a, I)],   j >= max(J)
//
// A work item is (J, tile of 64 consecutive I): the 64 x 64 operand S[a][i] = T_k[J][map[a][i]] is gathered from the packed
// row of J through the step's gather map, multiplied by the columns j >= max(J) of W on the FP64 tensor pipe
// (mma.sync.m8n8k4.f64, SASS DMMA; tcgen05 has no f64 kind) and written to the rows (J, j) of T_{k+1}, which are consecutive.
// What the first kernel (mat_step_dmma_kernel) lost, measured per step with ncu at BASELINE config 4 (rank 6 dim 64):
//  * the late steps are millions of small items and every CTA began with a serial unrank of J by one thread, a W tile from
//    L2 and a cold pipeline: 28.6 ms for step 4, whose traffic is 18 GB (3 ms of HBM time);
//  * gathers were staged through registers with two block-wide barriers per 32-deep chunk;
//  * whole 32-column blocks were multiplied although only the columns j >= max(J) are stored (1.5 - 2.5x the DMMAs).
// Here: CTAs are persistent (two per SM) and claim chunks of consecutive items from a global counter; W sits in shared
// memory for the whole kernel; four PRODUCER warps walk the item sequence, read the gather map (a tile ahead, into registers)
// and issue 8-byte cp.async gathers straight into a two-stage ring of operand tiles (completion through
// cp.async.mbarrier.arrive on the stage's `full` barrier; invalid entries are zero-filled by the same instruction with
// src-size 0), one producer lane keeps J by odometer steps and publishes the item descriptor; eight CONSUMER warps wait on
// `full`, multiply only the 8-column blocks that hold a column j >= max(J) (dealt alternately to the two column groups of
// warps; a branch per block count -- a predicated-off DMMA still takes its 16 cycles of the pipe), store, and release the
// stage on `empty`.  The shape is a template of the constants below: one CTA per SM with a five-stage ring, sixteen consumer
// and eight producer warps (NGROUP 2, NPROD 256, STAGES 5) measured SLOWER (105 ms against 83 ms for the whole chain).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <map>
#include <mutex>

#include "st_common.cuh"

namespace st {
namespace matpipe {

constexpr int TI = 64;   // rows I per tile
constexpr int TA = 64;   // contraction length held per tile (dim <= 64)
constexpr int LD = 68;   // row stride in doubles: 4 mod 16, conflict-free fragment loads
constexpr int NGROUP = 1;                      // consumer groups of 8 warps: group g multiplies the items n = g (mod NGROUP) of the CTA
constexpr int NCONS = 256 * NGROUP, NPROD = 128, NTHREADS = NCONS + NPROD;
constexpr int STAGES = 2;                      // operand tiles in the ring (two CTAs per SM: W + 2 tiles = 104 KB each)
constexpr int CTAS_PER_SM = 2;
constexpr int CH = 8;                          // items per claim
constexpr int kSmemBytes = ((1 + STAGES) * TA * LD) * 8 + 512 + ST_MAX_RANK * TA * 4;  // + the rank terms of the on-the-fly gather map

struct Desc {
  long long i0;        // first row I of the tile
  long long out_base;  // element offset of the output row (J, jlast) in T_{k+1}
  int jlast;           // first stored column
  int valid;
};

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// the arrival fires when every cp.async this thread has issued so far has landed (counts as one of the expected arrivals)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// bounded wait: a barrier that never completes must end the kernel with an error, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < (1 << 26); ++it) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return;
    if (it > 4) __nanosleep(it > 64 ? 256 : 32);  // (polls take issue slots from the warps that have work: 40 % of the kernel's instructions without this)
  }
  __trap();
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// one warp's 16 rows x NB 8-column blocks of a tile: NB x 2 accumulators over the whole contraction, then the store.
// Accumulator (i, jj) holds rows i * 8 + lane / 4, columns col0 + 16 jj + {0, 1} of the warp's blocks.
template <int NB>
__device__ __forceinline__ void tile_mma(const double* __restrict__ ap, const double* __restrict__ bp, int kend, double* __restrict__ obase, long long nI,
                                         int col0, int jlast, int chi, int rows_left) {
  double acc[2][NB][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < kend; kk += 4) {
    const double a0 = ap[kk * LD], a1 = ap[kk * LD + 8];
#pragma unroll
    for (int jj = 0; jj < NB; ++jj) {
      const double b = bp[kk * LD + jj * 16];
      dmma(acc[0][jj], a0, b);
      dmma(acc[1][jj], a1, b);
    }
  }
#pragma unroll
  for (int jj = 0; jj < NB; ++jj)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = col0 + 16 * jj + h;
      if (col < jlast || col >= chi) continue;
      double* orow = obase + (long long)(16 * jj + h) * nI;
      if (rows_left > 0) orow[0] = acc[0][jj][h];
      if (rows_left > 8) orow[8] = acc[1][jj][h];
    }
}

// ONFLY: no gather map in memory -- the producers rank sort(a, I) themselves (step 0, whose map would be used exactly once:
// 2.8 GB written and read back at rank 6 dim 64): one unrank of the thread's I per tile, then
// rank = base - PS[p] - F[m-p][a] with p = #{q: I[q] <= a} growing along the thread's ascending a (see mat_index_kernel).
template <bool ONFLY>
__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM) mat_pipe_kernel(PlanView P, int k, int m, const double* __restrict__ Tk, const double* __restrict__ W,
                                                               double* __restrict__ Tn, long long nJ, long long nI, long long nI1,
                                                               const int32_t* __restrict__ tbl, long long rlo, int clo, int chi,
                                                               unsigned long long* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ws = reinterpret_cast<double*>(smem_raw);            // [TA][LD]   W[a][j]
  double* Ss = Ws + TA * LD;                                     // [STAGES][TA][LD] S[a][i]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(Ss + STAGES * TA * LD);  // full[STAGES], empty[STAGES]
  Desc* desc = reinterpret_cast<Desc*>(bars + 2 * STAGES);       // [STAGES]
  long long* s_claim = reinterpret_cast<long long*>(desc + STAGES);  // [2]
  int32_t* Fm = reinterpret_cast<int32_t*>(smem_raw + ((1 + STAGES) * TA * LD) * 8 + 512);  // [m + 1][TA]: F[t][v] = C(d-1+t-v, t+1)
  const int d = (int)P.dim;
  const int tid = threadIdx.x;
  for (int e = tid; e < TA * LD; e += NTHREADS) {
    const int a = e / LD, j = e % LD;
    Ws[e] = (a < d && j < d) ? W[a * d + j] : 0.0;
  }
  if (ONFLY) {
    for (int e = tid; e < (m + 1) * TA; e += NTHREADS) {
      const int t = e / TA, v = e % TA;
      Fm[e] = v < d ? (int32_t)binom_at(P.binom, P.rank, d - 1 + t - v, t + 1) : 0;
    }
  }
  if (tid == 0) {
    for (int b = 0; b < STAGES; ++b) {
      mbar_init(saddr(bars + b), NPROD + 1);   // full: every producer's cp.async arrival + the descriptor writer
      mbar_init(saddr(bars + STAGES + b), 8);  // empty: one arrival per warp of the group that multiplied the tile
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long tilesI = (nI + TI - 1) / TI;
  const long long total = nJ * tilesI;

  if (tid >= NCONS) {
    // ---------------- producers ----------------
    const int pt = tid - NCONS;
    int32_t J[ST_MAX_RANK];
    long long jr_cur = -2;
    long long w = 0, cend = 0;  // the first claim happens at the first item
    unsigned long long next_claim = 0;
    int slot = 0;
    if (pt == 0) next_claim = atomicAdd(counter, (unsigned long long)CH);
    constexpr int AQ = NPROD / 64, PER = TA / AQ;  // a thread gathers the rows a = AQ u + aq of its column
    const int r = pt & 63, aq = pt >> 6;
    // the next item of this CTA's sequence (all producers compute the same sequence)
    auto next_item = [&]() {
      if (w == cend) {  // next chunk of items (its claim was issued a chunk ago)
        if (pt == 0) s_claim[slot] = (long long)next_claim;
        asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
        w = s_claim[slot];
        cend = w + CH;
        slot ^= 1;
        if (pt == 0) next_claim = atomicAdd(counter, (unsigned long long)CH);
      }
    };
    // gather-map entries of an item's column: issued a whole tile ahead, so that their latency overlaps the wait for the stage
    int32_t idx[PER];
    const int32_t gbase = ONFLY ? (int32_t)(binom_at(P.binom, P.rank, d + m, m + 1) - 1) : 0;
    auto load_idx = [&](long long wi) {
      if (wi >= total) return;
      const long long jrel = wi / tilesI, i0 = (wi - jrel * tilesI) * TI;
      const bool rvalid = i0 + r < nI;
      if (ONFLY) {
        int32_t I[ST_MAX_RANK];
        int32_t ps = 0;
        if (rvalid) {
          flat_unrank_r(P, i0 + r, m, I);
          for (int q = 0; q < m; ++q) ps += Fm[(m - 1 - q) * TA + I[q]];
        }
        int p = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          const int a = AQ * u + aq;
          if (rvalid && a < d) {
            while (p < m && I[p] <= a) {
              ps += Fm[(m - p) * TA + I[p]] - Fm[(m - 1 - p) * TA + I[p]];
              ++p;
            }
            idx[u] = gbase - ps - Fm[(m - p) * TA + a];
          } else {
            idx[u] = -1;
          }
        }
        return;
      }
      const int32_t* __restrict__ tcol = tbl + i0 + r;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int a = AQ * u + aq;
        idx[u] = (rvalid && a < d) ? __ldg(tcol + (long long)a * nI) : -1;
      }
    };
    next_item();
    load_idx(w);
    int n_end = 0;
    for (long long n = 0;; ++n) {
      const int buf = (int)(n % STAGES);
      mbar_wait(saddr(bars + STAGES + buf), (uint32_t)(((n / STAGES) & 1) ^ 1));
      const uint32_t full = saddr(bars + buf);
      if (w >= total) {  // no more work: tell every consumer group (one empty item each) and leave
        if (pt == 0) {
          desc[buf].valid = 0;
          mbar_arrive(full);
        }
        cp_async_arrive(full);
        if (++n_end == NGROUP) break;
        continue;
      }
      const long long jrel = w / tilesI, t = w - jrel * tilesI;
      const long long jr = rlo + jrel, i0 = t * TI;
      const double* __restrict__ row = Tk + jr * nI1;
      double* Sb = Ss + buf * TA * LD;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int a = AQ * u + aq;
        cp_async_8(saddr(Sb + a * LD + r), idx[u] >= 0 ? (const void*)(row + idx[u]) : (const void*)Tk, idx[u] >= 0 ? 8u : 0u);
      }
      cp_async_arrive(full);
      if (pt == 0) {
        if (jr != jr_cur) {
          if (jr == jr_cur + 1 && k > 0) {  // successor of a sorted k-tuple over range(d)
            int q = k - 1;
            while (q > 0 && J[q] == d - 1) --q;
            const int32_t v = J[q] + 1;
            for (int s = q; s < k; ++s) J[s] = v;
          } else if (k > 0) {
            flat_unrank_r(P, jr, k, J);
          }
          jr_cur = jr;
        }
        const int jl = max(k ? J[k - 1] : 0, clo);
        J[k] = jl;
        desc[buf].i0 = i0;
        desc[buf].out_base = flat_rank_r(P, J, k + 1) * nI;
        desc[buf].jlast = jl;
        desc[buf].valid = 1;
        mbar_arrive(full);  // (release: the descriptor is visible to whoever sees the phase complete)
      }
      ++w;
      next_item();
      load_idx(w);
    }
    return;
  }

  // ---------------- consumers ----------------
  const int lane = tid & 31, warp = (tid >> 5) & 7, group = tid >> 8;
  const int wr = warp & 3, wc = warp >> 2;  // 16-row block of the tile; column group
  const int kend = (d + 3) & ~3;
  const int jb1 = (chi + 7) >> 3;
  for (long long n = group;; n += NGROUP) {
    const int buf = (int)(n % STAGES);
    mbar_wait(saddr(bars + buf), (uint32_t)((n / STAGES) & 1));
    const Desc D = desc[buf];
    if (!D.valid) break;
    const double* Sb = Ss + buf * TA * LD;
    const int jb0 = D.jlast >> 3;
    if (D.i0 + wr * 16 < nI && jb0 + wc < jb1) {
      const int nblk = (jb1 - jb0 - wc + 1) >> 1;  // this warp's 8-column blocks: jb0 + wc + 2 jj, jj < nblk <= 4
      const double* ap = Sb + (lane & 3) * LD + wr * 16 + (lane >> 2);
      const double* bp = Ws + (lane & 3) * LD + (jb0 + wc) * 8 + (lane >> 2);
      double* obase = Tn + D.out_base + (long long)((jb0 + wc) * 8 + 2 * (lane & 3) - D.jlast) * nI + D.i0 + wr * 16 + (lane >> 2);
      const int col0 = (jb0 + wc) * 8 + 2 * (lane & 3);
      const int rows_left = (int)min((long long)16, nI - D.i0 - wr * 16) - (lane >> 2);  // rows i * 8 + lane / 4 below it are stored
      // (a branch per block count, not a predicate per DMMA: a predicated-off DMMA still takes its 16 cycles of the pipe)
      switch (nblk) {
        case 1: tile_mma<1>(ap, bp, kend, obase, nI, col0, D.jlast, chi, rows_left); break;
        case 2: tile_mma<2>(ap, bp, kend, obase, nI, col0, D.jlast, chi, rows_left); break;
        case 3: tile_mma<3>(ap, bp, kend, obase, nI, col0, D.jlast, chi, rows_left); break;
        default: tile_mma<4>(ap, bp, kend, obase, nI, col0, D.jlast, chi, rows_left); break;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(saddr(bars + STAGES + buf));
  }
}

// one 8-byte work counter per (device, stream), allocated once (never freed)
static unsigned long long* counter_for(cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, unsigned long long*> ctrs;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_pair(dev, stream);
  auto it = ctrs.find(key);
  if (it != ctrs.end()) return it->second;
  unsigned long long* p = nullptr;
  if (cudaMalloc(&p, 256) != cudaSuccess) return nullptr;
  ctrs[key] = p;
  return p;
}

// launches one step through the pipeline kernel (tbl == nullptr: the producers rank the gathers themselves); false (nothing
// enqueued): shape not covered (dim > 64, positions beyond int32)
bool launch_step(const PlanView& P, int k, int m, const double* Tk, const double* W, double* Tn, int64_t nJ, int64_t nI, int64_t nI1,
                 const int32_t* tbl, int64_t rlo, int clo, int chi, cudaStream_t stream) {
  if (P.dim > TA || nI < 1 || nJ < 1 || m < 1 || nI1 >= 2147483647LL) return false;
  const bool onfly = tbl == nullptr;
  unsigned long long* ctr = counter_for(stream);
  if (!ctr) return false;
  const void* fn = onfly ? reinterpret_cast<const void*>(mat_pipe_kernel<true>) : reinterpret_cast<const void*>(mat_pipe_kernel<false>);
  if (set_max_dynamic_smem(fn, kSmemBytes) != ST_OK) return false;
  cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), stream);
  const int64_t tilesI = (nI + TI - 1) / TI;
  const int64_t chunks = (nJ * tilesI + CH - 1) / CH;
  const int grid = (int)std::min<int64_t>(chunks, (int64_t)sm_count() * CTAS_PER_SM);
  if (onfly)
    mat_pipe_kernel<true><<<grid, NTHREADS, kSmemBytes, stream>>>(P, k, m, Tk, W, Tn, (long long)nJ, (long long)nI, (long long)nI1, tbl, (long long)rlo, clo, chi, ctr);
  else
    mat_pipe_kernel<false><<<grid, NTHREADS, kSmemBytes, stream>>>(P, k, m, Tk, W, Tn, (long long)nJ, (long long)nI, (long long)nI1, tbl, (long long)rlo, clo, chi, ctr);
  return true;
}

}  // namespace matpipe
}  // namespace st
