// Symmetrized tensordot with TWO free indices on each side (ra - k == 2 and rb - k == 2), non-materialising:
// BASELINE config 3 (rank 3 . rank 3 over one index, dim 1000, fp32 -> rank 4, 41.9 G components; symtensor/symalg.py:427-459).
//
//     C[i<=j<=k<=l] = 1/6 * ( G[ij|kl] + G[kl|ij] + G[ik|jl] + G[jl|ik] + G[il|jk] + G[jk|il] ),
//     G[ab|cd] = sum_J w_J A[a, b, J] B[J, c, d]          (J: the packed contracted tuples, w_J their multiplicity)
//
// The Gram matrix G (pair rows x pair columns, 1 TB at config 3) is never stored.  The output is cut into TILES of index
// blocks  I x J x K x L = 16 x 16 x 16 x 8  (L: block of the LARGEST index).  For one tile the six terms are THREE
// GEMMs of shape [128 x 256 x 2 Kc] -- the two orientations of a pairing share an accumulator through the concatenated
// contraction  G[ab|cd] + G[cd|ab] = sum_J [wA | B][ab, J] . [B | wA][cd, J]  -- with rows (., l) = 16 x 8 = 128 and
// columns 16 x 16 = 256 for every pairing, which is the full-rate tcgen05 shape (M = 128, N = 256, cta_group::1):
//     pairing 0: rows (k, l), columns (i, j);   pairing 1: rows (j, l), columns (i, k);   pairing 2: rows (i, l), columns (j, k).
// l sits in the ROW index of every pairing because an epilogue thread owns a row (a TMEM lane): the 32 lanes of a warp
// then hold 4 x 8 consecutive values of l, i.e. 4 whole 32-byte sectors of the packed output per global add (with l in the
// columns every lane hit its own sector and the adds alone took longer than the GEMMs).
// Operands: the expanded pair matrices X[K chunk][first][second][16 or 32 floats] (dim x dim rows, both triangles) of wA and B, each
// pre-split into the tf32 part the tensor core reads and the exact remainder (3xTF32: hi.hi + hi.lo + lo.hi), written
// once by expand_pairs_kernel; a tile's operand boxes (16 x 8 or 16 x 16 pair rows x 16 or 32 floats) are fetched by TMA
// (cp.async.bulk.tensor.4d, SASS UTMALDG; hardware swizzle, out-of-range rows zero-filled) into a ring of shared-memory
// stages guarded by full / empty mbarriers.
//
// Warp roles (384 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane; allocates the 512 TMEM
// columns = two 128 x 256 fp32 accumulators), warps 4..11 = epilogue.  The tensor core accumulates into TMEM with
// truncation, so a CHAIN is at most 256 products long (st_ops.cu: measured 1.4e-5 of sum|terms| at 1024, 3.5e-6 at 256):
// after every chain the issuer commits the accumulator to the epilogue (tmem-full barrier) and continues in the other
// accumulator; the epilogue warps drain it with tcgen05.ld and add the chains in REGISTERS with round-to-nearest (each
// thread keeps 128 columns of its row).  At the end of a pairing the thread scales by 1/6 and writes (first pairing) or adds
// (red.global.add.f32, the other two) its values into the tile's own contiguous 128 KB SLOT of a scratch buffer -- one
// page, L2-resident across the three pairings; the three contributions of a component come from the same CTA in a fixed
// order, so the result is deterministic.  A second, plain HBM-bound kernel (sym22_scatter_kernel) moves the slots into the
// packed permcls layout: the position of (i < j < k < l) is a sum of four per-index terms, components with repeated indices
// (diagonal tiles) go through the generic class rank.  Tiles are processed in batches of 32768 (4 GB of slots).
// [begin, end) ranges of the output are served by the tiles that intersect them -- the multi-GPU partition, no collective.
#include <cuda.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "st_common.cuh"

namespace st {
namespace s22 {

constexpr int BJ = 16, BL = 8;  // index blocks: i, j, k in blocks of 16, l (the LARGEST index) in blocks of 8
constexpr int TM = BJ * BL;   // 128 rows: (one of i, j, k) x l
constexpr int TN = BJ * BJ;   // 256 columns: the other two
constexpr int NTHREADS = 384;
constexpr int CHAIN_K = 256;  // products per TMEM accumulation chain
constexpr int SMEM_STAGE_BUDGET = 192 * 1024;
constexpr int64_t kBatchTiles = 32768;  // tiles per launch: their slots (4 GB) are the scratch part of the workspace
int64_t g_batch_tiles = kBatchTiles;    // (tests lower it -- key "sym22_batch_tiles" -- to run several batches at small sizes; the workspace is sized for kBatchTiles)

template <int KCH>
struct Geo {
  static constexpr int ROW_BYTES = TM * KCH * 4;
  static constexpr int COL_BYTES = TN * KCH * 4;
  static constexpr int STAGE_BYTES = 2 * ROW_BYTES + 2 * COL_BYTES;
  static constexpr int STAGES = SMEM_STAGE_BUDGET / STAGE_BYTES;
  static constexpr int CHAIN_STAGES = CHAIN_K / KCH;
  static constexpr uint32_t SBO = 8 * KCH * 4;                  // bytes between 8-row groups
  static constexpr uint64_t LAYOUT = KCH == 32 ? 2ull : 4ull;   // SWIZZLE_128B : SWIZZLE_64B
};

// the output ranges of one call (a GPU's shard: its part of every class with repeated indices + its part of class (1,1,1,1)),
// each with its own destination buffer; disjoint
constexpr int kMaxRanges = 8;
struct OutRanges {
  int32_t n;
  int64_t begin[kMaxRanges], end[kMaxRanges];
  float* out[kMaxRanges];
};

struct Params {
  PlanView P;          // plan of the rank-4 output
  int64_t begin, end;  // packed output coordinates served by this launch
  float* out;          // points at coordinate `begin` (sym22_scatter_kernel)
  float* scratch;      // one 128 KB slot per tile of the launch (tile t of the launch -> slot t)
  const unsigned long long* tiles;  // block numbers p | q << 16 | r << 32 | s << 48 of i, j, k (16 wide) and l (8 wide)
  int64_t ntiles;
  int32_t nst;         // stages per K segment (Kp / KCH)
  int32_t debug;       // ablation switches (tuning key "sym22_debug"): 1 no global adds, 2 no TMEM drains, 4 no TMA loads, 8 no MMAs
  int* err;            // set to 1 when a bounded barrier wait expired
};

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a barrier that never completes must end the kernel (and raise the error flag), not hang the GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return true;
    if ((it & 1023) == 1023 && *abort_flag) return false;
  }
  *abort_flag = 1;
  return false;
}
// one operand box: K chunk `c3` of the pair rows (first0 .., second0 ..) -- a 4-D tile of X[chunk][first][second][KCH]
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap* map, int second0, int first0, int chunk, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(second0), "r"(first0), "r"(chunk), "r"(bar), "l"(policy)
               : "memory");
}
// L2 policies: the operand stream must not push the output tiles out of L2 between the three pairings of a tile (a tile's
// 128 KB of outputs are written by pairing 0 and added to by pairings 1 and 2, ~100 us apart, while ~1 GB of operands
// streams through the 126 MB L2): operands evict-first (or evict-normal), outputs evict-last
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <int KCH>
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {  // K-major, hardware swizzle, 8-row groups SBO apart
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(Geo<KCH>::SBO >> 4) << 32) | ((uint64_t)1 << 46) | (Geo<KCH>::LAYOUT << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {  // D = F32, A = B = TF32, both K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// binomials C(n, 2..4) for n >= 0 by exact 64-bit arithmetic (every division is exact)
__device__ __forceinline__ int64_t c2(int64_t n) { return n * (n - 1) / 2; }
__device__ __forceinline__ int64_t c3(int64_t n) { return c2(n) * (n - 2) / 3; }
__device__ __forceinline__ int64_t c4(int64_t n) { return c3(n) * (n - 3) / 4; }

// packed permcls coordinate of the component (a <= b <= c <= e) of the rank-4 output, or -1
__device__ __forceinline__ int64_t element_coord(const PlanView& P, int64_t base1111, int a, int b, int c, int e) {
  if (!(a <= b && b <= c && c <= e) || e >= P.dim) return -1;
  if (a < b && b < c && c < e) {
    return base1111 - binom_at(P.binom, 4, P.dim - 1 - a, 4) - binom_at(P.binom, 4, P.dim - 1 - b, 3) - binom_at(P.binom, 4, P.dim - 1 - c, 2) -
           binom_at(P.binom, 4, P.dim - 1 - e, 1);
  }
  const int32_t idx[4] = {a, b, c, e};
  int32_t vals[ST_MAX_RANK];
  const int ci = classify_index(P, idx, vals);
  if (ci < 0) return -1;
  return P.cls[ci].offset + permcls_rank_vals(P, P.cls[ci], vals);
}

// Second pass: the tiles' slots (tile-major, L(i, j, k, l) = ((i 16 + j) 16 + k) 8 + l) -> the packed permcls output.  One CTA
// per tile; a thread handles consecutive L, so 8 lanes write the 8 consecutive l of one (i, j, k) = one 32-byte sector, and
// the many warps in flight hide the page walks of the scattered rows.  Components outside [begin, end), unsorted index
// combinations of diagonal tiles and indices beyond the tensor are skipped; every component of the range is written exactly
// once (by the one tile that holds it).  Strictly increasing indices take the four-term formula from shared-memory tables,
// repeated indices the generic class rank.
__global__ void __launch_bounds__(256) sym22_scatter_kernel(PlanView P, const unsigned long long* __restrict__ tiles, const float* __restrict__ scratch,
                                                            const OutRanges R) {
  __shared__ long long Ti[BJ], Tj[BJ], Tk[BJ], Tl[BL];
  const unsigned long long tw = tiles[blockIdx.x];
  const int i0 = (int)(tw & 0xffff) * BJ, j0 = (int)((tw >> 16) & 0xffff) * BJ, k0 = (int)((tw >> 32) & 0xffff) * BJ,
            l0 = (int)((tw >> 48) & 0xffff) * BL;
  if (threadIdx.x < BJ) {
    Ti[threadIdx.x] = binom_at(P.binom, 4, P.dim - 1 - (i0 + threadIdx.x), 4);
    Tj[threadIdx.x] = binom_at(P.binom, 4, P.dim - 1 - (j0 + threadIdx.x), 3);
    Tk[threadIdx.x] = binom_at(P.binom, 4, P.dim - 1 - (k0 + threadIdx.x), 2);
    if (threadIdx.x < BL) Tl[threadIdx.x] = binom_at(P.binom, 4, P.dim - 1 - (l0 + threadIdx.x), 1);
  }
  __syncthreads();
  const int64_t base1111 = P.cls[P.ncls - 1].offset + binom_at(P.binom, 4, P.dim, 4) - 1;
  const float* __restrict__ slot = scratch + (size_t)blockIdx.x * (TM * TN);
  for (int L = threadIdx.x; L < TM * TN; L += 256) {
    const int l = L & 7, k = (L >> 3) & 15, j = (L >> 7) & 15, i = L >> 11;
    const int gi = i0 + i, gj = j0 + j, gk = k0 + k, gl = l0 + l;
    if (!(gi <= gj && gj <= gk && gk <= gl) || gl >= P.dim) continue;
    int64_t coord;
    if (gi < gj && gj < gk && gk < gl) coord = base1111 - Ti[i] - Tj[j] - Tk[k] - Tl[l];
    else coord = element_coord(P, base1111, gi, gj, gk, gl);
#pragma unroll
    for (int q = 0; q < kMaxRanges; ++q)
      if (q < R.n && coord >= R.begin[q] && coord < R.end[q]) R.out[q][coord - R.begin[q]] = slot[L];
  }
}

template <int KCH>
__global__ void __launch_bounds__(NTHREADS, 1)
sym22_umma_kernel(const __grid_constant__ CUtensorMap mAhr, const __grid_constant__ CUtensorMap mAlr, const __grid_constant__ CUtensorMap mBhr,
                  const __grid_constant__ CUtensorMap mBlr, const __grid_constant__ CUtensorMap mAhc, const __grid_constant__ CUtensorMap mAlc,
                  const __grid_constant__ CUtensorMap mBhc, const __grid_constant__ CUtensorMap mBlc, const __grid_constant__ Params prm) {
  using G = Geo<KCH>;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)G::STAGES * G::STAGE_BYTES);
  // bars: full[STAGES], empty[STAGES], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G::STAGES + 4);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = saddr(bars), bar_empty = saddr(bars + G::STAGES), bar_tfull = saddr(bars + 2 * G::STAGES),
                 bar_tempty = saddr(bars + 2 * G::STAGES + 2);
  if (threadIdx.x == 0) {
    for (int s = 0; s < G::STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 8); }
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int nst = prm.nst;
  const int nst_total = 2 * nst;  // the two K segments [wA | B] . [B | wA]

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      bool ok = true;
      const uint64_t pol = (prm.debug & 16) ? policy_evict_normal() : policy_evict_first();
      for (int64_t t = blockIdx.x; t < prm.ntiles && ok; t += gridDim.x) {
        const unsigned long long tw = prm.tiles[t];
        const int i0 = (int)(tw & 0xffff) * BJ, j0 = (int)((tw >> 16) & 0xffff) * BJ, k0 = (int)((tw >> 32) & 0xffff) * BJ,
                  l0 = (int)((tw >> 48) & 0xffff) * BL;
        for (int pr = 0; pr < 3 && ok; ++pr) {
          const int rfir = pr == 0 ? k0 : (pr == 1 ? j0 : i0);   // first index of the row pairs (second: l)
          const int cfir = pr == 2 ? j0 : i0;                    // first index of the column pairs
          const int csec = pr == 0 ? j0 : k0;                    // second index of the column pairs
          for (int st = 0; st < nst_total; ++st) {
            if (!mbar_wait(bar_empty + 8 * stage, ph ^ 1, abort_flag)) { ok = false; break; }
            const uint32_t full = bar_full + 8 * stage;
            if (prm.debug & 4) { mbar_arrive(full); if (++stage == G::STAGES) { stage = 0; ph ^= 1; } continue; }
            mbar_expect_tx(full, G::STAGE_BYTES);
            const bool seg1 = st >= nst;
            const int kc = seg1 ? st - nst : st;  // K chunk
            const uint32_t base = saddr(stages + (size_t)stage * G::STAGE_BYTES);
            tma_load_box(base, seg1 ? &mBhr : &mAhr, l0, rfir, kc, full, pol);
            tma_load_box(base + G::ROW_BYTES, seg1 ? &mBlr : &mAlr, l0, rfir, kc, full, pol);
            tma_load_box(base + 2 * G::ROW_BYTES, seg1 ? &mAhc : &mBhc, csec, cfir, kc, full, pol);
            tma_load_box(base + 2 * G::ROW_BYTES + G::COL_BYTES, seg1 ? &mAlc : &mBlc, csec, cfir, kc, full, pol);
            if (++stage == G::STAGES) { stage = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TM, TN);
      int stage = 0;
      uint32_t ph = 0;
      int buf = 0;
      uint32_t tph[2] = {0, 0};
      bool ok = true;
      for (int64_t t = blockIdx.x; t < prm.ntiles && ok; t += gridDim.x) {
        for (int pr = 0; pr < 3 && ok; ++pr) {
          for (int c0 = 0; c0 < nst_total && ok; c0 += G::CHAIN_STAGES) {
            if (!mbar_wait(bar_tempty + 8 * buf, tph[buf] ^ 1, abort_flag)) { ok = false; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + (uint32_t)buf * TN;
            const int c1 = c0 + G::CHAIN_STAGES < nst_total ? c0 + G::CHAIN_STAGES : nst_total;
            for (int st = c0; st < c1; ++st) {
              if (!mbar_wait(bar_full + 8 * stage, ph, abort_flag)) { ok = false; break; }
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t base = saddr(stages + (size_t)stage * G::STAGE_BYTES);
#pragma unroll
              for (int kk = 0; kk < ((prm.debug & 8) ? 0 : KCH / 8); ++kk) {  // one MMA = 8 tf32 along K = 32 bytes inside the swizzled row
                const uint32_t ko = kk * 32;
                const uint64_t dAh = make_desc<KCH>(base + ko), dAl = make_desc<KCH>(base + G::ROW_BYTES + ko);
                const uint64_t dBh = make_desc<KCH>(base + 2 * G::ROW_BYTES + ko), dBl = make_desc<KCH>(base + 2 * G::ROW_BYTES + G::COL_BYTES + ko);
                mma_tf32(tacc, dAh, dBh, idesc, (st > c0 || kk > 0) ? 1u : 0u);
                mma_tf32(tacc, dAh, dBl, idesc, 1u);
                mma_tf32(tacc, dAl, dBh, idesc, 1u);
              }
              mma_commit(bar_empty + 8 * stage);  // the stage is free once these MMAs have read it
              if (++stage == G::STAGES) { stage = 0; ph ^= 1; }
            }
            if (!ok) break;
            mma_commit(bar_tfull + 8 * buf);  // the chain is complete: hand the accumulator to the epilogue
            tph[buf] ^= 1;
            buf ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    const int lq = warp & 3;                   // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;          // which 128 of the 256 columns
    const int m = lq * 32 + lane;              // row of the tile
    int buf = 0;
    uint32_t fph[2] = {0, 0};
    bool ok = true;
    const uint64_t opol = (prm.debug & 16) ? policy_evict_normal() : policy_evict_last();
    long long t_wait = 0, t_drain = 0, t_write = 0, t_fence = 0, t_bar = 0;  // debug 64: where an epilogue warp spends its time
    float acc[TN / 2];
    for (int64_t t = blockIdx.x; t < prm.ntiles && ok; t += gridDim.x) {
      for (int pr = 0; pr < 3 && ok; ++pr) {
#pragma unroll
        for (int c = 0; c < TN / 2; ++c) acc[c] = 0.f;
        for (int c0 = 0; c0 < nst_total && ok; c0 += G::CHAIN_STAGES) {
          const long long c_a = clock64();
          if (!mbar_wait(bar_tfull + 8 * buf, fph[buf], abort_flag)) { ok = false; break; }
          const long long c_b = clock64();
          t_wait += c_b - c_a;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr0 = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * TN + half * (TN / 2));
#pragma unroll
          for (int piece = 0; piece < ((prm.debug & 2) ? 0 : TN / 2); piece += 16) {
            uint32_t v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr0 + (uint32_t)piece));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[piece + j] += __uint_as_float(v[j]);
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
          t_drain += clock64() - c_b;
          fph[buf] ^= 1;
          buf ^= 1;
        }
        if (!ok) break;
        // ---- this pairing's share of the outputs: 1/6 of (G[rows|cols] + G[cols|rows])
        const float sixth = 1.0f / 6.0f;
        const long long c_w = clock64();
        if (!(prm.debug & 1)) {
          // The tile's 16 x 16 x 16 x 8 outputs go to ITS OWN contiguous 128 KB slot of a scratch buffer, at
          // L(i, j, k, l) = ((i 16 + j) 16 + k) 8 + l  -- not straight into the packed output: there the 256 (i, j) pairs of a
          // tile sit in ~100 different 2 MB pages and every scattered add of the epilogue cost a page walk (~500 clocks per
          // instruction: the adds took longer than the GEMMs).  In the slot the first pairing's rows (k, l) are whole
          // 128-byte lines per warp store, the other pairings add 4 sectors per warp instruction, everything within one page
          // and (evict-last) in L2; sym22_scatter_kernel then moves the slots to the packed layout with plain stores.
          float* slot = prm.scratch + (size_t)t * (TM * TN);
          const int x = m >> 3, z = m & 7;
          if (pr == 0) {  // rows (k, l), columns (i, j): L = n 128 + m
#pragma unroll
            for (int c = 0; c < TN / 2; ++c)
              asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(slot + (half * (TN / 2) + c) * TM + m), "f"(acc[c] * sixth), "l"(opol) : "memory");
          } else {
            // pairing 1: rows (j, l), columns (i, k): L = ((y 16 + x) 16 + w) 8 + z;  pairing 2: rows (i, l), columns (j, k): L = ((x 16 + y) 16 + w) 8 + z
#pragma unroll
            for (int c = 0; c < TN / 2; ++c) {
              const int n = half * (TN / 2) + c, y = n >> 4, w = n & 15;
              const int L = ((pr == 1 ? y * 16 + x : x * 16 + y) * 16 + w) * 8 + z;
              asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(slot + L), "f"(acc[c] * sixth), "l"(opol) : "memory");
            }
          }
        }
        // a component gets its three adds from three different threads of this CTA: keep them in pairing order (fp32 adds do
        // not commute bit for bit) -- nobody starts the next pairing's adds before everybody's adds of this one are performed
        const long long c_f = clock64();
        __threadfence();
        const long long c_g = clock64();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        t_write += c_f - c_w;
        t_fence += c_g - c_f;
        t_bar += clock64() - c_g;
      }
    }
    if ((prm.debug & 64) && lane == 0 && prm.err) {
      unsigned long long* dbg = reinterpret_cast<unsigned long long*>(prm.err) + 8;
      atomicAdd(dbg + 0, (unsigned long long)t_wait);
      atomicAdd(dbg + 1, (unsigned long long)t_drain);
      atomicAdd(dbg + 2, (unsigned long long)t_write);
      atomicAdd(dbg + 3, (unsigned long long)t_fence);
      atomicAdd(dbg + 4, (unsigned long long)t_bar);
      atomicAdd(dbg + 5, 1ULL);
    }
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && *abort_flag && prm.err) *prm.err = 1;
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// The same for ONE contracted index (k = 1: BASELINE config 3), where J is its own flat rank and has multiplicity 1: a CTA
// per (chunk, first), 32-bit index arithmetic, the rank of the sorted triple from per-value terms in shared memory
// (rank = C(d+2, 3) - 1 - F2[x] - F1[y] - F0[z]) -- ~25 instructions per element instead of ~300 (26 ms -> HBM time per operand).
__global__ void __launch_bounds__(256) expand_pairs_k1_kernel(PlanView P, const float* __restrict__ flat, float* __restrict__ hi,
                                                              float* __restrict__ lo, int Kp, int kch) {
  extern __shared__ int32_t F[];  // [3][d]: F[t][v] = C(d - 1 + t - v, t + 1)
  const int d = (int)P.dim;
  for (int e = threadIdx.x; e < 3 * d; e += blockDim.x) {
    const int t = e / d, v = e % d;
    F[e] = (int32_t)binom_at(P.binom, P.rank, d - 1 + t - v, t + 1);
  }
  __syncthreads();
  const int base = (int)(binom_at(P.binom, P.rank, d + 2, 3) - 1);
  const int nchunk = Kp / kch;
  for (int ca = blockIdx.x; ca < nchunk * d; ca += gridDim.x) {
    const int a = ca % d, chunk = ca / d;
    float* __restrict__ ho = hi + (int64_t)ca * d * kch;
    float* __restrict__ lo_o = lo + (int64_t)ca * d * kch;
    for (int e = threadIdx.x; e < d * kch; e += blockDim.x) {
      const int within = e % kch, b = e / kch;
      const int jj = chunk * kch + within;
      float v = 0.f;
      if (jj < d) {
        const int lo2 = min(a, b), hi2 = max(a, b);
        const int x = min(lo2, jj), z = max(hi2, jj), y = max(lo2, min(hi2, jj));
        v = flat[base - F[2 * d + x] - F[d + y] - F[z]];
      }
      const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      ho[e] = h;
      lo_o[e] = v - h;
    }
  }
}

// X[chunk][first][second][kch]: for all (first, second) in dim x dim the values X[first, second, J], J = chunk * kch + e < Kp
// (zero beyond K), stored K-chunk-major so that an operand box of the kernel -- 8 or 16 values of `first` times 16 consecutive
// values of `second`, one chunk -- is a few contiguous runs of 16 * kch * 4 bytes (with J innermost every 64-byte piece sat in
// its own 4 KB row: the kernel then ran at the speed of scattered HBM reads).  hi = the tf32 part the tensor core reads (the
// fp32 value with the low 13 mantissa bits cleared), lo = the exact remainder.  `weighted`: times the multiplicity of J.
__global__ void __launch_bounds__(256) expand_pairs_kernel(PlanView P, int k, const float* __restrict__ flat, float* __restrict__ hi,
                                                           float* __restrict__ lo, int64_t K, int64_t Kp, int kch, int weighted) {
  const int64_t d = P.dim;
  const int64_t total = d * d * Kp;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t within = e % kch, rowc = e / kch;    // rowc = (chunk * d + first) * d + second
    const int64_t b = rowc % d, ca = rowc / d;
    const int64_t a = ca % d, chunk = ca / d;
    const int64_t jj = chunk * kch + within;
    float v = 0.f;
    if (jj < K) {
      int32_t J[ST_MAX_RANK], m[ST_MAX_RANK + 2];
      flat_unrank_r(P, jj, k, J);
      int o = 0, q = 0;
      const int f0 = (int)(a < b ? a : b), f1 = (int)(a < b ? b : a);
      // merge (f0, f1) with the sorted tuple J
      int placed = 0;
      while (placed < 2 || q < k) {
        const int fv = placed == 0 ? f0 : f1;
        if (q >= k || (placed < 2 && fv <= J[q])) { m[o++] = fv; ++placed; }
        else m[o++] = J[q++];
      }
      double w = 1.0;
      if (weighted) {
        int rep = 0;
        for (int s = 0; s < k; ++s) {
          rep = (s > 0 && J[s] == J[s - 1]) ? rep + 1 : 1;
          w *= (double)(s + 1) / (double)rep;
        }
      }
      v = (float)(w * (double)flat[flat_rank_r(P, m, k + 2)]);
    }
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[e] = h;
    lo[e] = v - h;
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map(CUtensorMap* map, float* base, int64_t d, int64_t Kp, int kch, int box_second, int box_first) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return ST_ERR_UNSUPPORTED; }
  // X[chunk][first][second][kch]
  const cuuint64_t dims[4] = {(cuuint64_t)kch, (cuuint64_t)d, (cuuint64_t)d, (cuuint64_t)(Kp / kch)};
  const cuuint64_t strides[3] = {(cuuint64_t)kch * 4, (cuuint64_t)d * kch * 4, (cuuint64_t)d * d * kch * 4};
  const cuuint32_t box[4] = {(cuuint32_t)kch, (cuuint32_t)box_second, (cuuint32_t)box_first, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         kch == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return ST_ERR_CUDA; }
  return ST_OK;
}

int g_debug = 0;
int g_kch = 32;  // floats per stage row: 32 (SWIZZLE_128B, 2 stages of 96 KB; measured faster: 548 vs 654 ms on a 1/8 share of config 3) or 16 (SWIZZLE_64B, 4 stages); tuning key "sym22_kch"

static int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

int64_t contracted_count(int k, int64_t dim) {
  const HostPlan* hp = get_host_plan(std::max(k, 1), dim);
  if (!hp) return -1;
  // C(dim + k - 1, k)
  __int128 r = 1;
  for (int i = 1; i <= k; ++i) r = r * (dim + k - i) / i;  // exact: product of i consecutive integers is divisible by i!
  return (int64_t)r;
}

int workspace_bytes(int k, int64_t dim, int64_t* out) {
  const int64_t K = contracted_count(k, dim);
  if (K < 0) return ST_ERR_INVALID;
  const int64_t Kp = round_up(std::max<int64_t>(K, 1), 32);
  // control block, the four expanded operand arrays, the tile slots of one batch (at most kBatchTiles tiles of 128 KB)
  const int64_t nbj = (dim + BJ - 1) / BJ, nbl = (dim + BL - 1) / BL;
  const __int128 all_tiles = (__int128)nbj * nbj * nbj * nbl;
  const __int128 slots = all_tiles < (__int128)kBatchTiles ? all_tiles : (__int128)kBatchTiles;
  const __int128 bytes = (__int128)4 * dim * dim * Kp * 4 + 4096 + slots * (TM * TN * 4);
  if (bytes > (__int128)INT64_MAX) { set_error("workspace does not fit int64"); return ST_ERR_OVERFLOW; }
  *out = (int64_t)bytes;
  return ST_OK;
}

struct TileKey {
  int dev;
  int64_t dim;
  std::vector<std::pair<int64_t, int64_t>> ranges;
  bool operator<(const TileKey& o) const { return std::tie(dev, dim, ranges) < std::tie(o.dev, o.dim, o.ranges); }
};
struct TileList { unsigned long long* d; int64_t n; };
static std::mutex g_tmu;
static std::map<TileKey, TileList> g_tiles;

static int64_t rank1111(const HostPlan* hp, int64_t base, int64_t a, int64_t b, int64_t c, int64_t e) {
  const int64_t d = hp->dim;
  auto bn = [&](int64_t n, int k) { return n < 0 ? (int64_t)0 : hp->h_binom[n * 5 + k]; };
  return base - bn(d - 1 - a, 4) - bn(d - 1 - b, 3) - bn(d - 1 - c, 2) - bn(d - 1 - e, 1);
}

// tiles (p, q, r, s) that can hold components of [begin, end); s runs fastest so that the CTAs working side by side
// share the (i, j), (i, k) and (j, k) operand boxes in L2.  Pure host arithmetic (tests/test_cabi.py checks it against a
// brute-force enumeration through st_debug_sym22_tiles).
int g_tile_rgroup = 1;  // tiles of this many consecutive k blocks are interleaved in the list (1: s fastest); tuning key "sym22_rgroup"

void build_tiles(const HostPlan* hp, const std::vector<std::pair<int64_t, int64_t>>& ranges, std::vector<unsigned long long>& tiles) {
  const int64_t d = hp->dim;
  const int64_t off4 = hp->h_cls[hp->ncls - 1].offset;
  const int64_t base = off4 + hp->h_binom[d * 5 + 4] - 1;
  const int64_t nbj = (d + BJ - 1) / BJ, nbl = (d + BL - 1) / BL;
  tiles.clear();
  // Components with repeated indices live in the classes before (1,1,1,1).  In each of them the FIRST class-order value a
  // (the value with the highest multiplicity; the smaller one in class (2,2)) is the most significant digit of the position,
  // so the part of [begin, end) inside the class is an interval of a -- and a is a REPEATED index of the component: a tile
  // can hold it only where two neighbouring index blocks overlap in a.  (Round 1 gave every range that touched these classes
  // ALL diagonal tiles: 244,250 of them at dim 1000, all on the first GPU of a partition.)
  std::vector<int64_t> alo, ahi;
  bool any = false;
  {
    const PlanView P = hp->host_view();
    for (const auto& rg : ranges) {
      if (rg.second <= rg.first) continue;
      any = true;
      for (int c = 0; c + 1 < hp->ncls; ++c) {
        const ClassDesc& C = hp->h_cls[c];
        const int64_t lo = std::max(rg.first, C.offset), hi = std::min(rg.second, C.offset + C.size);
        if (lo >= hi) continue;
        int32_t v0[ST_MAX_RANK], v1[ST_MAX_RANK];
        permcls_unrank_vals(P, C, lo - C.offset, v0);
        permcls_unrank_vals(P, C, hi - 1 - C.offset, v1);
        alo.push_back(v0[0]);
        ahi.push_back(v1[0]);
      }
    }
  }
  if (!any) return;
  const int na = (int)alo.size();
  auto hits = [&](int64_t x0, int64_t x1) {  // does the value interval [x0, x1] meet one of the intervals of a?
    for (int i = 0; i < na; ++i)
      if (x0 <= ahi[i] && x1 >= alo[i]) return true;
    return false;
  };
  for (int64_t p = 0; p < nbj; ++p) {
    const int64_t i0 = p * BJ;
    for (int64_t q = p; q < nbj; ++q) {
      const int64_t j0 = q * BJ;
      for (int64_t r = q; r < nbj; ++r) {
        const int64_t k0 = r * BJ;
        for (int64_t s = k0 / BL; s < nbl; ++s) {
          const int64_t l0 = s * BL;
          const bool diag = p == q || q == r || l0 <= k0 + BJ - 1;
          bool take = false;
          if (diag && na) {  // a repeated index can sit in I_p (p == q), in I_q (q == r), or in the overlap of I_r and L_s
            if (p == q && hits(i0, i0 + BJ - 1)) take = true;
            if (q == r && hits(j0, j0 + BJ - 1)) take = true;
            if (l0 <= k0 + BJ - 1 && hits(std::max(k0, l0), std::min(k0 + BJ - 1, l0 + BL - 1))) take = true;
          }
          if (!take) {
            // smallest / largest strictly increasing tuple of the tile (componentwise bounds; lexicographic rank is monotone)
            const int64_t a0 = i0, b0 = std::max(j0, a0 + 1), c0 = std::max(k0, b0 + 1), e0 = std::max(l0, c0 + 1);
            const int64_t e1 = std::min(l0 + BL - 1, d - 1), c1 = std::min(k0 + BJ - 1, e1 - 1), b1 = std::min(j0 + BJ - 1, c1 - 1),
                          a1 = std::min(i0 + BJ - 1, b1 - 1);
            if (b0 <= j0 + BJ - 1 && c0 <= k0 + BJ - 1 && e0 <= e1 && a1 >= a0 && b1 >= b0 && c1 >= c0) {
              const int64_t lo = rank1111(hp, base, a0, b0, c0, e0), hi = rank1111(hp, base, a1, b1, c1, e1);
              for (const auto& rg : ranges) take = take || (rg.second > rg.first && lo < rg.second && hi >= rg.first);
            }
          }
          if (take) tiles.push_back((unsigned long long)p | ((unsigned long long)q << 16) | ((unsigned long long)r << 32) | ((unsigned long long)s << 48));
        }
      }
    }
  }
  // Order.  Generated with s fastest, the CTAs working side by side share the three column boxes (p, q), (p, r), (q, r) but every
  // one of their row boxes (r, s), (q, s), (p, s) is its own (6 MB per tile from HBM; the GEMM kernel reads 3.6 TB/s of DRAM
  // with the tensor pipe 72 % busy).  Interleaving groups of consecutive r (same p, q) s-major -- neighbours then share (q, s)
  // and (p, s) and differ in the column boxes of r -- was measured and is SLOWER: 515 ms (s fastest) / 519 / 549 / 568 / 576 ms for
  // groups of 2 / 4 / 8 / 16 on 1/8 of BASELINE config 3; the column boxes are twice the size of the row boxes.  Kept as a knob.
  if (g_tile_rgroup > 1) {
    const unsigned long long G = (unsigned long long)g_tile_rgroup;
    auto key = [G](unsigned long long w) {
      const unsigned long long p = w & 0xffff, q = (w >> 16) & 0xffff, r = (w >> 32) & 0xffff, sidx = (w >> 48) & 0xffff;
      return std::make_tuple(p, q, r / G, sidx, r);
    };
    std::stable_sort(tiles.begin(), tiles.end(), [&](unsigned long long a, unsigned long long b) { return key(a) < key(b); });
  }
}

void clear_tile_cache() {
  std::lock_guard<std::mutex> lk(g_tmu);
  for (auto& kv : g_tiles) cudaFree(kv.second.d);
  g_tiles.clear();
}

static int get_tiles(const HostPlan* hp, const std::vector<std::pair<int64_t, int64_t>>& ranges, TileList* out) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_tmu);
  const TileKey key{dev, hp->dim, ranges};
  auto it = g_tiles.find(key);
  if (it != g_tiles.end()) { *out = it->second; return ST_OK; }
  std::vector<unsigned long long> tiles;
  build_tiles(hp, ranges, tiles);
  TileList tl{nullptr, (int64_t)tiles.size()};
  if (tl.n) {
    rc = check_cuda(cudaMalloc(&tl.d, tiles.size() * sizeof(unsigned long long)), "cudaMalloc(tiles)");
    if (rc) return rc;
    rc = check_cuda(cudaMemcpy(tl.d, tiles.data(), tiles.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice), "cudaMemcpy(tiles)");
    if (rc) return rc;
  }
  if (g_tiles.size() > 64) {  // bounded cache
    for (auto& kv : g_tiles) cudaFree(kv.second.d);
    g_tiles.clear();
  }
  g_tiles[key] = tl;
  *out = tl;
  return ST_OK;
}

template <int KCH>
static int launch(const CUtensorMap* maps, const Params& prm, int grid, cudaStream_t stream) {
  using G = Geo<KCH>;
  const size_t smem = (size_t)G::STAGES * G::STAGE_BYTES + (2 * G::STAGES + 4) * 8 + 64;
  int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(sym22_umma_kernel<KCH>), (int)smem);
  if (rc) return rc;
  sym22_umma_kernel<KCH><<<grid, NTHREADS, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], prm);
  count_launch();
  return check_cuda(cudaGetLastError(), "sym22_umma_kernel");
}

// d_ws: workspace_bytes(); layout: [0, 4096) control (error flag), then wA hi, wA lo, B hi, B lo (dim x dim x Kp floats each)
int tensordot_sym22_ranges(int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, int nranges, const int64_t* begins,
                           const int64_t* ends, float* const* d_outs, void* d_ws, cudaStream_t stream);

int tensordot_sym22(int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out, int64_t begin, int64_t end, void* d_ws,
                    cudaStream_t stream) {
  return tensordot_sym22_ranges(k, dim, d_a_flat, d_b_flat, 1, &begin, &end, &d_out, d_ws, stream);
}

// the same for several disjoint output ranges at once (a tile that serves two of them runs once)
int tensordot_sym22_ranges(int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, int nranges, const int64_t* begins,
                           const int64_t* ends, float* const* d_outs, void* d_ws, cudaStream_t stream) {
  if (nranges < 1 || nranges > kMaxRanges) { set_error("1 .. %d output ranges per call", kMaxRanges); return ST_ERR_INVALID; }
  OutRanges R;
  R.n = nranges;
  std::vector<std::pair<int64_t, int64_t>> ranges;
  for (int q = 0; q < kMaxRanges; ++q) {
    R.begin[q] = q < nranges ? begins[q] : 0;
    R.end[q] = q < nranges ? ends[q] : 0;
    R.out[q] = q < nranges ? d_outs[q] : nullptr;
    if (q < nranges) ranges.emplace_back(begins[q], ends[q]);
  }
  if (dim >= 65536 * BL) { set_error("dim too large for the tile words"); return ST_ERR_UNSUPPORTED; }
  const int64_t K = contracted_count(k, dim);
  if (K <= 0) { set_error("nothing to contract"); return ST_ERR_INVALID; }
  const int kch = g_kch == 32 ? 32 : 16;
  const int64_t Kp = round_up(K, 32);
  const HostPlan* hp = get_host_plan(4, dim);
  if (!hp) return ST_ERR_INVALID;
  PlanView Pin, Pn;
  int rc = get_device_plan(k + 2, dim, &Pin);
  if (rc) return rc;
  rc = get_device_plan(4, dim, &Pn);
  if (rc) return rc;
  int* d_err = reinterpret_cast<int*>(d_ws);
  float* x = reinterpret_cast<float*>(reinterpret_cast<char*>(d_ws) + 4096);
  const int64_t n1 = dim * dim * Kp;
  float *ah = x, *al = x + n1, *bh = x + 2 * n1, *bl = x + 3 * n1;
  rc = check_cuda(cudaMemsetAsync(d_err, 0, 4096, stream), "cudaMemsetAsync(ctl)");
  if (rc) return rc;
  for (int q = 0; q < nranges; ++q) {
    if (ends[q] <= begins[q]) continue;
    rc = check_cuda(cudaMemsetAsync(d_outs[q], 0, (size_t)(ends[q] - begins[q]) * sizeof(float), stream), "cudaMemsetAsync(out)");
    if (rc) return rc;
  }
  const int eg = (int)std::min<int64_t>((n1 + 255) / 256, 148 * 32);
  if (k == 1 && dim <= 4096 && Kp % kch == 0 && Pin.flat_size < 2147483647LL) {
    const int cg = (int)std::min<int64_t>((Kp / kch) * dim, 148 * 64);
    expand_pairs_k1_kernel<<<cg, 256, 3 * dim * sizeof(int32_t), stream>>>(Pin, d_a_flat, ah, al, (int)Kp, kch);
    expand_pairs_k1_kernel<<<cg, 256, 3 * dim * sizeof(int32_t), stream>>>(Pin, d_b_flat, bh, bl, (int)Kp, kch);
  } else {
    expand_pairs_kernel<<<eg, 256, 0, stream>>>(Pin, k, d_a_flat, ah, al, K, Kp, kch, 1);
    expand_pairs_kernel<<<eg, 256, 0, stream>>>(Pin, k, d_b_flat, bh, bl, K, Kp, kch, 0);
  }
  count_launch(2);
  rc = check_cuda(cudaGetLastError(), "expand_pairs_kernel");
  if (rc) return rc;
  TileList tl;
  rc = get_tiles(hp, ranges, &tl);
  if (rc) return rc;
  if (tl.n == 0) return ST_OK;
  CUtensorMap maps[8];
  float* src[4] = {ah, al, bh, bl};
  for (int a = 0; a < 4 && !rc; ++a) rc = make_map(&maps[a], src[a], dim, Kp, kch, BL, BJ);      // row boxes (first: 16, l: 8)
  for (int a = 0; a < 4 && !rc; ++a) rc = make_map(&maps[4 + a], src[a], dim, Kp, kch, BJ, BJ);  // column boxes (16 x 16)
  if (rc) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* scratch = x + 4 * n1;
  const int64_t nbj = (dim + BJ - 1) / BJ, nbl = (dim + BL - 1) / BL;
  const int64_t cap = std::min<int64_t>(std::min<int64_t>(g_batch_tiles, kBatchTiles), nbj * nbj * nbj * nbl);
  // batches of tiles: GEMM kernel into the slots, then the scatter of the slots into the packed output (same stream).
  // (Measured in round 2: the scatter of a batch on a side stream, beside the GEMMs of the next batch and into a second slot
  // buffer, does not help -- 506 ms against 500 ms on 1/8 of BASELINE config 3: the GEMM kernel already draws 3.6 TB/s from
  // HBM, and the scatter is pure HBM traffic.)
  for (int64_t t0 = 0; t0 < tl.n; t0 += cap) {
    Params prm;
    prm.P = Pn;
    prm.begin = begins[0];
    prm.end = ends[0];
    prm.out = d_outs[0];
    prm.scratch = scratch;
    prm.tiles = tl.d + t0;
    prm.ntiles = std::min<int64_t>(cap, tl.n - t0);
    prm.nst = (int32_t)(Kp / kch);
    prm.debug = g_debug;
    prm.err = d_err;
    const int grid = (int)std::min<int64_t>(prm.ntiles, sms);
    rc = kch == 32 ? launch<32>(maps, prm, grid, stream) : launch<16>(maps, prm, grid, stream);
    if (rc) return rc;
    sym22_scatter_kernel<<<(unsigned)prm.ntiles, 256, 0, stream>>>(Pn, prm.tiles, scratch, R);
    count_launch();
    rc = check_cuda(cudaGetLastError(), "sym22_scatter_kernel");
    if (rc) return rc;
  }
  return ST_OK;
}

}  // namespace s22
}  // namespace st

extern "C" int64_t st_debug_sym22_tiles(int64_t dim, int64_t begin, int64_t end, unsigned long long* h_out, int64_t cap) {
  const st::HostPlan* hp = st::get_host_plan(4, dim);
  if (!hp) return -1;
  std::vector<unsigned long long> tiles;
  st::s22::build_tiles(hp, {{begin, end}}, tiles);
  for (int64_t i = 0; i < (int64_t)tiles.size() && i < cap; ++i) h_out[i] = tiles[i];
  return (int64_t)tiles.size();
}

extern "C" int64_t st_debug_sym22_tiles_ranges(int64_t dim, int nranges, const int64_t* begins, const int64_t* ends, unsigned long long* h_out,
                                               int64_t cap) {
  const st::HostPlan* hp = st::get_host_plan(4, dim);
  if (!hp || nranges < 0 || (nranges && (!begins || !ends))) return -1;
  std::vector<std::pair<int64_t, int64_t>> ranges;
  for (int q = 0; q < nranges; ++q) ranges.emplace_back(begins[q], ends[q]);
  std::vector<unsigned long long> tiles;
  st::s22::build_tiles(hp, ranges, tiles);
  for (int64_t i = 0; i < (int64_t)tiles.size() && i < cap; ++i) h_out[i] = tiles[i];
  return (int64_t)tiles.size();
}

// count and cost terms of the tile list of several ranges: stats[0] = tiles, stats[1] = sum over the tiles of 1 / (number of l
// blocks of the tile's k block) -- the fewer tiles share their three column boxes, the more a tile costs (sharding.py)
extern "C" int st_debug_sym22_tiles_stats(int64_t dim, int nranges, const int64_t* begins, const int64_t* ends, double* stats) {
  const st::HostPlan* hp = st::get_host_plan(4, dim);
  if (!hp || !stats || nranges < 0 || (nranges && (!begins || !ends))) return ST_ERR_INVALID;
  std::vector<std::pair<int64_t, int64_t>> ranges;
  for (int q = 0; q < nranges; ++q) ranges.emplace_back(begins[q], ends[q]);
  std::vector<unsigned long long> tiles;
  st::s22::build_tiles(hp, ranges, tiles);
  const int64_t nbl = (dim + st::s22::BL - 1) / st::s22::BL;
  double inv = 0.0;
  for (unsigned long long w : tiles) {
    const int64_t r = (int64_t)((w >> 32) & 0xffff);
    inv += 1.0 / (double)std::max<int64_t>(1, nbl - r * st::s22::BJ / st::s22::BL);
  }
  stats[0] = (double)tiles.size();
  stats[1] = inv;
  return ST_OK;
}
