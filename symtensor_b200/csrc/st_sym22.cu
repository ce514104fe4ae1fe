// Symmetrized tensordot with TWO free indices on each side (ra - k == 2 and rb - k == 2), non-materialising:
// BASELINE config 3 (rank 3 . rank 3 over one index, dim 1000, fp32 -> rank 4, 41.9 G components; symtensor/symalg.py:427-459).
//
//     C[i<=j<=k<=l] = 1/6 * ( G[ij|kl] + G[kl|ij] + G[ik|jl] + G[jl|ik] + G[il|jk] + G[jk|il] ),
//     G[ab|cd] = sum_J w_J A[a, b, J] B[J, c, d]          (J: the packed contracted tuples, w_J their multiplicity)
//
// The Gram matrix G (pair rows x pair columns, 1 TB at config 3) is never stored.  The output is cut into TILES of index
// blocks  I x J x K x L = 8 x 16 x 16 x 16  (I: block of the smallest index).  For one tile the six terms are THREE
// GEMMs of shape [128 x 256 x 2 Kc] -- the two orientations of a pairing share an accumulator through the concatenated
// contraction  G[ab|cd] + G[cd|ab] = sum_J [wA | B][ab, J] . [B | wA][cd, J]  -- with rows (i, .) = 8 x 16 = 128 and
// columns 16 x 16 = 256 for every pairing, which is the full-rate tcgen05 shape (M = 128, N = 256, cta_group::1):
//     pairing 0: rows (i, j), columns (k, l);   pairing 1: rows (i, k), columns (j, l);   pairing 2: rows (i, l), columns (j, k).
// Operands: the expanded pair matrices X[first][second][J] (dim x dim rows of Kp floats, both triangles) of wA and B, each
// pre-split into the tf32 part the tensor core reads and the exact remainder (3xTF32: hi.hi + hi.lo + lo.hi), written
// once by expand_pairs_kernel; a tile's operand boxes (8 x 16 or 16 x 16 pair rows x 16 or 32 floats) are fetched by TMA
// (cp.async.bulk.tensor.3d, SASS UTMALDG; hardware swizzle, out-of-range rows zero-filled) into a ring of shared-memory
// stages guarded by full / empty mbarriers.
//
// Warp roles (384 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane; allocates the 512 TMEM
// columns = two 128 x 256 fp32 accumulators), warps 4..11 = epilogue.  The tensor core accumulates into TMEM with
// truncation, so a CHAIN is at most 256 products long (st_ops.cu: measured 1.4e-5 of sum|terms| at 1024, 3.5e-6 at 256):
// after every chain the issuer commits the accumulator to the epilogue (tmem-full barrier) and continues in the other
// accumulator; the epilogue warps drain it with tcgen05.ld and add the chains in REGISTERS with round-to-nearest (each
// thread keeps 128 columns of its row).  At the end of a pairing the thread scales by 1/6 and adds its values straight
// into the packed permcls output with red.global.add.f32 (the position of (i, j, k, l) is a sum of four per-index terms,
// kept in shared-memory tables per tile); the output range is zeroed first and every component receives exactly three
// adds, all from the same CTA in a fixed order, so the result is deterministic.  Elements with repeated indices
// (diagonal tiles) go through the generic class rank.  [begin, end) ranges of the output are served by the tiles that
// intersect them -- the multi-GPU partition, no collective.
#include <cuda.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "st_common.cuh"

namespace st {
namespace s22 {

constexpr int BI = 8, BJ = 16;
constexpr int TM = BI * BJ;   // 128 rows
constexpr int TN = BJ * BJ;   // 256 columns
constexpr int NTHREADS = 384;
constexpr int CHAIN_K = 256;  // products per TMEM accumulation chain
constexpr int SMEM_STAGE_BUDGET = 192 * 1024;

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

struct Tables {
  long long colterm[3][TN];
  long long rowterm[3][TM];
};

struct Params {
  PlanView P;          // plan of the rank-4 output
  int64_t begin, end;  // packed output coordinates served by this launch
  float* out;          // points at coordinate `begin`
  const unsigned long long* tiles;  // p | q << 16 | r << 32 | s << 48
  int64_t ntiles;
  int32_t nst;         // stages per K segment (Kp / KCH)
  int32_t pad_;
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
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

// one output of a tile that is not "fast" (repeated indices, the ragged edge, or a launch range that cuts the tile)
__device__ __noinline__ void slow_add(const PlanView& P, int64_t base1111, int64_t begin, int64_t end, float* out, int gi, int gj, int gk, int gl,
                                      float v) {
  const int64_t coord = element_coord(P, base1111, gi, gj, gk, gl);
  if (coord >= begin && coord < end) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(out + (coord - begin)), "f"(v) : "memory");
}

template <int KCH>
__global__ void __launch_bounds__(NTHREADS, 1)
sym22_umma_kernel(const __grid_constant__ CUtensorMap mAhr, const __grid_constant__ CUtensorMap mAlr, const __grid_constant__ CUtensorMap mBhr,
                  const __grid_constant__ CUtensorMap mBlr, const __grid_constant__ CUtensorMap mAhc, const __grid_constant__ CUtensorMap mAlc,
                  const __grid_constant__ CUtensorMap mBhc, const __grid_constant__ CUtensorMap mBlc, const __grid_constant__ Params prm) {
  using G = Geo<KCH>;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stages = smem;
  Tables* tab = reinterpret_cast<Tables*>(smem + (size_t)G::STAGES * G::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(tab) + sizeof(Tables));
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
      for (int64_t t = blockIdx.x; t < prm.ntiles && ok; t += gridDim.x) {
        const unsigned long long tw = prm.tiles[t];
        const int i0 = (int)(tw & 0xffff) * BI, j0 = (int)((tw >> 16) & 0xffff) * BJ, k0 = (int)((tw >> 32) & 0xffff) * BJ,
                  l0 = (int)((tw >> 48) & 0xffff) * BJ;
        for (int pr = 0; pr < 3 && ok; ++pr) {
          const int rsec = pr == 0 ? j0 : (pr == 1 ? k0 : l0);   // second index of the row pairs (first: i)
          const int cfir = pr == 0 ? k0 : j0;                    // first index of the column pairs
          const int csec = pr == 2 ? k0 : l0;                    // second index of the column pairs
          for (int st = 0; st < nst_total; ++st) {
            if (!mbar_wait(bar_empty + 8 * stage, ph ^ 1, abort_flag)) { ok = false; break; }
            const uint32_t full = bar_full + 8 * stage;
            mbar_expect_tx(full, G::STAGE_BYTES);
            const bool seg1 = st >= nst;
            const int kc = (seg1 ? st - nst : st) * KCH;
            const uint32_t base = saddr(stages + (size_t)stage * G::STAGE_BYTES);
            tma_load_3d(base, seg1 ? &mBhr : &mAhr, kc, rsec, i0, full);
            tma_load_3d(base + G::ROW_BYTES, seg1 ? &mBlr : &mAlr, kc, rsec, i0, full);
            tma_load_3d(base + 2 * G::ROW_BYTES, seg1 ? &mAhc : &mBhc, kc, csec, cfir, full);
            tma_load_3d(base + 2 * G::ROW_BYTES + G::COL_BYTES, seg1 ? &mAlc : &mBlc, kc, csec, cfir, full);
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
              for (int kk = 0; kk < KCH / 8; ++kk) {  // one MMA = 8 tf32 along K = 32 bytes inside the swizzled row
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
    const PlanView& P = prm.P;
    const int et = threadIdx.x - 128;          // 0..255
    const int lq = warp & 3;                   // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;          // which 128 of the 256 columns
    const int m = lq * 32 + lane;              // row of the tile
    const int64_t base1111 = P.cls[P.ncls - 1].offset + binom_at(P.binom, 4, P.dim, 4) - 1;
    int buf = 0;
    uint32_t fph[2] = {0, 0};
    bool ok = true;
    float acc[TN / 2];
    for (int64_t t = blockIdx.x; t < prm.ntiles && ok; t += gridDim.x) {
      const unsigned long long tw = prm.tiles[t];
      const int i0 = (int)(tw & 0xffff) * BI, j0 = (int)((tw >> 16) & 0xffff) * BJ, k0 = (int)((tw >> 32) & 0xffff) * BJ,
                l0 = (int)((tw >> 48) & 0xffff) * BJ;
      // all indices distinct, inside the tensor and inside the launch range: positions are sums of table terms
      bool fast = i0 + BI - 1 < j0 && j0 + BJ - 1 < k0 && k0 + BJ - 1 < l0 && l0 + BJ - 1 < P.dim;
      if (fast) {
        const int64_t cmin = element_coord(P, base1111, i0, j0, k0, l0);
        const int64_t cmax = element_coord(P, base1111, i0 + BI - 1, j0 + BJ - 1, k0 + BJ - 1, l0 + BJ - 1);
        fast = cmin >= prm.begin && cmax < prm.end;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // everybody is done with the previous tile's tables
      if (fast) {
        // term of an index at output position q (0-based): C(d - 1 - g, 4 - q)
        {
          const int y = et >> 4, z = et & 15;  // column n = et: (first, second)
          const long long tj = binom_at(P.binom, 4, P.dim - 1 - (j0 + y), 3), tk_f = binom_at(P.binom, 4, P.dim - 1 - (k0 + y), 2);
          const long long tl = binom_at(P.binom, 4, P.dim - 1 - (l0 + z), 1), tk_s = binom_at(P.binom, 4, P.dim - 1 - (k0 + z), 2);
          tab->colterm[0][et] = tk_f + tl;   // columns (k, l)
          tab->colterm[1][et] = tj + tl;     // columns (j, l)
          tab->colterm[2][et] = tj + tk_s;   // columns (j, k)
        }
        if (et < TM) {
          const int i = et >> 4, x = et & 15;  // row m = et: (i, second)
          const long long ti = binom_at(P.binom, 4, P.dim - 1 - (i0 + i), 4);
          tab->rowterm[0][et] = ti + binom_at(P.binom, 4, P.dim - 1 - (j0 + x), 3);  // rows (i, j)
          tab->rowterm[1][et] = ti + binom_at(P.binom, 4, P.dim - 1 - (k0 + x), 2);  // rows (i, k)
          tab->rowterm[2][et] = ti + binom_at(P.binom, 4, P.dim - 1 - (l0 + x), 1);  // rows (i, l)
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int pr = 0; pr < 3 && ok; ++pr) {
#pragma unroll
        for (int c = 0; c < TN / 2; ++c) acc[c] = 0.f;
        for (int c0 = 0; c0 < nst_total && ok; c0 += G::CHAIN_STAGES) {
          if (!mbar_wait(bar_tfull + 8 * buf, fph[buf], abort_flag)) { ok = false; break; }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr0 = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * TN + half * (TN / 2));
#pragma unroll
          for (int piece = 0; piece < TN / 2; piece += 16) {
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
          fph[buf] ^= 1;
          buf ^= 1;
        }
        if (!ok) break;
        // ---- this pairing's share of the outputs: 1/6 of (G[rows|cols] + G[cols|rows])
        const float sixth = 1.0f / 6.0f;
        if (fast) {
          const int64_t row_off = base1111 - prm.begin - tab->rowterm[pr][m];
          const long long* ct = tab->colterm[pr] + half * (TN / 2);
#pragma unroll
          for (int c = 0; c < TN / 2; ++c) {
            float* p = prm.out + (row_off - ct[c]);
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(acc[c] * sixth) : "memory");
          }
        } else {
          const int i = m >> 4, x = m & 15;
#pragma unroll
          for (int c = 0; c < TN / 2; ++c) {
            const int n = half * (TN / 2) + c;
            const int y = n >> 4, z = n & 15;
            int gi = i0 + i, gj, gk, gl;
            if (pr == 0) { gj = j0 + x; gk = k0 + y; gl = l0 + z; }
            else if (pr == 1) { gk = k0 + x; gj = j0 + y; gl = l0 + z; }
            else { gl = l0 + x; gj = j0 + y; gk = k0 + z; }
            slow_add(P, base1111, prm.begin, prm.end, prm.out, gi, gj, gk, gl, acc[c] * sixth);
          }
        }
        // a component gets its three adds from three different threads of this CTA: keep them in pairing order (fp32 adds do
        // not commute bit for bit) -- nobody starts the next pairing's adds before everybody's adds of this one are performed
        __threadfence();
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
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

// X[first][second][J] for all (first, second) in dim x dim and J < Kp (zero beyond K): the tf32 part the tensor core reads
// (the fp32 value with the low 13 mantissa bits cleared) and the exact remainder.  `weighted`: times the multiplicity of J.
__global__ void __launch_bounds__(256) expand_pairs_kernel(PlanView P, int k, const float* __restrict__ flat, float* __restrict__ hi,
                                                           float* __restrict__ lo, int64_t K, int64_t Kp, int weighted) {
  const int64_t d = P.dim;
  const int64_t total = d * d * Kp;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / Kp, jj = e - row * Kp;
    float v = 0.f;
    if (jj < K) {
      const int a = (int)(row / d), b = (int)(row - (int64_t)a * d);
      int32_t J[ST_MAX_RANK], m[ST_MAX_RANK + 2];
      flat_unrank_r(P, jj, k, J);
      int o = 0, q = 0;
      const int f0 = a < b ? a : b, f1 = a < b ? b : a;
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
  const cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)d, (cuuint64_t)d};
  const cuuint64_t strides[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)d * Kp * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kch, (cuuint32_t)box_second, (cuuint32_t)box_first};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         kch == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return ST_ERR_CUDA; }
  return ST_OK;
}

int g_kch = 16;  // floats per stage row: 16 (SWIZZLE_64B, 4 stages) or 32 (SWIZZLE_128B, 2 stages); tuning key "sym22_kch"

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
  const __int128 bytes = (__int128)4 * dim * dim * Kp * 4 + 4096;
  if (bytes > (__int128)INT64_MAX) { set_error("workspace does not fit int64"); return ST_ERR_OVERFLOW; }
  *out = (int64_t)bytes;
  return ST_OK;
}

struct TileKey {
  int dev;
  int64_t dim, begin, end;
  bool operator<(const TileKey& o) const { return std::tie(dev, dim, begin, end) < std::tie(o.dev, o.dim, o.begin, o.end); }
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
void build_tiles(const HostPlan* hp, int64_t begin, int64_t end, std::vector<unsigned long long>& tiles) {
  const int64_t d = hp->dim;
  const int64_t off4 = hp->h_cls[hp->ncls - 1].offset;
  const int64_t base = off4 + hp->h_binom[d * 5 + 4] - 1;
  const int64_t nbi = (d + BI - 1) / BI, nbj = (d + BJ - 1) / BJ;
  tiles.clear();
  if (end <= begin) return;
  for (int64_t p = 0; p < nbi; ++p) {
    const int64_t i0 = p * BI;
    for (int64_t q = i0 / BJ; q < nbj; ++q) {
      const int64_t j0 = q * BJ;
      for (int64_t r = q; r < nbj; ++r) {
        const int64_t k0 = r * BJ;
        for (int64_t s = r; s < nbj; ++s) {
          const int64_t l0 = s * BJ;
          const bool diag = i0 + BI - 1 >= j0 || q == r || r == s;
          bool take = diag && begin < off4;  // components with repeated indices live in the classes before (1,1,1,1)
          if (!take) {
            // smallest / largest strictly increasing tuple of the tile (componentwise bounds; lexicographic rank is monotone)
            const int64_t a0 = i0, b0 = std::max(j0, a0 + 1), c0 = std::max(k0, b0 + 1), e0 = std::max(l0, c0 + 1);
            const int64_t e1 = std::min(l0 + BJ - 1, d - 1), c1 = std::min(k0 + BJ - 1, e1 - 1), b1 = std::min(j0 + BJ - 1, c1 - 1),
                          a1 = std::min(i0 + BI - 1, b1 - 1);
            if (b0 <= j0 + BJ - 1 && c0 <= k0 + BJ - 1 && e0 <= e1 && a1 >= a0 && b1 >= b0 && c1 >= c0) {
              const int64_t lo = rank1111(hp, base, a0, b0, c0, e0), hi = rank1111(hp, base, a1, b1, c1, e1);
              take = lo < end && hi >= begin;
            }
          }
          if (take) tiles.push_back((unsigned long long)p | ((unsigned long long)q << 16) | ((unsigned long long)r << 32) | ((unsigned long long)s << 48));
        }
      }
    }
  }
}

static int get_tiles(const HostPlan* hp, int64_t begin, int64_t end, TileList* out) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_tmu);
  const TileKey key{dev, hp->dim, begin, end};
  auto it = g_tiles.find(key);
  if (it != g_tiles.end()) { *out = it->second; return ST_OK; }
  std::vector<unsigned long long> tiles;
  build_tiles(hp, begin, end, tiles);
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
  const size_t smem = (size_t)G::STAGES * G::STAGE_BYTES + sizeof(Tables) + (2 * G::STAGES + 4) * 8 + 64;
  int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(sym22_umma_kernel<KCH>), (int)smem);
  if (rc) return rc;
  sym22_umma_kernel<KCH><<<grid, NTHREADS, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], prm);
  count_launch();
  return check_cuda(cudaGetLastError(), "sym22_umma_kernel");
}

// d_ws: workspace_bytes(); layout: [0, 4096) control (error flag), then wA hi, wA lo, B hi, B lo (dim x dim x Kp floats each)
int tensordot_sym22(int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out, int64_t begin, int64_t end, void* d_ws,
                    cudaStream_t stream) {
  if (dim >= 65536 * BI) { set_error("dim too large for the tile words"); return ST_ERR_UNSUPPORTED; }
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
  rc = check_cuda(cudaMemsetAsync(d_out, 0, (size_t)(end - begin) * sizeof(float), stream), "cudaMemsetAsync(out)");
  if (rc) return rc;
  const int eg = (int)std::min<int64_t>((n1 + 255) / 256, 148 * 32);
  expand_pairs_kernel<<<eg, 256, 0, stream>>>(Pin, k, d_a_flat, ah, al, K, Kp, 1);
  expand_pairs_kernel<<<eg, 256, 0, stream>>>(Pin, k, d_b_flat, bh, bl, K, Kp, 0);
  count_launch(2);
  rc = check_cuda(cudaGetLastError(), "expand_pairs_kernel");
  if (rc) return rc;
  TileList tl;
  rc = get_tiles(hp, begin, end, &tl);
  if (rc) return rc;
  if (tl.n == 0) return ST_OK;
  CUtensorMap maps[8];
  float* src[4] = {ah, al, bh, bl};
  for (int a = 0; a < 4 && !rc; ++a) rc = make_map(&maps[a], src[a], dim, Kp, kch, BJ, BI);      // row boxes (i: 8, second: 16)
  for (int a = 0; a < 4 && !rc; ++a) rc = make_map(&maps[4 + a], src[a], dim, Kp, kch, BJ, BJ);  // column boxes (16 x 16)
  if (rc) return rc;
  Params prm;
  prm.P = Pn;
  prm.begin = begin;
  prm.end = end;
  prm.out = d_out;
  prm.tiles = tl.d;
  prm.ntiles = tl.n;
  prm.nst = (int32_t)(Kp / kch);
  prm.pad_ = 0;
  prm.err = d_err;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<int64_t>(tl.n, sms);
  return kch == 32 ? launch<32>(maps, prm, grid, stream) : launch<16>(maps, prm, grid, stream);
}

}  // namespace s22
}  // namespace st

extern "C" int64_t st_debug_sym22_tiles(int64_t dim, int64_t begin, int64_t end, unsigned long long* h_out, int64_t cap) {
  const st::HostPlan* hp = st::get_host_plan(4, dim);
  if (!hp) return -1;
  std::vector<unsigned long long> tiles;
  st::s22::build_tiles(hp, begin, end, tiles);
  for (int64_t i = 0; i < (int64_t)tiles.size() && i < cap; ++i) h_out[i] = tiles[i];
  return (int64_t)tiles.size();
}
