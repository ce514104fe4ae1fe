// Packed-space implementations of the three tensor-valued ops of the hot path and the layout converters
// they use.  Every op works on packed components only -- the reference's defaults densify to d^r and average
// r! transposes (symtensor/symalg.py:206-283, 294-316, 427-459, 475-496); here the formulas of SURVEY.md A.3
// are evaluated directly:
//
//   multiply.outer   C_K = C(n,ra)^-1  sum_{S subset of positions, |S| = ra}  A[K_S] B[K_S^c]          (n = ra + rb)
//   tensordot        C_K = C(n,ra-k)^-1 sum_S sum_{J in [d]^k} A[K_S, J] B[J, K_S^c]                  (n = ra + rb - 2k)
//                        = C(n,ra-k)^-1 sum_S G[rank(K_S)][rank(K_S^c)],   G = Aexp . Bexp^T  (pair-packed Gram matrix,
//                          rows = packed free indices, columns = packed contracted tuples weighted by multiplicity)
//   contract_all_indices_with_matrix   the partially-symmetric mode chain
//                        T_{k+1}[j_1..j_{k+1}; I''] = sum_a W[a, j_{k+1}] T_k[j_1..j_k; sort(a, I'')],  T_0 = A, T_r = C
//                    every T_k is stored packed x packed (flat order in both index groups); the "unpack" of the
//                    contracted mode is a gather into the shared-memory tile of the step's GEMM.
//
// Operands and intermediates use the FLAT order (combinations_with_replacement, symtensor/flat_symtensor.py:
// 39-50): the position of a sorted sub-multi-index is a sum of r binomials, so gathering A[K_S] needs no sort
// (a subsequence of a sorted sequence is sorted) and no class lookup.  Results are written in the caller's
// layout (permcls ranges [begin, end) for sharding over GPUs).
#include <algorithm>
#include <vector>

#include "st_common.cuh"

namespace st {

// ------------------------------------------------------------------------------------------------------
// layout converters (also the "pack / unpack" step either side of the ops: permcls <-> flat re-ordering)
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void permcls_to_flat_kernel(PlanView P, const T* __restrict__ in, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P.flat_size; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t s[ST_MAX_RANK], vals[ST_MAX_RANK];
    flat_unrank_sorted(P, i, s);
    const int c = classify_index(P, s, vals);
    out[i] = in[P.cls[c].offset + permcls_rank_vals(P, P.cls[c], vals)];
  }
}

template <typename T>
__global__ void flat_to_permcls_kernel(PlanView P, const T* __restrict__ in, T* __restrict__ out, int64_t begin, int64_t end) {
  for (int64_t c = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < end; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t K[ST_MAX_RANK];
    out[c - begin] = permcls_coord_sorted(P, c, K) ? in[flat_rank_sorted(P, K)] : T(0);
  }
}

// Row-walk converters (round 2; ranks <= 8, dims <= 255): a warp serves 32 consecutive permcls coordinates, finds their
// sorted multi-indices with the warp-uniform odometer of st_common.cuh (RowCursor) and ranks them in the flat layout with
// per-entry terms from shared memory -- ~150 instructions per component instead of the ~2,500 of an unrank / classify /
// rank round trip (12 ms and 8 ms either side of contract_all_indices_with_matrix at BASELINE config 4).
constexpr int kConvSpan = 2048;

template <typename T, bool TO_FLAT>
__global__ void __launch_bounds__(256) rowwalk_convert_kernel(PlanView P, const T* __restrict__ in, T* __restrict__ out, int64_t begin, int64_t end) {
  __shared__ int64_t Fl[8 * 256];   // Fl[t][v] = C(d - 1 + t - v, t + 1): the flat rank is C(d+r-1, r) - 1 - sum_p Fl[r-1-p][K[p]]
  __shared__ int32_t B23[2 * kRowBinomStride];
  __shared__ unsigned long long cinfo[32];
  const int d = (int)P.dim, rk = P.rank;
  for (int e = threadIdx.x; e < 8 * 256; e += blockDim.x) {
    const int tt = e >> 8, v = e & 255;
    Fl[e] = (v < d && tt < rk) ? binom_at(P.binom, P.rank, d - 1 + tt - v, tt + 1) : 0;
  }
  for (int e = threadIdx.x; e < 2 * kRowBinomStride; e += blockDim.x) {
    const int n = e % kRowBinomStride;
    B23[e] = e < kRowBinomStride ? n * (n - 1) / 2 : n * (n - 1) * (n - 2) / 6;
  }
  for (int e = threadIdx.x; e < P.ncls && e < 32; e += blockDim.x) cinfo[e] = row_class_info(P.cls[e], rk);
  __syncthreads();
  const int64_t base = P.flat_size - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t ntasks = (end - begin + kConvSpan - 1) / kConvSpan;
  for (int64_t task = (int64_t)blockIdx.x * wpb + warp; task < ntasks; task += (int64_t)gridDim.x * wpb) {
    const int64_t s0 = begin + task * kConvSpan, s1 = s0 + kConvSpan < end ? s0 + kConvSpan : end;
    RowCursor rc;
    rowcursor_seek(P, rc, s0);
    for (int64_t b = s0; b < s1; b += 32) {
      const int64_t c = b + lane, be = b + 32 < s1 ? b + 32 : s1;
      RowLatch R;
      rowcursor_serve(P, rc, c, be, R);
      if (c >= be) continue;
      if (R.state != 1) {
        if (!TO_FLAT) out[c - begin] = T(0);
        continue;
      }
      int32_t K[8];
      row_component(R.valsp, R.b, R.m, R.o, cinfo[R.ci], B23, K);
      int64_t pos = base;
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (p < rk) pos -= Fl[((rk - 1 - p) << 8) + K[p]];
      if (TO_FLAT) out[pos] = in[c];
      else out[c - begin] = in[pos];
    }
  }
}

// debug: the row walk on the DEVICE, exactly as the kernels run it (st_debug_rowwalk replays it on the host)
__global__ void __launch_bounds__(256) rowwalk_debug_kernel(PlanView P, int64_t begin, int64_t end, int span, int32_t* __restrict__ idx, int32_t* __restrict__ dbg) {
  __shared__ int32_t B23[2 * kRowBinomStride];
  __shared__ unsigned long long cinfo[32];
  for (int e = threadIdx.x; e < 2 * kRowBinomStride; e += blockDim.x) {
    const int n = e % kRowBinomStride;
    B23[e] = e < kRowBinomStride ? n * (n - 1) / 2 : n * (n - 1) * (n - 2) / 6;
  }
  for (int e = threadIdx.x; e < P.ncls && e < 32; e += blockDim.x) cinfo[e] = row_class_info(P.cls[e], P.rank);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t ntasks = (end - begin + span - 1) / span;
  for (int64_t task = (int64_t)blockIdx.x * wpb + warp; task < ntasks; task += (int64_t)gridDim.x * wpb) {
    const int64_t s0 = begin + task * span, s1 = s0 + span < end ? s0 + span : end;
    RowCursor rc;
    rowcursor_seek(P, rc, s0);
    for (int64_t b = s0; b < s1; b += 32) {
      const int64_t c = b + lane, be = b + 32 < s1 ? b + 32 : s1;
      RowLatch R;
      rowcursor_serve(P, rc, c, be, R);
      if (c >= be) continue;
      int32_t* o = idx + (c - begin) * P.rank;
      if (dbg) {
        int32_t* g = dbg + (c - begin) * 8;
        g[0] = (int32_t)(R.valsp & 0xffffffffu); g[1] = (int32_t)(R.valsp >> 32); g[2] = R.b; g[3] = R.m; g[4] = R.o; g[5] = R.ci; g[6] = R.state;
        g[7] = (int32_t)(rc.valsp & 0xffffffffu);
      }
      if (R.state != 1) { for (int i = 0; i < P.rank; ++i) o[i] = -1; continue; }
      int32_t K[8];
      row_component(R.valsp, R.b, R.m, R.o, cinfo[R.ci], B23, K);
      for (int i = 0; i < P.rank; ++i) o[i] = K[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// symmetrized outer product: one thread per packed output component, C(n, ra) gathered products
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) outer_kernel(PlanView P, int ra, int rb, const T* __restrict__ af, const T* __restrict__ bf,
                                                    T* __restrict__ out, int64_t begin, int64_t end, double inv_count, int op) {
  const int n = ra + rb;
  for (int64_t c = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < end; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t K[ST_MAX_RANK];
    if (!permcls_coord_sorted(P, c, K)) { out[c - begin] = T(0); continue; }
    double acc = 0.0;
    // all position subsets of size ra, as bit masks in increasing order (Gosper's hack)
    for (uint32_t mask = (1u << ra) - 1u; mask < (1u << n);) {
      int32_t sa[ST_MAX_RANK], sb[ST_MAX_RANK];
      int ia = 0, ib = 0;
      for (int p = 0; p < n; ++p) {
        if ((mask >> p) & 1u) sa[ia++] = K[p]; else sb[ib++] = K[p];
      }
      const double va = (double)af[flat_rank_r(P, sa, ra)], vb = (double)bf[flat_rank_r(P, sb, rb)];
      acc += op == ST_OUTER_MULTIPLY ? va * vb : op == ST_OUTER_ADD ? va + vb : va - vb;  // (warp-uniform)
      if (ra == 0) break;
      const uint32_t lo = mask & (0u - mask), hi = mask + lo;
      mask = (((mask ^ hi) >> 2) / lo) | hi;
    }
    out[c - begin] = (T)(acc * inv_count);
  }
}

// fused outer -> contract_all_indices_with_vector: sum_K gamma_K C_K prod x[K], the rank-n tensor is never stored
template <typename T>
__global__ void __launch_bounds__(256) outer_vec_kernel(PlanView P, int ra, int rb, const T* __restrict__ af, const T* __restrict__ bf,
                                                        const T* __restrict__ x, double* __restrict__ partials, int64_t begin, int64_t end,
                                                        double inv_count) {
  __shared__ double red[32];
  const int n = ra + rb;
  double total = 0.0;
  for (int64_t c = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < end; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t K[ST_MAX_RANK];
    if (!permcls_coord_sorted(P, c, K)) continue;
    double acc = 0.0;
    for (uint32_t mask = (1u << ra) - 1u; mask < (1u << n);) {
      int32_t sa[ST_MAX_RANK], sb[ST_MAX_RANK];
      int ia = 0, ib = 0;
      for (int p = 0; p < n; ++p) {
        if ((mask >> p) & 1u) sa[ia++] = K[p]; else sb[ib++] = K[p];
      }
      acc += (double)af[flat_rank_r(P, sa, ra)] * (double)bf[flat_rank_r(P, sb, rb)];
      if (ra == 0) break;
      const uint32_t lo = mask & (0u - mask), hi = mask + lo;
      mask = (((mask ^ hi) >> 2) / lo) | hi;
    }
    double w = (double)P.cls[class_of_coord(P, c)].gamma;
    for (int p = 0; p < n; ++p) w *= (double)x[K[p]];
    total += acc * inv_count * w;
  }
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = total;
  __syncthreads();
  if (warp == 0) {
    double s = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) partials[blockIdx.x] = s;
  }
}

// ---- symmetrized outer product, compile-time ranks (RA + RB <= 8) -----------------------------------------------
// The flat rank of a sorted r-tuple s is  C(d+r-1, r) - 1 - sum_i F[r-1-i][s[i]],  F[t][v] = C(d-1+t-v, t+1): a sum of
// per-element terms.  F (a few KB) lives in shared memory, the per-component values G[t][p] = F[t][K[p]] in
// registers, and the C(n, RA) position subsets are unrolled at compile time: every subset costs n - 2 integer
// adds, two gathers and one FMA (the run-time-rank kernel spends ~150 instructions per subset on splitting K,
// ranking twice through the global binomial table and Gosper's hack).
__host__ __device__ constexpr int popc_const(unsigned m) { return m == 0 ? 0 : (int)(m & 1u) + popc_const(m >> 1); }

template <typename T, int RA, int RB>
__device__ __forceinline__ double outer_terms(const int32_t* __restrict__ F, int d, const int32_t* K, const T* __restrict__ af,
                                              const T* __restrict__ bf, int base_a, int base_b) {
  constexpr int N = RA + RB, TM = RA > RB ? RA : RB;
  int32_t G[TM][N];
#pragma unroll
  for (int t = 0; t < TM; ++t)
#pragma unroll
    for (int p = 0; p < N; ++p) G[t][p] = F[t * d + K[p]];
  double acc = 0.0;
#pragma unroll
  for (unsigned mask = 0; mask < (1u << N); ++mask) {
    if (__popc(mask) != RA) continue;  // (folded after unrolling; a constexpr helper here was compiled into a RUNTIME popcount loop per mask: 45 % of the kernel's instructions)
    int sa = 0, sb = 0, ia = 0, ib = 0;
#pragma unroll
    for (int p = 0; p < N; ++p) {
      if ((mask >> p) & 1u) { sa += G[RA - 1 - ia][p]; ++ia; }
      else { sb += G[RB - 1 - ib][p]; ++ib; }
    }
    if (sizeof(T) == 4) acc += (double)((float)af[base_a - sa] * (float)bf[base_b - sb]);  // fp32: one rounding of the product (6e-8), fp64 sum
    else acc += (double)af[base_a - sa] * (double)bf[base_b - sb];
  }
  return acc;
}

template <typename T, int RA, int RB, bool VEC>
__global__ void __launch_bounds__(256) outer_fast_kernel(PlanView P, const T* __restrict__ af, const T* __restrict__ bf, T* __restrict__ out,
                                                         const T* __restrict__ x, double* __restrict__ partials, int64_t begin, int64_t end,
                                                         double inv_count) {
  constexpr int N = RA + RB, TM = RA > RB ? RA : RB;
  extern __shared__ int32_t Fs[];  // [TM][d]
  __shared__ double red[32];
  const int d = (int)P.dim;
  for (int e = threadIdx.x; e < TM * d; e += blockDim.x) {
    const int tt = e / d, v = e % d;
    Fs[e] = (int32_t)binom_at(P.binom, P.rank, d - 1 + tt - v, tt + 1);
  }
  __syncthreads();
  const int base_a = (int)(binom_at(P.binom, P.rank, d + RA - 1, RA) - 1), base_b = (int)(binom_at(P.binom, P.rank, d + RB - 1, RB) - 1);
  double total = 0.0;
  // A CTA takes CHUNKS of consecutive coordinates (not a grid-stride): consecutive components share their leading indices,
  // so the operand entries a chunk gathers (~4 distinct floats per component) stay in L1 -- the operands (494 KB each at
  // BASELINE config 5) do not fit L1 as a whole, and with a grid-stride every 256 components started cold (L2 latency).
  constexpr int64_t kChunk = 8192;
  const int64_t nchunks = (end - begin + kChunk - 1) / kChunk;
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x)
  for (int64_t c = begin + ch * kChunk + threadIdx.x; c < end && c < begin + (ch + 1) * kChunk; c += blockDim.x) {
    int32_t K[ST_MAX_RANK];
    if (!permcls_coord_sorted(P, c, K)) {
      if (!VEC) out[c - begin] = T(0);
      continue;
    }
    int32_t Kr[N];
#pragma unroll
    for (int p = 0; p < N; ++p) Kr[p] = K[p];
    const double acc = outer_terms<T, RA, RB>(Fs, d, Kr, af, bf, base_a, base_b);
    if (VEC) {
      double w = (double)P.cls[class_of_coord(P, c)].gamma;
#pragma unroll
      for (int p = 0; p < N; ++p) w *= (double)x[Kr[p]];
      total += acc * inv_count * w;
    } else {
      out[c - begin] = (T)(acc * inv_count);
    }
  }
  if (VEC) {
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = total;
    __syncthreads();
    if (warp == 0) {
      double s = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) partials[blockIdx.x] = s;
    }
  }
}

// ---- symmetrized outer product, row walk (round 2) ----------------------------------------------------------------
// outer_fast_kernel spends ~70 % of its instructions on the per-component unrank (divisions, binary searches) and most of
// the rest on the n - 2 integer adds per subset.  Here
//  * a warp serves 32 CONSECUTIVE coordinates at a time (lane = coordinate: coalesced stores, neighbouring operand entries)
//    and finds their multi-indices by walking the few rows that cover them with a warp-uniform odometer (RowCursor in
//    st_common.cuh): one full unrank per span of kSpan coordinates instead of one per component;
//  * the rank sums of a subset are split into the low and the high half of the sorted positions: the partial sums of the
//    2^(n/2) half-subsets are formed once per component (registers), a subset then costs one add per operand.
// 256 threads, at most 128 registers (two CTAs per SM): (128 threads, 96 registers) and (256, 80) spill and measured 15 % slower.
// dim <= 255 (the latched multi-index is packed in bytes, the rank-term table has a fixed row stride).
// (A first version walked rows of ONE varying value -- ~4 components each at rank 8 dim 40, seven rows per batch -- and was
// slower than the per-component unrank: the uniform walk costs whole warp instructions, 3,700 per batch against ~600 for
// the 32 components' products.  With the last three values in the lanes' hands a row has ~300 components.)
constexpr int kRowsStride = 256;  // entries per rank-term row in shared memory
constexpr int kRowsSpan = 2048;   // coordinates per warp task (one seek each)

template <typename T, int RA, int RB>
__device__ __forceinline__ double outer_terms_split(const int32_t* __restrict__ Fs, const int32_t* K, const T* __restrict__ af,
                                                    const T* __restrict__ bf, int base_a, int base_b) {
  constexpr int N = RA + RB, TM = RA > RB ? RA : RB, L = N / 2, H = N - L;
  // high half: the partial rank sums of all its half-subsets (registers); low half: formed per half-subset in the loop below
  int32_t saHi[1 << H], sbHi[1 << H];
  {
    int32_t G[TM][H];
#pragma unroll
    for (int p = 0; p < H; ++p)
#pragma unroll
      for (int t = 0; t < TM; ++t) G[t][p] = Fs[t * kRowsStride + K[L + p]];
#pragma unroll
    for (unsigned hi = 0; hi < (1u << H); ++hi) {
      const int jh = __popc(hi);
      if (jh > RA || RA - jh > L || H - jh > RB) continue;
      int sa = 0, sb = 0, ia = RA - jh, ib = L - (RA - jh);  // entries the low half has taken
#pragma unroll
      for (int p = 0; p < H; ++p) {
        if ((hi >> p) & 1u) { sa += G[RA - 1 - ia][p]; ++ia; }
        else { sb += G[RB - 1 - ib][p]; ++ib; }
      }
      saHi[hi] = sa; sbHi[hi] = sb;
    }
  }
  int32_t G[TM][L > 0 ? L : 1];
#pragma unroll
  for (int p = 0; p < L; ++p)
#pragma unroll
    for (int t = 0; t < TM; ++t) G[t][p] = Fs[t * kRowsStride + K[p]];
  double acc = 0.0;
#pragma unroll
  for (unsigned lo = 0; lo < (1u << L); ++lo) {
    const int ja = __popc(lo);
    if (ja > RA || L - ja > RB) continue;
    int saLo = base_a, sbLo = base_b;
    {
      int ia = 0, ib = 0;
#pragma unroll
      for (int p = 0; p < L; ++p) {
        if ((lo >> p) & 1u) { saLo -= G[RA - 1 - ia][p]; ++ia; }
        else { sbLo -= G[RB - 1 - ib][p]; ++ib; }
      }
    }
    if (sizeof(T) == 4) {
      float part = 0.f;  // fp32: the products of one low half-subset (at most C(H, H/2) of them) in fp32, the halves in fp64
#pragma unroll
      for (unsigned hi = 0; hi < (1u << H); ++hi) {
        if (__popc(hi) != RA - ja) continue;
        part = fmaf((float)af[saLo - saHi[hi]], (float)bf[sbLo - sbHi[hi]], part);
      }
      acc += (double)part;
    } else {
#pragma unroll
      for (unsigned hi = 0; hi < (1u << H); ++hi) {
        if (__popc(hi) != RA - ja) continue;
        acc += (double)af[saLo - saHi[hi]] * (double)bf[sbLo - sbHi[hi]];
      }
    }
  }
  return acc;
}

template <typename T, int RA, int RB, bool VEC, int THREADS = 256, int MINB = 2>
__global__ void __launch_bounds__(THREADS, MINB) outer_rows_kernel(PlanView P, const T* __restrict__ af, const T* __restrict__ bf, T* __restrict__ out,
                                                         const T* __restrict__ x, double* __restrict__ partials, int64_t begin, int64_t end,
                                                         double inv_count) {
  constexpr int N = RA + RB, TM = RA > RB ? RA : RB;
  __shared__ int32_t Fs[TM * kRowsStride];
  __shared__ int32_t B23[2 * kRowBinomStride];  // C(n, 2), C(n, 3): the lanes' decode of their offset in a row
  __shared__ unsigned long long cinfo[32];      // row_class_info per class (a rank <= 8 has at most 22 classes)
  __shared__ double red[32];
  const int d = (int)P.dim;
  for (int e = threadIdx.x; e < TM * kRowsStride; e += blockDim.x) {
    const int tt = e / kRowsStride, v = e % kRowsStride;
    Fs[e] = v < d ? (int32_t)binom_at(P.binom, P.rank, d - 1 + tt - v, tt + 1) : 0;
  }
  for (int e = threadIdx.x; e < 2 * kRowBinomStride; e += blockDim.x) {
    const int n = e % kRowBinomStride;
    B23[e] = e < kRowBinomStride ? n * (n - 1) / 2 : n * (n - 1) * (n - 2) / 6;
  }
  for (int e = threadIdx.x; e < P.ncls && e < 32; e += blockDim.x) cinfo[e] = row_class_info(P.cls[e], N);
  __syncthreads();
  const int base_a = (int)(binom_at(P.binom, P.rank, d + RA - 1, RA) - 1), base_b = (int)(binom_at(P.binom, P.rank, d + RB - 1, RB) - 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double total = 0.0;
  const int64_t ntasks = (end - begin + kRowsSpan - 1) / kRowsSpan;
  for (int64_t task = (int64_t)blockIdx.x * wpb + warp; task < ntasks; task += (int64_t)gridDim.x * wpb) {
    const int64_t s0 = begin + task * kRowsSpan, s1 = s0 + kRowsSpan < end ? s0 + kRowsSpan : end;
    RowCursor rc;
    rowcursor_seek(P, rc, s0);
    for (int64_t b = s0; b < s1; b += 32) {
      const int64_t c = b + lane, be = b + 32 < s1 ? b + 32 : s1;
      RowLatch R;
      rowcursor_serve(P, rc, c, be, R);
      if (c >= be) continue;
      if (R.state != 1) {
        if (!VEC) out[c - begin] = T(0);
        continue;
      }
      int32_t K[8];
      row_component(R.valsp, R.b, R.m, R.o, cinfo[R.ci], B23, K);
      const double acc = outer_terms_split<T, RA, RB>(Fs, K, af, bf, base_a, base_b);
      if (VEC) {
        double w = (double)P.cls[R.ci].gamma;
#pragma unroll
        for (int p = 0; p < N; ++p) w *= (double)x[K[p]];
        total += acc * inv_count * w;
      } else {
        out[c - begin] = (T)(acc * inv_count);
      }
    }
  }
  if (VEC) {
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) red[warp] = total;
    __syncthreads();
    if (warp == 0) {
      double s = lane < wpb ? red[lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) partials[blockIdx.x] = s;
    }
  }
}

// tensordot epilogue with compile-time free ranks: C_K = inv_count * sum_S G[rank(K_S)][rank(K_S^c)] (same scheme as
// outer_fast_kernel: rank terms from shared memory, subsets unrolled)
template <typename T, int NA, int NB>
__global__ void __launch_bounds__(256) gram_gather_fast_kernel(PlanView P, const T* __restrict__ G, int64_t ncols, T* __restrict__ out, int64_t begin,
                                                               int64_t end, double inv_count) {
  constexpr int N = NA + NB, TM = NA > NB ? NA : NB;
  extern __shared__ int32_t Fs[];  // [TM][d]
  const int d = (int)P.dim;
  for (int e = threadIdx.x; e < TM * d; e += blockDim.x) {
    const int tt = e / d, v = e % d;
    Fs[e] = (int32_t)binom_at(P.binom, P.rank, d - 1 + tt - v, tt + 1);
  }
  __syncthreads();
  const int base_a = (int)(binom_at(P.binom, P.rank, d + NA - 1, NA) - 1), base_b = (int)(binom_at(P.binom, P.rank, d + NB - 1, NB) - 1);
  for (int64_t c = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < end; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t K[ST_MAX_RANK];
    if (!permcls_coord_sorted(P, c, K)) { out[c - begin] = T(0); continue; }
    int32_t Gt[TM][N];
#pragma unroll
    for (int tt = 0; tt < TM; ++tt)
#pragma unroll
      for (int p = 0; p < N; ++p) Gt[tt][p] = Fs[tt * d + K[p]];
    double acc = 0.0;
#pragma unroll
    for (unsigned mask = 0; mask < (1u << N); ++mask) {
      if (__popc(mask) != NA) continue;
      int sa = 0, sb = 0, ia = 0, ib = 0;
#pragma unroll
      for (int p = 0; p < N; ++p) {
        if ((mask >> p) & 1u) { sa += Gt[NA - 1 - ia][p]; ++ia; }
        else { sb += Gt[NB - 1 - ib][p]; ++ib; }
      }
      acc += (double)G[(int64_t)(base_a - sa) * ncols + (base_b - sb)];
    }
    out[c - begin] = (T)(acc * inv_count);
  }
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double* out64, float* out32) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) {
      if (out64) *out64 = t;
      if (out32) *out32 = (float)t;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// tensordot: expansion of an operand to [packed free indices] x [packed contracted tuples]
//   exp[p][J] = w(J) * flat[rank(sort(F_p, J))],  w(J) = multiplicity of the sorted k-tuple J (A side) or 1 (B side)
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void expand_kernel(PlanView P, int rfree, int k, const T* __restrict__ flat, T* __restrict__ ex, int64_t nfree, int64_t ncon,
                              int weighted) {
  const int64_t total = nfree * ncon;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / ncon, j = i - p * ncon;
    int32_t f[ST_MAX_RANK], J[ST_MAX_RANK], m[ST_MAX_RANK];
    flat_unrank_r(P, p, rfree, f);
    flat_unrank_r(P, j, k, J);
    // merge the two sorted tuples
    int a = 0, b = 0, o = 0;
    while (a < rfree || b < k) m[o++] = (b >= k || (a < rfree && f[a] <= J[b])) ? f[a++] : J[b++];
    double w = 1.0;
    if (weighted) {  // k! / prod(count!) distinct orderings of J
      int rep = 0;
      for (int q = 0; q < k; ++q) {
        rep = (q > 0 && J[q] == J[q - 1]) ? rep + 1 : 1;
        w *= (double)(q + 1) / (double)rep;
      }
    }
    ex[i] = (T)(w * (double)flat[flat_rank_r(P, m, rfree + k)]);
  }
}

// C[M x N] = A[M x K] . B[N x K]^T  (row-major, any sizes): 64 x 64 tile per CTA, 4 x 4 outputs per thread
template <typename T>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, int64_t M, int64_t N,
                                                      int64_t K) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ T As[TK][TM + 4];
  __shared__ T Bs[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int64_t k0 = 0; k0 < K; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int r = e / TK, kk = e % TK;
      As[kk][r] = (m0 + r < M && k0 + kk < K) ? A[(m0 + r) * K + k0 + kk] : T(0);
      Bs[kk][r] = (n0 + r < N && k0 + kk < K) ? B[(n0 + r) * K + k0 + kk] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[m * N + n] = acc[i][j];
    }
}

// ------------------------------------------------------------------------------------------------------
// fp32 Gram matrix on the 5th-generation tensor cores: C[M x N] = A[M x K] * B[N x K]^T with tcgen05.mma kind::tf32
// (SASS UTCHMMA), accumulator in TMEM, fp32 accuracy through the 3xTF32 split
//     a = a_hi + a_lo (a_hi: the 19 bits the tensor core reads, a_lo: the exact remainder),
//     a b ~ a_hi b_hi + a_hi b_lo + a_lo b_hi                       (relative error ~2^-21 per product)
// and two-level accumulation: the tensor core adds into its fp32 accumulator with truncation (measured: the error
// grows with the chain length, 1.4e-5 of sum|terms| at K = 1024), so a chain is at most 256 long and the chains are
// added in registers with round-to-nearest.  One CTA (4 warps) per 128 x 128 tile; the operand chunks (32 along K)
// are split and laid out by the CTA's threads in the canonical K-major no-swizzle layout (core matrices of 8 rows x
// 16 bytes; cute::UMMA::SmemDescriptor / InstrDescriptor bit layouts), one elected thread issues the MMAs and
// commits them to an mbarrier.  Three CTAs share an SM (66 KB of shared memory, 128 TMEM columns each), which is
// what overlaps one CTA's staging with another's MMAs in this first version (tools/membench/umma_probe.cu:
// 38.6 TFLOP/s at 4096 x 4096 x 1024 with the split, 4x the CUDA-core kernel it replaces).
// ------------------------------------------------------------------------------------------------------
namespace umma {
constexpr int BM = 128, BN = 128, BK = 32, CHAIN = 256;
constexpr uint32_t LBO = 128, SBO = (BK / 4) * 128;  // bytes: next 16-byte K chunk, next 8-row group
constexpr int TILE_BYTES = (BM / 8) * SBO;           // 16 KB
constexpr size_t SMEM_BYTES = 4 * TILE_BYTES + 1024;

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {  // D = F32, A = B = TF32, K-major, M x N
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
__device__ __forceinline__ bool wait_bounded(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < (1 << 26); ++it) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
}  // namespace umma

__global__ void __launch_bounds__(128) gram_umma_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int64_t M,
                                                        int64_t N, int64_t K) {
  using namespace umma;
  extern __shared__ __align__(1024) unsigned char smem_umma[];
  unsigned char* tiles = smem_umma;  // A hi, A lo, B hi, B lo
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&tmem_base_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc(BM, BN);
  const bool vec4 = (K & 3) == 0;
  float acc[BN];  // this thread's row of the tile: the chains are added here with round-to-nearest
#pragma unroll
  for (int j = 0; j < BN; ++j) acc[j] = 0.f;
  uint32_t phase = 0;
  bool ok = true;
  // drain the TMEM accumulator into the registers (warp w owns TMEM lanes 32 w .. 32 w + 31 = rows of the tile)
  auto drain = [&]() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[c0 + j] += __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };
  int in_chain = 0;  // K elements accumulated in TMEM since the last drain
  for (int64_t k0 = 0; k0 < K; k0 += BK) {
    // stage: 16-byte chunks (row r, chunk kc) of the A and B operand chunks, split into hi / lo
    for (int e = tid; e < BM * (BK / 4); e += 128) {
      const int r = e / (BK / 4), kc = e % (BK / 4);
      const int64_t kk = k0 + kc * 4;
      float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec4) {
        if (m0 + r < M && kk < K) { const float4 v = *reinterpret_cast<const float4*>(A + (m0 + r) * K + kk); a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
        if (n0 + r < N && kk < K) { const float4 v = *reinterpret_cast<const float4*>(B + (n0 + r) * K + kk); b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w; }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (m0 + r < M && kk + q < K) a[q] = A[(m0 + r) * K + kk + q];
          if (n0 + r < N && kk + q < K) b[q] = B[(n0 + r) * K + kk + q];
        }
      }
      float4 ah, al, bh, bl;
      ah.x = tf32_hi(a[0]); ah.y = tf32_hi(a[1]); ah.z = tf32_hi(a[2]); ah.w = tf32_hi(a[3]);
      bh.x = tf32_hi(b[0]); bh.y = tf32_hi(b[1]); bh.z = tf32_hi(b[2]); bh.w = tf32_hi(b[3]);
      al.x = a[0] - ah.x; al.y = a[1] - ah.y; al.z = a[2] - ah.z; al.w = a[3] - ah.w;
      bl.x = b[0] - bh.x; bl.y = b[1] - bh.y; bl.z = b[2] - bh.z; bl.w = b[3] - bh.w;
      const uint32_t off = (uint32_t)(r >> 3) * SBO + (uint32_t)kc * LBO + (uint32_t)(r & 7) * 16;
      *reinterpret_cast<float4*>(tiles + 0 * TILE_BYTES + off) = ah;
      *reinterpret_cast<float4*>(tiles + 1 * TILE_BYTES + off) = al;
      *reinterpret_cast<float4*>(tiles + 2 * TILE_BYTES + off) = bh;
      *reinterpret_cast<float4*>(tiles + 3 * TILE_BYTES + off) = bl;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core (async proxy) reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = saddr(tiles);
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {  // one MMA = 8 tf32 along K = two 16-byte chunks
        const uint32_t koff = ks * 2 * LBO;
        const uint64_t dAh = make_desc(base + 0 * TILE_BYTES + koff), dAl = make_desc(base + 1 * TILE_BYTES + koff);
        const uint64_t dBh = make_desc(base + 2 * TILE_BYTES + koff), dBl = make_desc(base + 3 * TILE_BYTES + koff);
        mma_tf32(tmem_base, dAh, dBh, idesc, (in_chain > 0 || ks > 0) ? 1u : 0u);
        mma_tf32(tmem_base, dAh, dBl, idesc, 1u);
        mma_tf32(tmem_base, dAl, dBh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(&bar)) : "memory");
    }
    in_chain += BK;
    // the tiles are overwritten by the next stage: wait for the MMAs of this one
    // (the decision to give up is taken by the whole block: a per-thread exit would leave the barriers below divergent)
    if (!__syncthreads_and(wait_bounded(saddr(&bar), phase) ? 1 : 0)) { ok = false; break; }
    phase ^= 1;
    if (in_chain >= CHAIN || k0 + BK >= K) {
      drain();
      in_chain = 0;
    }
    __syncthreads();
  }
  const int64_t row = m0 + warp * 32 + lane;
  if (row < M) {
    const float poison = __uint_as_float(0x7fc00000u);  // a timed-out barrier must not pass as a result
#pragma unroll
    for (int j = 0; j < BN; ++j)
      if (n0 + j < N) C[row * N + n0 + j] = ok ? acc[j] : poison;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

// epilogue: C_K = inv_count * sum_S G[rank(K_S)][rank(K_S^c)], written in the permcls layout of rank n
template <typename T>
__global__ void __launch_bounds__(256) gram_gather_kernel(PlanView P, int na, int nb, const T* __restrict__ G, int64_t ncols, T* __restrict__ out,
                                                          int64_t begin, int64_t end, double inv_count) {
  const int n = na + nb;
  for (int64_t c = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < end; c += (int64_t)gridDim.x * blockDim.x) {
    int32_t K[ST_MAX_RANK];
    if (!permcls_coord_sorted(P, c, K)) { out[c - begin] = T(0); continue; }
    double acc = 0.0;
    for (uint32_t mask = (1u << na) - 1u; mask < (1u << n);) {
      int32_t sa[ST_MAX_RANK], sb[ST_MAX_RANK];
      int ia = 0, ib = 0;
      for (int p = 0; p < n; ++p) {
        if ((mask >> p) & 1u) sa[ia++] = K[p]; else sb[ib++] = K[p];
      }
      acc += (double)G[flat_rank_r(P, sa, na) * ncols + flat_rank_r(P, sb, nb)];
      if (na == 0) break;
      const uint32_t lo = mask & (0u - mask), hi = mask + lo;
      mask = (((mask ^ hi) >> 2) / lo) | hi;
    }
    out[c - begin] = (T)(acc * inv_count);
  }
}

// ------------------------------------------------------------------------------------------------------
// matrix contraction: one step of the mode chain
//   Tn[(J, j)][I] = sum_a W[a][j] Tk[J][sort(a, I)]     J: sorted k-tuple, j >= max(J), I: sorted m-tuple (m = r-k-1)
// CTA: one J, a tile of TI rows I and TJ columns j; the A-operand tile S[i][a] is GATHERED from the packed
// row Tk[J][.] (the contracted mode is unpacked in shared memory only), W streams through shared memory.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) mat_step_kernel(PlanView P, int k, int m, const T* __restrict__ Tk, const T* __restrict__ W,
                                                       T* __restrict__ Tn, int64_t nJ, int64_t nI, int64_t nI1, int64_t jbase, int jlo, int jhi) {
  // rows J = jbase + blockIdx.x / tilesI (a range of the sorted k-tuples: the multi-GPU partition by the first output mode);
  // columns j restricted to [jlo, jhi) (step 0 of a partition; [0, d) otherwise).  Tk / Tn are indexed by GLOBAL rows.
  constexpr int TI = 64, TJ = 64, TK = 16;
  __shared__ T Ss[TK][TI + 4];
  __shared__ T Ws[TK][TJ + 4];
  __shared__ int32_t Is[TI][ST_MAX_RANK];  // the sorted m-tuples of the tile's rows
  __shared__ int32_t Js[ST_MAX_RANK];
  const int64_t d = P.dim;
  const int64_t tilesI = (nI + TI - 1) / TI;
  const int64_t jidx = jbase + blockIdx.x / tilesI, i0 = (blockIdx.x % tilesI) * TI;
  const int64_t j0 = (int64_t)blockIdx.y * TJ;
  if (threadIdx.x == 0) flat_unrank_r(P, jidx, k, Js);
  for (int r = threadIdx.x; r < TI; r += 256)
    if (i0 + r < nI) flat_unrank_r(P, i0 + r, m, Is[r]);
  __syncthreads();
  const int jlast = max(k ? Js[k - 1] : 0, jlo);
  if (j0 + TJ <= jlast || j0 >= jhi) return;  // every column of this tile is below max(J) (not a sorted (k+1)-tuple) or outside the range
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const T* __restrict__ row = Tk + jidx * nI1;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int64_t a0 = 0; a0 < d; a0 += TK) {
    for (int e = threadIdx.x; e < TI * TK; e += 256) {
      const int r = e % TI, aa = e / TI;
      const int64_t a = a0 + aa;
      T v = T(0);
      if (i0 + r < nI && a < d) {
        // flat rank of sort(a, I): walk the merged tuple from its largest element down
        int64_t pos = binom_at(P.binom, P.rank, d + m, m + 1) - 1;
        int q = m - 1;
        bool placed = false;
        for (int t = 0; t <= m; ++t) {
          int32_t z;
          if (!placed && (q < 0 || Is[r][q] <= a)) { z = (int32_t)a; placed = true; }
          else z = Is[r][q--];
          pos -= binom_at(P.binom, P.rank, d - 1 + t - z, t + 1);
        }
        v = row[pos];
      }
      Ss[aa][r] = v;
    }
    for (int e = threadIdx.x; e < TJ * TK; e += 256) {
      const int c = e % TJ, aa = e / TJ;
      Ws[aa][c] = (a0 + aa < d && j0 + c < d) ? W[(a0 + aa) * d + j0 + c] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Ss[kk][ty * 4 + i]; b[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  // store: row of the output = flat rank of the sorted (k+1)-tuple (J, j)
  int32_t Jn[ST_MAX_RANK];
  for (int q = 0; q < k; ++q) Jn[q] = Js[q];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t jj = j0 + tx * 4 + j;
    if (jj >= d || jj < jlast || jj >= jhi) continue;
    Jn[k] = (int32_t)jj;
    const int64_t orow = flat_rank_r(P, Jn, k + 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t ii = i0 + ty * 4 + i;
      if (ii < nI) Tn[orow * nI + ii] = acc[i][j];
    }
  }
}

// ---- fp64 mode-chain step on the FP64 tensor pipe (mma.sync m8n8k4, DMMA) ---------------------------------
// flat rank (rank m + 1) of sort(a, I): the merged tuple walked from its largest element down
ST_HD int64_t merged_rank(const PlanView& P, const int32_t* I, int m, int64_t a) {
  const int64_t d = P.dim;
  int64_t pos = binom_at(P.binom, P.rank, d + m, m + 1) - 1;
  int q = m - 1;
  bool placed = false;
  for (int t = 0; t <= m; ++t) {
    int32_t z;
    if (!placed && (q < 0 || I[q] <= a)) { z = (int32_t)a; placed = true; }
    else z = I[q--];
    pos -= binom_at(P.binom, P.rank, d - 1 + t - z, t + 1);
  }
  return pos;
}

// gather map of one step, built once and shared by all J: tbl[a * nI + i] = flat rank of sort(a, I_i).
// With F[t][v] = C(d-1+t-v, t+1) the rank of the merged tuple is  base - PS[p] - F[m-p][a],  p = #{q: I[q] <= a},
// PS[p] = sum_{q<p} F[m-q][I[q]] + sum_{q>=p} F[m-1-q][I[q]]:  one unrank of I, then a few instructions per a (p only grows).
// A thread takes kIdxRun consecutive I: one unrank, then successors of the sorted tuple.  (Runs of 4 were measured SLOWER, 3.3 ms
// against 2.9 ms for step 0 of BASELINE config 4: the stores of a warp then fall 16 bytes apart.)
constexpr int kIdxRun = 1;
__global__ void __launch_bounds__(256) mat_index_kernel(PlanView P, int m, int64_t nI, int32_t* __restrict__ tbl) {
  const int64_t i0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * kIdxRun;
  if (i0 >= nI) return;
  int32_t I[ST_MAX_RANK];
  flat_unrank_r(P, i0, m, I);
  const int64_t d = P.dim;
  const int64_t base = binom_at(P.binom, P.rank, d + m, m + 1) - 1;
  for (int rr = 0; rr < kIdxRun && i0 + rr < nI; ++rr) {
    const int64_t i = i0 + rr;
    if (rr) {  // successor of a sorted m-tuple over range(d), lexicographic
      int q = m - 1;
      while (q > 0 && I[q] == d - 1) --q;
      const int32_t v = I[q] + 1;
      for (int s = q; s < m; ++s) I[s] = v;
    }
    int64_t ps = 0;
    for (int q = 0; q < m; ++q) ps += binom_at(P.binom, P.rank, d - 1 + (m - 1 - q) - I[q], m - q);
    int p = 0;
    for (int64_t a = 0; a < d; ++a) {
      while (p < m && I[p] <= a) {
        ps += binom_at(P.binom, P.rank, d - 1 + (m - p) - I[p], m - p + 1) - binom_at(P.binom, P.rank, d - 1 + (m - 1 - p) - I[p], m - p);
        ++p;
      }
      tbl[a * nI + i] = (int32_t)(base - ps - binom_at(P.binom, P.rank, d - 1 + (m - p) - a, m - p + 1));
    }
  }
}

__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// CTA: one J, a strip of NT tiles of 64 rows I, 64 columns j.  The A-operand chunk S[a][i] is gathered from the packed
// row Tk[J][.] through the step's gather map (or, when the map would be too large, by ranking on the fly), W
// streams through shared memory; 8 warps, each a 16 x 32 block of the tile as 2 x 4 m8n8k4 accumulators.  The
// (tile, chunk) sequence of the strip is software-pipelined: the gathers of the next chunk are in flight (in
// registers) while the tensor pipe works on the current one.  Warps whose 32 columns lie entirely below max(J)
// (not sorted (k+1)-tuples) skip the arithmetic.
template <bool USE_TBL>
__global__ void __launch_bounds__(256, 2) mat_step_dmma_kernel(PlanView P, int k, int m, const double* __restrict__ Tk, const double* __restrict__ W,
                                                            double* __restrict__ Tn, int64_t nJ, int64_t nI, int64_t nI1,
                                                            const int32_t* __restrict__ tblI, int NT, int64_t jbase, int jlo, int jhi) {
  constexpr int TI = 64, TJ = 64, TK = 32, LD = 68;  // LD = 4 mod 16: conflict-free fragment loads
  constexpr int PER = TI * TK / 256;                 // gathered components (and W entries) per thread and chunk
  __shared__ double Ss[TK][LD];
  __shared__ double Ws[TK][LD];
  __shared__ int64_t orow[TJ];
  __shared__ int32_t Js[ST_MAX_RANK];
  const int64_t d = P.dim;
  const int64_t tilesI = (nI + TI - 1) / TI;
  const int64_t strips = (tilesI + NT - 1) / NT;
  const int64_t jidx = jbase + blockIdx.x / strips;  // (a range of rows and, at step 0, of columns: see mat_step_kernel)
  const int64_t t0 = (blockIdx.x % strips) * NT;
  const int64_t t1 = t0 + NT < tilesI ? t0 + NT : tilesI;
  const int64_t j0 = (int64_t)blockIdx.y * TJ;
  if (threadIdx.x == 0) flat_unrank_r(P, jidx, k, Js);
  __syncthreads();
  const int jlast = max(k ? Js[k - 1] : 0, jlo);
  if (j0 + TJ <= jlast || j0 >= jhi) return;  // every column of this tile is below max(J) or outside the column range
  if (threadIdx.x < TJ) {        // output row of every column: flat rank of the sorted (k+1)-tuple (J, j)
    const int64_t jj = j0 + threadIdx.x;
    int32_t Jn[ST_MAX_RANK];
    for (int q = 0; q < k; ++q) Jn[q] = Js[q];
    Jn[k] = (int32_t)jj;
    orow[threadIdx.x] = (jj < d && jj >= jlast && jj < jhi) ? flat_rank_r(P, Jn, k + 1) : -1;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp & 3, wc = warp >> 2;  // 16-row block, 32-column block
  const bool active = j0 + wc * 32 + 32 > jlast && j0 + wc * 32 < d && j0 + wc * 32 < jhi;
  const double* __restrict__ row = Tk + jidx * nI1;
  const int nchunk = (int)((d + TK - 1) / TK);
  double sreg[PER], wreg[PER];
  // request chunk c of tile t: this thread's PER components of S (row r = e % TI, a = a0 + e / TI) and entries of W
  auto fetch = [&](int64_t t, int c) {
    const int64_t i0 = t * TI, a0 = (int64_t)c * TK;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int e = threadIdx.x + u * 256;
      const int r = e % TI, aa = e / TI;
      const int64_t a = a0 + aa;
      double v = 0.0;
      if (i0 + r < nI && a < d) {
        int64_t pos;
        if (USE_TBL) {
          pos = (int64_t)__ldg(tblI + a * nI + i0 + r);
        } else {
          int32_t I[ST_MAX_RANK];
          flat_unrank_r(P, i0 + r, m, I);
          pos = merged_rank(P, I, m, a);
        }
        v = row[pos];
      }
      sreg[u] = v;
      const int cc = e % TJ;
      wreg[u] = (a < d && j0 + cc < d) ? W[a * d + j0 + cc] : 0.0;
    }
  };
  fetch(t0, 0);
  for (int64_t t = t0; t < t1; ++t) {
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      __syncthreads();  // the previous chunk's fragments have been read
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int e = threadIdx.x + u * 256;
        Ss[e / TI][e % TI] = sreg[u];
        Ws[e / TJ][e % TJ] = wreg[u];
      }
      __syncthreads();
      if (c + 1 < nchunk) fetch(t, c + 1);
      else if (t + 1 < t1) fetch(t + 1, 0);
      if (active) {
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
          double af[2], bf[4];
#pragma unroll
          for (int i = 0; i < 2; ++i) af[i] = Ss[kk + (lane & 3)][wr * 16 + i * 8 + (lane >> 2)];
#pragma unroll
          for (int j = 0; j < 4; ++j) bf[j] = Ws[kk + (lane & 3)][wc * 32 + j * 8 + (lane >> 2)];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j], af[i], bf[j]);
        }
      }
    }
    if (active) {
      // store: accumulator (i, j) holds rows wr*16 + i*8 + lane/4, columns wc*32 + j*8 + 2*(lane%4) + {0, 1}
      const int64_t i0 = t * TI;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t orw = orow[wc * 32 + j * 8 + 2 * (lane & 3) + h];
          if (orw < 0) continue;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int64_t ii = i0 + wr * 16 + i * 8 + (lane >> 2);
            if (ii < nI) Tn[orw * nI + ii] = acc[i][j][h];
          }
        }
    }
  }
}

// Last step of the chain (m == 0: no row indices left): Tn[(J, j)] = sum_a W[a][j] Tk[J][a] for j >= max(J), a plain
// [nJ x d] x [d x d] product whose rows are the sorted k-tuples J.  CTA: 64 consecutive J (their tuples by one
// unrank and 63 successor steps), 64 columns.  The output position of (J, j) is  c_J - (d - 1 - j)  with
// c_J = flat rank of (J, d - 1): consecutive j are consecutive output entries.
__global__ void __launch_bounds__(256, 2) mat_last_dmma_kernel(PlanView P, int k, const double* __restrict__ Tk, const double* __restrict__ W,
                                                               double* __restrict__ Tn, int64_t nJ, int64_t rbase) {
  // rows [rbase, nJ) of the sorted k-tuples (global row numbers; Tk / Tn indexed globally)
  constexpr int TI = 64, TJ = 64, TK = 32, LD = 68, LDS = 36;  // LDS = 4 mod 16: conflict-free stores (a fastest) and fragment loads
  __shared__ double Ss[TI][LDS];  // [row][a]: the rows of Tk are contiguous in a
  __shared__ double Ws[TK][LD];
  __shared__ int64_t cJ[TI];
  __shared__ int32_t jl[TI];
  const int64_t d = P.dim;
  const int64_t r0 = rbase + (int64_t)blockIdx.x * TI;
  const int64_t j0 = (int64_t)blockIdx.y * TJ;
  // every row's own J, in parallel (round 1 let one thread walk the 64 successors: ~5,000 clocks before the first load, the
  // whole cost of the step at BASELINE config 4)
  if (threadIdx.x < TI && r0 + threadIdx.x < nJ) {
    int32_t Jn[ST_MAX_RANK];
    flat_unrank_r(P, r0 + threadIdx.x, k, Jn);
    Jn[k] = (int32_t)(d - 1);
    cJ[threadIdx.x] = flat_rank_r(P, Jn, k + 1);
    jl[threadIdx.x] = k ? Jn[k - 1] : 0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp & 3, wc = warp >> 2;
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int64_t a0 = 0; a0 < d; a0 += TK) {
    __syncthreads();
    for (int e = threadIdx.x; e < TI * TK; e += 256) {
      const int aa = e % TK, r = e / TK;
      Ss[r][aa] = (r0 + r < nJ && a0 + aa < d) ? Tk[(r0 + r) * d + a0 + aa] : 0.0;
    }
    for (int e = threadIdx.x; e < TJ * TK; e += 256) {
      const int c = e % TJ, aa = e / TJ;
      Ws[aa][c] = (a0 + aa < d && j0 + c < d) ? W[(a0 + aa) * d + j0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; kk += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) af[i] = Ss[wr * 16 + i * 8 + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Ws[kk + (lane & 3)][wc * 32 + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j], af[i], bf[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = wr * 16 + i * 8 + (lane >> 2);
    if (r0 + r >= nJ) continue;
    const int64_t c = cJ[r];
    const int jlast = jl[r];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t jj = j0 + wc * 32 + j * 8 + 2 * (lane & 3) + h;
        if (jj < d && jj >= jlast) Tn[c - (d - 1 - jj)] = acc[i][j][h];
      }
  }
}

static const int64_t kMatTableMaxBytes = (int64_t)4 << 30;
int g_mat_dmma = 1;  // fp64 mode chain on the FP64 tensor pipe (0: the DFMA register-tile kernel; test hook)
// bytes of the gather map of step k (0: the step ranks on the fly)
static int64_t mat_table_bytes(const HostPlan* hp, int rank, int k);

static int grid_1d(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = 148 * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int64_t flat_size_host(const HostPlan* hp, int r) {
  return r == 0 ? 1 : hp->h_binom[(hp->dim + r - 1) * (hp->rank + 1) + r];
}

static int64_t mat_table_bytes(const HostPlan* hp, int rank, int k) {
  const int m = rank - k - 1;
  const int64_t nI = flat_size_host(hp, m), nI1 = flat_size_host(hp, m + 1), nJ = flat_size_host(hp, k);
  (void)nJ;
  if (m == 0 || nI1 >= 2147483647LL) return 0;  // the last step needs no map / does not fit int32
  const __int128 bytes = (__int128)nI * hp->dim * 4;
  return bytes <= kMatTableMaxBytes ? (int64_t)bytes : 0;
}

static double binom_double(int n, int k) {
  double v = 1.0;
  for (int i = 1; i <= k; ++i) v = v * (double)(n - k + i) / (double)i;
  return v;
}

int g_conv_rows = 1;  // row-walk layout converters (0: one unrank / rank round trip per component; test hook "conv_rows")

template <typename T>
static int permcls_to_flat(int rank, int64_t dim, const T* d_in, T* d_out, cudaStream_t stream) {
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  if (get_host_plan(rank, dim)->flat_overflow) { set_error("flat size does not fit int64"); return ST_ERR_OVERFLOW; }
  if (!d_in || !d_out) { set_error("null pointer"); return ST_ERR_INVALID; }
  if (g_conv_rows && rank >= 1 && rank <= 8 && dim <= 255) {
    const int64_t ntasks = (P.total + kConvSpan - 1) / kConvSpan;
    rowwalk_convert_kernel<T, true><<<(unsigned)std::min<int64_t>((ntasks + 7) / 8, (int64_t)sm_count() * 8), 256, 0, stream>>>(P, d_in, d_out, 0, P.total);
  } else {
    permcls_to_flat_kernel<T><<<grid_1d(P.flat_size, 256), 256, 0, stream>>>(P, d_in, d_out);
  }
  count_launch();
  return check_cuda(cudaGetLastError(), "permcls_to_flat_kernel");
}

template <typename T>
static int flat_to_permcls(int rank, int64_t dim, const T* d_in, T* d_out, int64_t begin, int64_t end, cudaStream_t stream) {
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  if (begin < 0 || end < begin || end > P.total) { set_error("range [%lld, %lld) outside [0, %lld]", (long long)begin, (long long)end, (long long)P.total); return ST_ERR_INVALID; }
  if (end == begin) return ST_OK;
  if (!d_in || !d_out) { set_error("null pointer"); return ST_ERR_INVALID; }
  if (g_conv_rows && rank >= 1 && rank <= 8 && dim <= 255) {
    const int64_t ntasks = (end - begin + kConvSpan - 1) / kConvSpan;
    rowwalk_convert_kernel<T, false><<<(unsigned)std::min<int64_t>((ntasks + 7) / 8, (int64_t)sm_count() * 8), 256, 0, stream>>>(P, d_in, d_out, begin, end);
  } else {
    flat_to_permcls_kernel<T><<<grid_1d(end - begin, 256), 256, 0, stream>>>(P, d_in, d_out, begin, end);
  }
  count_launch();
  return check_cuda(cudaGetLastError(), "flat_to_permcls_kernel");
}

int g_gram_umma = 1;   // fp32 tensordot Gram matrix on tcgen05 (0: the CUDA-core kernel; test hook)
int g_outer_fast = 1;  // compile-time-rank outer kernels (0: the run-time-rank kernel; test hook)
int g_outer_rows = 1;  // ... with the warp-uniform row walk (0: one unrank per component, outer_fast_kernel; test hook)

// launch the compile-time-rank kernel for (ra, rb) if there is one (ra >= rb; the symmetrized product commutes, so the
// caller swaps the operands otherwise); false: not instantiated / operands too large for 32-bit ranks
template <typename T, bool VEC>
static bool launch_outer_fast(const HostPlan* hp, const PlanView& P, int ra, int rb, const T* a, const T* b, T* out, const T* x, double* partials,
                              int64_t begin, int64_t end, int grid, cudaStream_t stream, int* used_grid = nullptr) {
  if (used_grid) *used_grid = grid;
  if (!g_outer_fast || ra < rb || rb < 1 || ra > 4) return false;
  if (flat_size_host(hp, ra) >= 2147483647LL || hp->dim >= 2147483647LL / 8) return false;
  if (g_outer_rows && hp->dim <= 255) {  // row walk: one warp task per kRowsSpan coordinates, as many CTAs as are resident
    const double inv = 1.0 / binom_double(ra + rb, ra);
    const int64_t ntasks = (end - begin + kRowsSpan - 1) / kRowsSpan;
#define ST_OUTER_ROWS_LAUNCH(KERNEL, THREADS)                                                                                        \
  {                                                                                                                                 \
    int occ = 0;                                                                                                                    \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KERNEL, THREADS, 0) != cudaSuccess || occ < 1) occ = 1;                   \
    const int wpb = THREADS / 32;                                                                                                   \
    const int g = (int)std::min<int64_t>(std::min<int64_t>((ntasks + wpb - 1) / wpb, (int64_t)sm_count() * occ), VEC ? grid : (1 << 30)); \
    KERNEL<<<std::max(g, 1), THREADS, 0, stream>>>(P, a, b, out, x, partials, begin, end, inv);                                       \
    if (used_grid) *used_grid = std::max(g, 1);                                                                                     \
    return true;                                                                                                                    \
  }
#define ST_OUTER_ROWS_CASE(RA, RB) \
  if (ra == RA && rb == RB) ST_OUTER_ROWS_LAUNCH((outer_rows_kernel<T, RA, RB, VEC>), 256)
    ST_OUTER_ROWS_CASE(1, 1) ST_OUTER_ROWS_CASE(2, 1) ST_OUTER_ROWS_CASE(2, 2) ST_OUTER_ROWS_CASE(3, 1) ST_OUTER_ROWS_CASE(3, 2)
    ST_OUTER_ROWS_CASE(3, 3) ST_OUTER_ROWS_CASE(4, 1) ST_OUTER_ROWS_CASE(4, 2) ST_OUTER_ROWS_CASE(4, 3) ST_OUTER_ROWS_CASE(4, 4)
#undef ST_OUTER_ROWS_CASE
#undef ST_OUTER_ROWS_LAUNCH
  }
  const size_t smem = (size_t)ra * hp->dim * sizeof(int32_t);
  if (smem > 40 * 1024) return false;
  const double inv = 1.0 / binom_double(ra + rb, ra);
#define ST_OUTER_CASE(RA, RB) \
  if (ra == RA && rb == RB) { outer_fast_kernel<T, RA, RB, VEC><<<grid, 256, smem, stream>>>(P, a, b, out, x, partials, begin, end, inv); return true; }
  ST_OUTER_CASE(1, 1) ST_OUTER_CASE(2, 1) ST_OUTER_CASE(2, 2) ST_OUTER_CASE(3, 1) ST_OUTER_CASE(3, 2) ST_OUTER_CASE(3, 3)
  ST_OUTER_CASE(4, 1) ST_OUTER_CASE(4, 2) ST_OUTER_CASE(4, 3) ST_OUTER_CASE(4, 4)
#undef ST_OUTER_CASE
  return false;
}

template <typename T>
static int outer(int ra, int rb, int64_t dim, const T* d_a_flat, const T* d_b_flat, T* d_out, int64_t begin, int64_t end, cudaStream_t stream,
                 int op = ST_OUTER_MULTIPLY) {
  if (ra < 0 || rb < 0 || ra + rb > ST_MAX_RANK) { set_error("ranks %d + %d exceed %d", ra, rb, ST_MAX_RANK); return ST_ERR_INVALID; }
  if (op != ST_OUTER_MULTIPLY && op != ST_OUTER_ADD && op != ST_OUTER_SUBTRACT) { set_error("unknown outer op %d", op); return ST_ERR_INVALID; }
  PlanView P;
  int rc = get_device_plan(ra + rb, dim, &P);
  if (rc) return rc;
  if (begin < 0 || end < begin || end > P.total) { set_error("range [%lld, %lld) outside [0, %lld]", (long long)begin, (long long)end, (long long)P.total); return ST_ERR_INVALID; }
  if (end == begin) return ST_OK;
  if (!d_a_flat || !d_b_flat || !d_out) { set_error("null pointer"); return ST_ERR_INVALID; }
  {
    const HostPlan* hp = get_host_plan(ra + rb, dim);
    const bool sw = ra < rb;  // A (x) B symmetrized == B (x) A symmetrized
    if (op != ST_OUTER_MULTIPLY || !hp || !launch_outer_fast<T, false>(hp, P, sw ? rb : ra, sw ? ra : rb, sw ? d_b_flat : d_a_flat, sw ? d_a_flat : d_b_flat, d_out, nullptr,
                                            nullptr, begin, end, grid_1d(end - begin, 256), stream))
      outer_kernel<T><<<grid_1d(end - begin, 256), 256, 0, stream>>>(P, ra, rb, d_a_flat, d_b_flat, d_out, begin, end,
                                                                      1.0 / binom_double(ra + rb, ra), op);
  }
  count_launch();
  return check_cuda(cudaGetLastError(), "outer_kernel");
}

static const int kOuterVecCtas = 148 * 8;

template <typename T>
static int outer_vec(int ra, int rb, int64_t dim, const T* d_a_flat, const T* d_b_flat, const T* d_x, T* d_out, double* d_ws, int64_t begin,
                     int64_t end, cudaStream_t stream) {
  if (ra < 0 || rb < 0 || ra + rb > ST_MAX_RANK) { set_error("ranks %d + %d exceed %d", ra, rb, ST_MAX_RANK); return ST_ERR_INVALID; }
  PlanView P;
  int rc = get_device_plan(ra + rb, dim, &P);
  if (rc) return rc;
  if (begin < 0 || end < begin || end > P.total) { set_error("range outside the packed tensor"); return ST_ERR_INVALID; }
  if (!d_a_flat || !d_b_flat || !d_out || !d_ws || (dim > 0 && !d_x)) { set_error("null pointer"); return ST_ERR_INVALID; }
  int grid = std::min(grid_1d(std::max<int64_t>(end - begin, 1), 256), kOuterVecCtas);
  {
    const HostPlan* hp = get_host_plan(ra + rb, dim);
    const bool sw = ra < rb;
    if (!hp || !launch_outer_fast<T, true>(hp, P, sw ? rb : ra, sw ? ra : rb, sw ? d_b_flat : d_a_flat, sw ? d_a_flat : d_b_flat, nullptr, d_x, d_ws,
                                           begin, end, grid, stream, &grid))
      outer_vec_kernel<T><<<grid, 256, 0, stream>>>(P, ra, rb, d_a_flat, d_b_flat, d_x, d_ws, begin, end, 1.0 / binom_double(ra + rb, ra));
  }
  if (sizeof(T) == 8) sum_partials_kernel<<<1, 256, 0, stream>>>(d_ws, grid, reinterpret_cast<double*>(d_out), nullptr);
  else sum_partials_kernel<<<1, 256, 0, stream>>>(d_ws, grid, nullptr, reinterpret_cast<float*>(d_out));
  count_launch(2);
  return check_cuda(cudaGetLastError(), "outer_vec_kernel");
}

// st_sym22.cu: non-materialising tcgen05 kernel for two free indices on each side (fp32)
namespace s22 {
int tensordot_sym22(int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out, int64_t begin, int64_t end, void* d_ws,
                    cudaStream_t stream);
int workspace_bytes(int k, int64_t dim, int64_t* out);
int tensordot_sym22_ranges(int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, int nranges, const int64_t* begins,
                           const int64_t* ends, float* const* d_outs, void* d_ws, cudaStream_t stream);
}
int g_sym22 = 1;             // fp32 tensordot with 2 + 2 free indices through the tiled kernel (0: always the materialised Gram matrix)
int64_t g_sym22_min_dim = 96;  // ... from this dimension on (tuning key "sym22_min_dim": tests lower it to reach the kernel at small sizes)
static bool use_sym22(int ra, int rb, int k, int64_t dim, int elem_size) {
  return g_sym22 && elem_size == 4 && k >= 1 && ra - k == 2 && rb - k == 2 && dim >= g_sym22_min_dim;
}

struct TdotShape {
  int na, nb, n, R;
  int64_t M, N, K;
};

static int tdot_shape(int ra, int rb, int k, int64_t dim, TdotShape* s) {
  if (ra < 0 || rb < 0 || k < 0 || k > ra || k > rb) { set_error("cannot contract %d axes of rank-%d and rank-%d tensors", k, ra, rb); return ST_ERR_INVALID; }
  s->na = ra - k;
  s->nb = rb - k;
  s->n = s->na + s->nb;
  s->R = std::max(std::max(ra, rb), s->n);
  if (s->R > ST_MAX_RANK) { set_error("rank %d exceeds %d", s->R, ST_MAX_RANK); return ST_ERR_INVALID; }
  const HostPlan* hp = get_host_plan(s->R, dim);
  if (!hp) return ST_ERR_INVALID;
  s->M = flat_size_host(hp, s->na);
  s->N = flat_size_host(hp, s->nb);
  s->K = flat_size_host(hp, k);
  return ST_OK;
}

template <typename T>
static int tensordot(int ra, int rb, int k, int64_t dim, const T* d_a_flat, const T* d_b_flat, T* d_out, int64_t begin, int64_t end,
                     void* d_ws, cudaStream_t stream) {
  TdotShape s;
  int rc = tdot_shape(ra, rb, k, dim, &s);
  if (rc) return rc;
  if (k == 0) return outer<T>(ra, rb, dim, d_a_flat, d_b_flat, d_out, begin, end, stream);
  PlanView P, Pn;
  rc = get_device_plan(s.R, dim, &P);  // operand gathers: binomials up to the largest rank involved
  if (rc) return rc;
  rc = get_device_plan(s.n, s.n ? dim : 1, &Pn);  // the output's permcls layout (rank 0: one component, dim 1)
  if (rc) return rc;
  if (begin < 0 || end < begin || end > Pn.total) { set_error("range outside the packed output"); return ST_ERR_INVALID; }
  if (end == begin) return ST_OK;
  if (!d_a_flat || !d_b_flat || !d_out || !d_ws) { set_error("null pointer"); return ST_ERR_INVALID; }
  if (use_sym22(ra, rb, k, dim, (int)sizeof(T)))
    return s22::tensordot_sym22(k, dim, reinterpret_cast<const float*>(d_a_flat), reinterpret_cast<const float*>(d_b_flat),
                                reinterpret_cast<float*>(d_out), begin, end, d_ws, stream);
  T* aex = reinterpret_cast<T*>(d_ws);
  T* bex = aex + s.M * s.K;
  T* G = bex + s.N * s.K;
  expand_kernel<T><<<grid_1d(s.M * s.K, 256), 256, 0, stream>>>(P, s.na, k, d_a_flat, aex, s.M, s.K, 1);
  expand_kernel<T><<<grid_1d(s.N * s.K, 256), 256, 0, stream>>>(P, s.nb, k, d_b_flat, bex, s.N, s.K, 0);
  const dim3 grid((unsigned)((s.N + 63) / 64), (unsigned)((s.M + 63) / 64));
  if (grid.y > 65535) { set_error("Gram matrix with %lld rows needs the tiled (non-materialising) kernel", (long long)s.M); return ST_ERR_UNSUPPORTED; }
  if (sizeof(T) == 4 && g_gram_umma) {
    rc = set_max_dynamic_smem(reinterpret_cast<const void*>(gram_umma_kernel), (int)umma::SMEM_BYTES);
    if (rc) return rc;
    const dim3 ugrid((unsigned)((s.N + umma::BN - 1) / umma::BN), (unsigned)((s.M + umma::BM - 1) / umma::BM));
    if (ugrid.y > 65535) { set_error("Gram matrix with %lld rows needs the tiled (non-materialising) kernel", (long long)s.M); return ST_ERR_UNSUPPORTED; }
    gram_umma_kernel<<<ugrid, 128, umma::SMEM_BYTES, stream>>>(reinterpret_cast<const float*>(aex), reinterpret_cast<const float*>(bex),
                                                              reinterpret_cast<float*>(G), s.M, s.N, s.K);
  } else {
    gemm_nt_kernel<T><<<grid, 256, 0, stream>>>(aex, bex, G, s.M, s.N, s.K);
  }
  // the gather runs on the plan of the output rank; sub-tuple ranks only need binomials up to rank n there
  {
    const double inv = 1.0 / binom_double(s.n, s.na);
    const int tm = std::max(s.na, s.nb);
    const size_t fsm = (size_t)tm * dim * sizeof(int32_t);
    const int gg = grid_1d(end - begin, 256);
    bool done = false;
    if (g_outer_fast && s.na >= 1 && s.nb >= 1 && tm <= 3 && fsm <= 40 * 1024 && s.M < 2147483647LL && s.N < 2147483647LL) {
#define ST_GRAM_CASE(NA, NB) \
  if (s.na == NA && s.nb == NB) { gram_gather_fast_kernel<T, NA, NB><<<gg, 256, fsm, stream>>>(Pn, G, s.N, d_out, begin, end, inv); done = true; }
      ST_GRAM_CASE(1, 1) ST_GRAM_CASE(1, 2) ST_GRAM_CASE(2, 1) ST_GRAM_CASE(2, 2) ST_GRAM_CASE(1, 3) ST_GRAM_CASE(3, 1)
      ST_GRAM_CASE(2, 3) ST_GRAM_CASE(3, 2) ST_GRAM_CASE(3, 3)
#undef ST_GRAM_CASE
    }
    if (!done) gram_gather_kernel<T><<<gg, 256, 0, stream>>>(Pn, s.na, s.nb, G, s.N, d_out, begin, end, inv);
  }
  count_launch(4);
  return check_cuda(cudaGetLastError(), "tensordot kernels");
}

// rows of the sorted k-tuples over range(d) whose FIRST element is below v (lexicographic = flat order)
static int64_t rows_below(const HostPlan* hp, int k, int64_t v) {
  if (k == 0) return v > 0 ? 1 : 0;
  const int64_t d = hp->dim;
  if (v >= d) return flat_size_host(hp, k);
  if (v <= 0) return 0;
  return flat_size_host(hp, k) - hp->h_binom[(d - v + k - 1) * (hp->rank + 1) + k];
}

// elements of one ping-pong intermediate for the output modes j1 in [jlo, jhi)
static int64_t mat_range_elems(const HostPlan* hp, int rank, int64_t jlo, int64_t jhi) {
  int64_t maxT = 0;
  for (int k = 1; k <= rank; ++k) maxT = std::max(maxT, (rows_below(hp, k, jhi) - rows_below(hp, k, jlo)) * flat_size_host(hp, rank - k));
  return maxT;
}

// st_mat.cu: persistent producer / consumer pipeline for one step (fp64, dim <= 64, gather map present)
namespace matpipe {
bool launch_step(const PlanView& P, int k, int m, const double* Tk, const double* W, double* Tn, int64_t nJ, int64_t nI, int64_t nI1,
                 const int32_t* tbl, int64_t rlo, int clo, int chi, cudaStream_t stream);
}
int g_mat_pipe = 1;  // (0: the first DMMA kernel for every step; test hook "mat_pipe")
// steps with at most this many rows J rank their gathers in the producers instead of building a map ("mat_onfly_rows").  Off by
// default: at BASELINE config 4 step 0 took 7.8 ms that way against 2.9 (map) + 3.3 ms -- the per-tile unrank of every column
// makes the producers the bottleneck.  The path also serves steps whose map would not fit the workspace.
int64_t g_mat_onfly_rows = 0;

// C = (W^T)^{(x) r} . A restricted to the output components whose FIRST (smallest) mode j1 lies in [jlo, jhi): the flat
// range [rows_below(r, jlo), rows_below(r, jhi)) of the output, written to d_out_slice (which starts at that position).
// Every intermediate T_k of the chain shards with it (rows J whose first element is in the range), so a GPU of a
// partition by j1 holds and computes only its slice of the chain: the multi-GPU scheme of SURVEY.md 8e, no collective.
template <typename T>
static int contract_mat(int rank, int64_t dim, const T* d_a_flat, const T* d_W, T* d_out_slice, void* d_ws, cudaStream_t stream,
                        int64_t jlo = 0, int64_t jhi = -1) {
  if (rank < 0 || rank > ST_MAX_RANK) { set_error("rank %d outside [0, %d]", rank, ST_MAX_RANK); return ST_ERR_INVALID; }
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  const HostPlan* hp = get_host_plan(rank, dim);
  if (jhi < 0) jhi = dim;
  if (jlo < 0 || jhi < jlo || jhi > dim) { set_error("mode range [%lld, %lld) outside [0, %lld]", (long long)jlo, (long long)jhi, (long long)dim); return ST_ERR_INVALID; }
  if (jhi == jlo && rank > 0 && dim > 0) return ST_OK;  // empty slice
  if (!d_a_flat || !d_out_slice || (rank > 0 && (!d_W || !d_ws))) { set_error("null pointer"); return ST_ERR_INVALID; }
  if (rank == 0 || dim == 0) return check_cuda(cudaMemcpyAsync(d_out_slice, d_a_flat, sizeof(T) * (size_t)P.flat_size, cudaMemcpyDeviceToDevice, stream), "cudaMemcpyAsync");
  const bool whole = jlo == 0 && jhi == dim;
  int64_t maxT = 0;
  if (whole) { for (int k = 0; k <= rank; ++k) maxT = std::max(maxT, flat_size_host(hp, k) * flat_size_host(hp, rank - k)); }
  else maxT = mat_range_elems(hp, rank, jlo, jhi);
  T* buf[2] = {reinterpret_cast<T*>(d_ws), reinterpret_cast<T*>(d_ws) + maxT};
  const T* src = d_a_flat;  // global row indexing: row 0 of step 0 is the whole input
  for (int k = 0; k < rank; ++k) {
    const int m = rank - k - 1;
    const int64_t nI = flat_size_host(hp, m), nI1 = flat_size_host(hp, m + 1);
    const int64_t rlo = k == 0 ? 0 : rows_below(hp, k, jlo), rhi = k == 0 ? 1 : rows_below(hp, k, jhi);  // rows J of this step
    const int64_t olo = rows_below(hp, k + 1, jlo);                                          // first output row
    const int64_t nJ = rhi - rlo;
    const int clo = k == 0 ? (int)jlo : 0, chi = k == 0 ? (int)jhi : (int)dim;               // columns j (step 0 only: the partition)
    T* dst_slice = (k == rank - 1) ? d_out_slice : buf[k & 1];
    T* dst = dst_slice - olo * nI;  // indexed by global output rows; only rows of the slice are touched
    const int64_t tilesI = (nI + 63) / 64;
    const int64_t gx = nJ * tilesI;
    if (gx > 2147483647LL) { set_error("mode-chain step %d needs %lld CTAs", k, (long long)gx); return ST_ERR_UNSUPPORTED; }
    const dim3 grid((unsigned)gx, (unsigned)((dim + 63) / 64));
    if (sizeof(T) == 8 && g_mat_dmma) {
      const int64_t tb = mat_table_bytes(hp, rank, k);
      int32_t* tbl = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(d_ws) + 2 * (size_t)maxT * sizeof(T));
      // strips of NT row tiles per CTA (software-pipelined): as long as there are enough CTAs to fill the machine
      int NT = 1;
      while (NT < 8 && nJ * ((tilesI + 2 * NT - 1) / (2 * NT)) >= 148 * 8) NT *= 2;
      const dim3 sgrid((unsigned)(nJ * ((tilesI + NT - 1) / NT)), (unsigned)((dim + 63) / 64));
      if (m == 0) {
        const dim3 lgrid((unsigned)((nJ + 63) / 64), (unsigned)((dim + 63) / 64));
        mat_last_dmma_kernel<<<lgrid, 256, 0, stream>>>(P, k, reinterpret_cast<const double*>(src), reinterpret_cast<const double*>(d_W),
                                                       reinterpret_cast<double*>(dst), rhi, rlo);
      } else if (g_mat_pipe && (nJ <= g_mat_onfly_rows || tb == 0) && dim <= 64 &&
                 matpipe::launch_step(P, k, m, reinterpret_cast<const double*>(src), reinterpret_cast<const double*>(d_W), reinterpret_cast<double*>(dst),
                                      nJ, nI, nI1, nullptr, rlo, clo, chi, stream)) {
        // (few rows J -- step 0: the gather map would be used about once; the producers rank the gathers themselves)
      } else if (tb > 0) {
        mat_index_kernel<<<(unsigned)((nI + 256 * kIdxRun - 1) / (256 * kIdxRun)), 256, 0, stream>>>(P, m, nI, tbl);
        count_launch();
        if (!g_mat_pipe || !matpipe::launch_step(P, k, m, reinterpret_cast<const double*>(src), reinterpret_cast<const double*>(d_W),
                                                 reinterpret_cast<double*>(dst), nJ, nI, nI1, tbl, rlo, clo, chi, stream))
          mat_step_dmma_kernel<true><<<sgrid, 256, 0, stream>>>(P, k, m, reinterpret_cast<const double*>(src), reinterpret_cast<const double*>(d_W),
                                                                reinterpret_cast<double*>(dst), nJ, nI, nI1, tbl, NT, rlo, clo, chi);
      } else {
        mat_step_dmma_kernel<false><<<sgrid, 256, 0, stream>>>(P, k, m, reinterpret_cast<const double*>(src), reinterpret_cast<const double*>(d_W),
                                                               reinterpret_cast<double*>(dst), nJ, nI, nI1, nullptr, NT, rlo, clo, chi);
      }
    } else {
      mat_step_kernel<T><<<grid, 256, 0, stream>>>(P, k, m, src, d_W, dst, nJ, nI, nI1, rlo, clo, chi);
    }
    count_launch();
    src = dst;  // (global row indexing again)
  }
  return check_cuda(cudaGetLastError(), "mat_step_kernel");
}

}  // namespace st

using namespace st;

extern "C" {

int st_debug_rowwalk_device(int rank, int64_t dim, int64_t begin, int64_t end, int64_t span, int32_t* d_idx, void* stream) {
  return st_debug_rowwalk_device2(rank, dim, begin, end, span, d_idx, nullptr, stream);
}
int st_debug_rowwalk_device2(int rank, int64_t dim, int64_t begin, int64_t end, int64_t span, int32_t* d_idx, int32_t* d_dbg, void* stream) {
  PlanView P;
  int rc = get_device_plan(rank, dim, &P);
  if (rc) return rc;
  if (rank < 1 || rank > 8 || dim > 255 || span < 32 || span % 32 || begin < 0 || end < begin || end > P.total || !d_idx) { set_error("bad arguments"); return ST_ERR_INVALID; }
  if (end == begin) return ST_OK;
  const int64_t ntasks = (end - begin + span - 1) / span;
  rowwalk_debug_kernel<<<(unsigned)std::min<int64_t>((ntasks + 7) / 8, 148 * 8), 256, 0, (cudaStream_t)stream>>>(P, begin, end, (int)span, d_idx, d_dbg);
  return check_cuda(cudaGetLastError(), "rowwalk_debug_kernel");
}

int st_permcls_to_flat_f64(int rank, int64_t dim, const double* d_permcls, double* d_flat, void* stream) {
  return permcls_to_flat<double>(rank, dim, d_permcls, d_flat, (cudaStream_t)stream);
}
int st_permcls_to_flat_f32(int rank, int64_t dim, const float* d_permcls, float* d_flat, void* stream) {
  return permcls_to_flat<float>(rank, dim, d_permcls, d_flat, (cudaStream_t)stream);
}
int st_flat_to_permcls_f64(int rank, int64_t dim, const double* d_flat, double* d_permcls, int64_t begin, int64_t end, void* stream) {
  return flat_to_permcls<double>(rank, dim, d_flat, d_permcls, begin, end, (cudaStream_t)stream);
}
int st_flat_to_permcls_f32(int rank, int64_t dim, const float* d_flat, float* d_permcls, int64_t begin, int64_t end, void* stream) {
  return flat_to_permcls<float>(rank, dim, d_flat, d_permcls, begin, end, (cudaStream_t)stream);
}

int st_outer_f64(int ra, int rb, int64_t dim, const double* d_a_flat, const double* d_b_flat, double* d_out, int64_t begin, int64_t end,
                 void* stream) {
  return outer<double>(ra, rb, dim, d_a_flat, d_b_flat, d_out, begin, end, (cudaStream_t)stream);
}
int st_outer_f32(int ra, int rb, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out, int64_t begin, int64_t end,
                 void* stream) {
  return outer<float>(ra, rb, dim, d_a_flat, d_b_flat, d_out, begin, end, (cudaStream_t)stream);
}

int st_outer_op_f64(int op, int ra, int rb, int64_t dim, const double* d_a_flat, const double* d_b_flat, double* d_out, int64_t begin, int64_t end,
                    void* stream) {
  return outer<double>(ra, rb, dim, d_a_flat, d_b_flat, d_out, begin, end, (cudaStream_t)stream, op);
}
int st_outer_op_f32(int op, int ra, int rb, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out, int64_t begin, int64_t end,
                    void* stream) {
  return outer<float>(ra, rb, dim, d_a_flat, d_b_flat, d_out, begin, end, (cudaStream_t)stream, op);
}

int64_t st_outer_vec_workspace_bytes(void) { return (int64_t)sizeof(double) * kOuterVecCtas; }
int st_outer_vec_f64(int ra, int rb, int64_t dim, const double* d_a_flat, const double* d_b_flat, const double* d_x, double* d_out,
                     void* d_workspace, int64_t begin, int64_t end, void* stream) {
  return outer_vec<double>(ra, rb, dim, d_a_flat, d_b_flat, d_x, d_out, reinterpret_cast<double*>(d_workspace), begin, end, (cudaStream_t)stream);
}
int st_outer_vec_f32(int ra, int rb, int64_t dim, const float* d_a_flat, const float* d_b_flat, const float* d_x, float* d_out,
                     void* d_workspace, int64_t begin, int64_t end, void* stream) {
  return outer_vec<float>(ra, rb, dim, d_a_flat, d_b_flat, d_x, d_out, reinterpret_cast<double*>(d_workspace), begin, end, (cudaStream_t)stream);
}

int st_tensordot_workspace_bytes(int ra, int rb, int k, int64_t dim, int elem_size, int64_t* out_bytes) {
  TdotShape s;
  int rc = tdot_shape(ra, rb, k, dim, &s);
  if (rc) return rc;
  if (!out_bytes) { set_error("null pointer"); return ST_ERR_INVALID; }
  if (use_sym22(ra, rb, k, dim, elem_size)) return s22::workspace_bytes(k, dim, out_bytes);
  const __int128 elems = k == 0 ? 0 : (__int128)s.M * s.K + (__int128)s.N * s.K + (__int128)s.M * s.N;
  if (elems * elem_size > (__int128)INT64_MAX) { set_error("workspace does not fit int64"); return ST_ERR_OVERFLOW; }
  *out_bytes = (int64_t)elems * elem_size;
  return ST_OK;
}
int st_tensordot_is_tiled(int ra, int rb, int k, int64_t dim, int elem_size) { return use_sym22(ra, rb, k, dim, elem_size) ? 1 : 0; }
int st_tensordot_f64(int ra, int rb, int k, int64_t dim, const double* d_a_flat, const double* d_b_flat, double* d_out, int64_t begin,
                     int64_t end, void* d_workspace, void* stream) {
  return tensordot<double>(ra, rb, k, dim, d_a_flat, d_b_flat, d_out, begin, end, d_workspace, (cudaStream_t)stream);
}
int st_tensordot_f32(int ra, int rb, int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, float* d_out, int64_t begin,
                     int64_t end, void* d_workspace, void* stream) {
  return tensordot<float>(ra, rb, k, dim, d_a_flat, d_b_flat, d_out, begin, end, d_workspace, (cudaStream_t)stream);
}

int st_tensordot_ranges_f32(int ra, int rb, int k, int64_t dim, const float* d_a_flat, const float* d_b_flat, int nranges, const int64_t* begins,
                            const int64_t* ends, float* const* d_outs, void* d_workspace, void* stream) {
  if (nranges < 1 || !begins || !ends || !d_outs) { set_error("null pointer / no range"); return ST_ERR_INVALID; }
  if (!use_sym22(ra, rb, k, dim, 4)) {  // shapes without the tiled kernel: one call per range
    for (int q = 0; q < nranges; ++q) {
      const int rc = tensordot<float>(ra, rb, k, dim, d_a_flat, d_b_flat, d_outs[q], begins[q], ends[q], d_workspace, (cudaStream_t)stream);
      if (rc) return rc;
    }
    return ST_OK;
  }
  PlanView Pn;
  int rc = get_device_plan(4, dim, &Pn);
  if (rc) return rc;
  if (!d_a_flat || !d_b_flat || !d_workspace) { set_error("null pointer"); return ST_ERR_INVALID; }
  for (int q = 0; q < nranges; ++q) {
    if (begins[q] < 0 || ends[q] < begins[q] || ends[q] > Pn.total || (ends[q] > begins[q] && !d_outs[q])) { set_error("range %d outside the packed output / null buffer", q); return ST_ERR_INVALID; }
    for (int p = 0; p < q; ++p)
      if (begins[q] < ends[p] && begins[p] < ends[q]) { set_error("output ranges %d and %d overlap", p, q); return ST_ERR_INVALID; }
  }
  return s22::tensordot_sym22_ranges(k, dim, d_a_flat, d_b_flat, nranges, begins, ends, d_outs, d_workspace, (cudaStream_t)stream);
}

int st_contract_mat_workspace_bytes(int rank, int64_t dim, int elem_size, int64_t* out_bytes) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return ST_ERR_INVALID;
  if (!out_bytes) { set_error("null pointer"); return ST_ERR_INVALID; }
  __int128 maxT = 0;
  for (int k = 0; k <= rank; ++k) maxT = std::max(maxT, (__int128)flat_size_host(hp, k) * flat_size_host(hp, rank - k));
  int64_t tbl = 0;
  if (elem_size == 8)
    for (int k = 0; k < rank; ++k) tbl = std::max(tbl, mat_table_bytes(hp, rank, k));
  if (maxT * 2 * elem_size + tbl > (__int128)INT64_MAX) { set_error("workspace does not fit int64"); return ST_ERR_OVERFLOW; }
  *out_bytes = (int64_t)(maxT * 2 * elem_size) + tbl;
  return ST_OK;
}
int st_contract_mat_range_bounds(int rank, int64_t dim, int64_t jlo, int64_t jhi, int64_t* flat_begin, int64_t* flat_end) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return ST_ERR_INVALID;
  if (jlo < 0 || jhi < jlo || jhi > dim) { set_error("mode range [%lld, %lld) outside [0, %lld]", (long long)jlo, (long long)jhi, (long long)dim); return ST_ERR_INVALID; }
  if (flat_begin) *flat_begin = rows_below(hp, rank, jlo);
  if (flat_end) *flat_end = rows_below(hp, rank, jhi);
  return ST_OK;
}
int st_contract_mat_range_workspace_bytes(int rank, int64_t dim, int64_t jlo, int64_t jhi, int elem_size, int64_t* out_bytes) {
  const HostPlan* hp = get_host_plan(rank, dim);
  if (!hp) return ST_ERR_INVALID;
  if (!out_bytes) { set_error("null pointer"); return ST_ERR_INVALID; }
  if (jlo < 0 || jhi < jlo || jhi > dim) { set_error("mode range [%lld, %lld) outside [0, %lld]", (long long)jlo, (long long)jhi, (long long)dim); return ST_ERR_INVALID; }
  if (jlo == 0 && jhi == dim) return st_contract_mat_workspace_bytes(rank, dim, elem_size, out_bytes);
  int64_t tbl = 0;
  if (elem_size == 8)
    for (int k = 0; k < rank; ++k) tbl = std::max(tbl, mat_table_bytes(hp, rank, k));
  *out_bytes = mat_range_elems(hp, rank, jlo, jhi) * 2 * elem_size + tbl;
  return ST_OK;
}
int st_contract_mat_range_f64(int rank, int64_t dim, const double* d_a_flat, const double* d_W, double* d_out_slice, int64_t jlo, int64_t jhi,
                              void* d_workspace, void* stream) {
  return contract_mat<double>(rank, dim, d_a_flat, d_W, d_out_slice, d_workspace, (cudaStream_t)stream, jlo, jhi);
}
int st_contract_mat_range_f32(int rank, int64_t dim, const float* d_a_flat, const float* d_W, float* d_out_slice, int64_t jlo, int64_t jhi,
                              void* d_workspace, void* stream) {
  return contract_mat<float>(rank, dim, d_a_flat, d_W, d_out_slice, d_workspace, (cudaStream_t)stream, jlo, jhi);
}
int st_contract_mat_f64(int rank, int64_t dim, const double* d_a_flat, const double* d_W, double* d_out_flat, void* d_workspace, void* stream) {
  return contract_mat<double>(rank, dim, d_a_flat, d_W, d_out_flat, d_workspace, (cudaStream_t)stream);
}
int st_contract_mat_f32(int rank, int64_t dim, const float* d_a_flat, const float* d_W, float* d_out_flat, void* d_workspace, void* stream) {
  return contract_mat<float>(rank, dim, d_a_flat, d_W, d_out_flat, d_workspace, (cudaStream_t)stream);
}

}  // extern "C"
