// Probe: C[M x N] = A[M x K] * B[N x K]^T in fp32 accuracy on the 5th-generation tensor cores: tcgen05.mma kind::tf32
// with the 3xTF32 split (hi/lo), accumulator in TMEM, operands staged to shared memory by the CTA's threads in the
// canonical K-major no-swizzle layout.  One CTA per 128 x 128 tile; single-buffered (correctness and a first rate).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int BM = 128, BN = 128, BK = 32;          // tile; BK tf32 elements = 128 bytes per row
constexpr uint32_t LBO = 128, SBO = (BK / 4) * 128;  // bytes: next 16-byte K chunk, next 8-row group
constexpr int TILE_BYTES = (BM / 8) * SBO;           // 16 KB

// shared memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(LBO >> 4) << 16;
  d |= (uint64_t)(SBO >> 4) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// element (r, k) of a staged tile
__device__ __forceinline__ uint32_t tile_off(int r, int k) { return (uint32_t)(r >> 3) * SBO + (uint32_t)(k >> 2) * LBO + (uint32_t)(r & 7) * 16 + (uint32_t)(k & 3) * 4; }

__global__ void __launch_bounds__(128, 1) umma_gemm_nt(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
                                                       int split, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tiles = smem;  // A hi, A lo, B hi, B lo
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc(BM, BN);
  uint32_t phase = 0;
  bool ok = true;
  for (int k0 = 0; k0 < K; k0 += BK) {
    // stage: thread t handles 16-byte chunks (row r, chunk kc); hi = tf32-truncated value, lo = remainder
    for (int e = tid; e < BM * (BK / 4); e += 128) {
      const int r = e / (BK / 4), kc = e % (BK / 4);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (m0 + r < M && k0 + kc * 4 < K) a = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * K + k0 + kc * 4);
      if (n0 + r < N && k0 + kc * 4 < K) b = *reinterpret_cast<const float4*>(B + (size_t)(n0 + r) * K + k0 + kc * 4);
      auto hi = [](float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); };
      const float4 ah = make_float4(hi(a.x), hi(a.y), hi(a.z), hi(a.w)), bh = make_float4(hi(b.x), hi(b.y), hi(b.z), hi(b.w));
      const float4 al = make_float4(a.x - ah.x, a.y - ah.y, a.z - ah.z, a.w - ah.w), bl = make_float4(b.x - bh.x, b.y - bh.y, b.z - bh.z, b.w - bh.w);
      const uint32_t off = tile_off(r, kc * 4);
      *reinterpret_cast<float4*>(tiles + 0 * TILE_BYTES + off) = ah;
      *reinterpret_cast<float4*>(tiles + 1 * TILE_BYTES + off) = al;
      *reinterpret_cast<float4*>(tiles + 2 * TILE_BYTES + off) = bh;
      *reinterpret_cast<float4*>(tiles + 3 * TILE_BYTES + off) = bl;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async-proxy (tensor core) reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = smem_u32(tiles);
      for (int ks = 0; ks < BK / 8; ++ks) {  // one MMA = 8 tf32 along K = two 16-byte chunks
        const uint32_t koff = ks * 2 * LBO;
        const uint64_t dAh = make_desc(base + 0 * TILE_BYTES + koff), dAl = make_desc(base + 1 * TILE_BYTES + koff);
        const uint64_t dBh = make_desc(base + 2 * TILE_BYTES + koff), dBl = make_desc(base + 3 * TILE_BYTES + koff);
        umma_tf32(tmem_base, dAh, dBh, idesc, (k0 > 0 || ks > 0) ? 1u : 0u);
        if (split) {
          umma_tf32(tmem_base, dAh, dBl, idesc, 1u);
          umma_tf32(tmem_base, dAl, dBh, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // the tiles are overwritten by the next stage: wait for the MMAs of this one
    if (!mbar_wait_bounded(smem_u32(&bar), phase)) { ok = false; break; }
    phase ^= 1;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
    // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (accumulator rows), 8 columns per load
    const int row = m0 + warp * 32 + lane;
    for (int c0 = 0; c0 < BN; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < M)
        for (int j = 0; j < 8; ++j)
          if (n0 + c0 + j < N) C[(size_t)row * N + n0 + c0 + j] = __uint_as_float(v[j]);
    }
  } else if (tid == 0) {
    atomicExch(status, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

// ---- version 2: operands split once in global memory (hi / lo arrays), staged with 16-byte cp.async straight into the
// canonical layout, two stages: the copies of k-block i overlap the MMAs of k-block i - 1 --------------------------
__global__ void split_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = in[i], h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
constexpr int NST = 3, CHAIN = 256;
__global__ void __launch_bounds__(128, 1) umma_gemm_nt_v2(const float* __restrict__ Ah, const float* __restrict__ Al, const float* __restrict__ Bh,
                                                          const float* __restrict__ Bl, float* __restrict__ C, int M, int N, int K, int* status, int repeat) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[NST];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc(BM, BN);
  const uint32_t sbase = smem_u32(smem);
  const int nkb = (K + BK - 1) / BK;
  float acc[BN];
#pragma unroll
  for (int j = 0; j < BN; ++j) acc[j] = 0.f;
  int waited[NST] = {0, 0, 0};  // completed MMA batches consumed per stage (one commit per k-block on its stage)
  bool ok = true;
  auto wait_kb = [&](int kb) {  // until the MMAs of k-block kb (and all before it) have completed
    const int st = kb % NST, target = kb / NST + 1;
    while (ok && waited[st] < target) {
      if (!mbar_wait_bounded(smem_u32(&bars[st]), (uint32_t)(waited[st] & 1))) ok = false;
      ++waited[st];
    }
  };
  // this thread copies the 16-byte chunk kc = tid % 8 of the rows r = tid / 8 + 16 i (i = 0 .. 7) of all four tiles:
  // constant strides in shared memory (2 SBO) and in global memory (16 K floats), nothing to recompute per chunk
  const int kc_t = tid & 7, r_t = tid >> 3;
  const uint32_t off_t = tile_off(r_t, kc_t * 4);
  auto stage_copy = [&](int kb) {
    const uint32_t base = sbase + (kb % NST) * 4 * TILE_BYTES + off_t;
    const int kk = kb * BK + kc_t * 4;
    const bool vk = kk < K;
    const float* pah = Ah + (size_t)(m0 + r_t) * K + kk;
    const float* pal = Al + (size_t)(m0 + r_t) * K + kk;
    const float* pbh = Bh + (size_t)(n0 + r_t) * K + kk;
    const float* pbl = Bl + (size_t)(n0 + r_t) * K + kk;
    const size_t gstep = (size_t)16 * K;
#pragma unroll
    for (int i = 0; i < BM / 16; ++i) {
      const bool va = vk && m0 + r_t + 16 * i < M, vb = vk && n0 + r_t + 16 * i < N;
      const uint32_t d = base + (uint32_t)i * 2 * SBO;
      cp16(d + 0 * TILE_BYTES, va ? pah + i * gstep : Ah, va ? 16u : 0u);
      cp16(d + 1 * TILE_BYTES, va ? pal + i * gstep : Al, va ? 16u : 0u);
      cp16(d + 2 * TILE_BYTES, vb ? pbh + i * gstep : Bh, vb ? 16u : 0u);
      cp16(d + 3 * TILE_BYTES, vb ? pbl + i * gstep : Bl, vb ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
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
  stage_copy(0);
  if (nkb > 1) stage_copy(1);
  int in_chain = 0;
  for (int kb = 0; kb < nkb && ok; ++kb) {
    const int st = kb % NST;
    if (kb + 2 < nkb) {
      if (kb >= 1) wait_kb(kb - 1);  // stage (kb + 2) % 3 is free once the MMAs of k-block kb - 1 are done
      if (repeat != -1) stage_copy(kb + 2); else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 2;" ::: "memory");  // k-block kb has landed (this thread's copies)
    } else if (kb + 1 < nkb) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = sbase + st * 4 * TILE_BYTES;
      // descriptors: start address field = (address >> 4); one k-step = 2 LBO = 256 bytes = 16 units, one tile = TILE_BYTES / 16 units
      const uint64_t d0 = make_desc(base);
      constexpr uint64_t TU = TILE_BYTES >> 4, KU = (2 * LBO) >> 4;
      for (int rep = 0; rep < (repeat < 0 ? (repeat == -2 ? 0 : 1) : repeat); ++rep) {
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t dAh = d0 + ks * KU, dAl = dAh + TU, dBh = dAh + 2 * TU, dBl = dAh + 3 * TU;
          umma_tf32(tmem_base, dAh, dBh, idesc, (in_chain > 0 || ks > 0 || rep > 0) ? 1u : 0u);
          umma_tf32(tmem_base, dAh, dBl, idesc, 1u);
          umma_tf32(tmem_base, dAl, dBh, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[st])) : "memory");
    }
    in_chain += BK;
    if (in_chain >= CHAIN || kb + 1 == nkb) {  // end of a chain: its MMAs done, add the TMEM accumulator to the registers
      wait_kb(kb);
      if (!ok) break;
      drain();
      in_chain = 0;
      __syncthreads();
    }
  }
  const int row = m0 + warp * 32 + lane;
  if (row < M) {
#pragma unroll
    for (int j = 0; j < BN; ++j)
      if (n0 + j < N) C[(size_t)row * N + n0 + j] = ok ? acc[j] : __uint_as_float(0x7fc00000u);
  }
  if (!ok && tid == 0) atomicExch(status, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 256, N = argc > 2 ? atoi(argv[2]) : 384, K = argc > 3 ? atoi(argv[3]) : 160;
  std::vector<float> hA((size_t)M * K), hB((size_t)N * K), hC((size_t)M * N);
  srand(1);
  for (auto& v : hA) v = (float)rand() / RAND_MAX + 0.5f;
  for (auto& v : hB) v = (float)rand() / RAND_MAX - 0.25f;
  float *dA, *dB, *dC;
  int* dS;
  cudaMalloc(&dA, hA.size() * 4);
  cudaMalloc(&dB, hB.size() * 4);
  cudaMalloc(&dC, hC.size() * 4);
  cudaMalloc(&dS, 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 4 * TILE_BYTES + 1024;
  cudaFuncSetAttribute(umma_gemm_nt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  for (int split = 0; split < 2; ++split) {
    cudaMemset(dC, 0, hC.size() * 4);
    cudaMemset(dS, 0, 4);
    umma_gemm_nt<<<grid, 128, smem>>>(dA, dB, dC, M, N, K, split, dS);
    cudaError_t err = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double maxrel = 0, maxabs = 0;
    for (int i = 0; i < M; i += 7)
      for (int j = 0; j < N; j += 5) {
        double ref = 0, sc = 0;
        for (int k = 0; k < K; ++k) { ref += (double)hA[(size_t)i * K + k] * hB[(size_t)j * K + k]; sc += fabs((double)hA[(size_t)i * K + k] * hB[(size_t)j * K + k]); }
        const double d = fabs(hC[(size_t)i * N + j] - ref);
        if (d / sc > maxrel) maxrel = d / sc;
        if (d > maxabs) maxabs = d;
      }
    printf("split=%d  err=%s  timeout=%d  max |diff| / sum|terms| = %.3e  (max abs %.3e)\n", split, cudaGetErrorString(err), st, maxrel, maxabs);
    if (err != cudaSuccess) return 1;
  }
  {  // version 2: pre-split operands, cp.async double buffering
    float *dAh, *dAl, *dBh, *dBl;
    cudaMalloc(&dAh, hA.size() * 4); cudaMalloc(&dAl, hA.size() * 4); cudaMalloc(&dBh, hB.size() * 4); cudaMalloc(&dBl, hB.size() * 4);
    split_kernel<<<1024, 256>>>(dA, dAh, dAl, hA.size());
    split_kernel<<<1024, 256>>>(dB, dBh, dBl, hB.size());
    const size_t smem2 = NST * 4 * TILE_BYTES + 1024;
    cudaFuncSetAttribute(umma_gemm_nt_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    cudaMemset(dC, 0, hC.size() * 4);
    cudaMemset(dS, 0, 4);
    umma_gemm_nt_v2<<<grid, 128, smem2>>>(dAh, dAl, dBh, dBl, dC, M, N, K, dS, 1);
    cudaError_t err = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double maxrel = 0;
    for (int i = 0; i < M; i += 7)
      for (int j = 0; j < N; j += 5) {
        double ref = 0, sc = 0;
        for (int k = 0; k < K; ++k) { ref += (double)hA[(size_t)i * K + k] * hB[(size_t)j * K + k]; sc += fabs((double)hA[(size_t)i * K + k] * hB[(size_t)j * K + k]); }
        const double d = fabs(hC[(size_t)i * N + j] - ref);
        if (d / sc > maxrel) maxrel = d / sc;
      }
    printf("v2 (split, cp.async x2, chains of %d): err=%s timeout=%d  max |diff| / sum|terms| = %.3e\n", CHAIN, cudaGetErrorString(err), st, maxrel);
    if (err != cudaSuccess) return 1;
    if (argc > 4) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      for (int repeat : {1, 4, 16, -1, -2}) {  // -1: no copies after the prologue, -2: no MMAs (ablations)
        umma_gemm_nt_v2<<<grid, 128, smem2>>>(dAh, dAl, dBh, dBl, dC, M, N, K, dS, repeat);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) umma_gemm_nt_v2<<<grid, 128, smem2>>>(dAh, dAl, dBh, dBl, dC, M, N, K, dS, repeat);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("v2 repeat=%2d: %.3f ms per GEMM  %.2f TFLOP/s useful (2MNK x repeat), %.1f TFLOP/s tensor-pipe (x3)\n", repeat, ms / 10,
               2.0 * M * N * K * repeat / (ms / 10 * 1e-3) / 1e12, 6.0 * M * N * K * repeat / (ms / 10 * 1e-3) / 1e12);
      }
    }
  }
  // rate
  if (argc > 4) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int split = 0; split < 2; ++split) {
      umma_gemm_nt<<<grid, 128, smem>>>(dA, dB, dC, M, N, K, split, dS);
      cudaEventRecord(e0);
      for (int i = 0; i < 10; ++i) umma_gemm_nt<<<grid, 128, smem>>>(dA, dB, dC, M, N, K, split, dS);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("split=%d: %.3f ms per GEMM  %.2f TFLOP/s (2MNK)\n", split, ms / 10, 2.0 * M * N * K / (ms / 10 * 1e-3) / 1e12);
    }
  }
  printf("done\n");
  return 0;
}
