// Throw-away probe 4: per-warp cp.async.bulk (TMA 1-D bulk copy) rings feeding a table dot product from shared
// memory, against a plain LDG streaming read.  Mimics the planned vec_ring_kernel: 148 CTAs, NW warps per CTA,
// every warp owns R ring slots of B bytes; tiles of TILE bytes are dealt round-robin to the warps of the grid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ringbench ringbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ring kernel: TILE, B in elements (double)
__global__ void __launch_bounds__(1024, 1) ring_kernel(const double* __restrict__ a, long long n, double* out, int R, int Bel, int tile_el, int tbl_n, int dot) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
  double* tbl = reinterpret_cast<double*>(smem);
  double* ring = tbl + tbl_n;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)NW * R * Bel);
  for (int i = threadIdx.x; i < tbl_n; i += blockDim.x) tbl[i] = 1.0 + i * 1e-9;
  double* myring = ring + (size_t)warp * R * Bel;
  const uint32_t bar0 = smem_u32(bars + warp * R);
  if (lane == 0) for (int r = 0; r < R; ++r) mbar_init(bar0 + 8 * r, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long W = (long long)gridDim.x * NW, gw = (long long)blockIdx.x * NW + warp;
  const long long ntiles = n / tile_el;
  const int spt = tile_el / Bel;  // sub-chunks per tile
  // produce cursor
  long long pt = gw;
  int ps = 0, pc = 0;
  auto issue = [&]() {
    if (pt < ntiles) {
      if (lane == 0) {
        const int slot = pc % R;
        const uint32_t bar = bar0 + 8 * slot;
        fence_proxy_async();
        mbar_expect_tx(bar, Bel * 8);
        bulk_g2s(smem_u32(myring + (size_t)slot * Bel), a + pt * tile_el + (long long)ps * Bel, Bel * 8, bar);
      }
      ++pc;
      if (++ps == spt) { ps = 0; pt += W; }
    }
  };
  for (int r = 0; r < R; ++r) issue();
  double s0 = 0, s1 = 0;
  int c = 0;
  int toff = 1;
  for (long long t = gw; t < ntiles; t += W) {
    for (int k = 0; k < spt; ++k, ++c) {
      const int slot = c % R;
      mbar_wait(bar0 + 8 * slot, (c / R) & 1);
      const double* sd = myring + (size_t)slot * Bel;
      if (dot) {
        if (toff + Bel + 64 > tbl_n) toff = 1;
        const double* tp = tbl + toff;
#pragma unroll 4
        for (int e = lane; e < Bel; e += 64) { s0 += sd[e] * tp[e]; s1 += sd[e + 32] * tp[e + 32]; }
        toff += Bel;
      } else {
#pragma unroll 4
        for (int e = lane; e < Bel; e += 64) { s0 += sd[e]; s1 += sd[e + 32]; }
      }
      __syncwarp();
      issue();
    }
  }
  double s = s0 + s1;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) atomicAdd(out, s);
}

// plain LDG read: U 16-byte vectors per lane in flight
template <int U>
__global__ void __launch_bounds__(1024) ldg_kernel(const double2* __restrict__ a, long long nvec, double* out) {
  const long long nthreads = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double s0 = 0, s1 = 0;
  long long i = tid;
  for (; i + (U - 1) * nthreads < nvec; i += U * nthreads) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(a + i + u * nthreads);
#pragma unroll
    for (int u = 0; u < U; ++u) { s0 += v[u].x; s1 += v[u].y; }
  }
  for (; i < nvec; i += nthreads) { const double2 v = __ldcs(a + i); s0 += v.x; s1 += v.y; }
  double s = s0 + s1;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

static double* g_a;
static double* g_out;
static long long g_n;

static void run_ring(int NW, int R, int Bkb_x4, int tile_kb, int tbl_kb, int dot) {
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int Bel = Bkb_x4 * 256 / 8;  // B given in units of 256 bytes
  const int tile_el = tile_kb * 1024 / 8;
  const int tbl_n = tbl_kb * 1024 / 8;
  const size_t smem = (size_t)tbl_n * 8 + (size_t)NW * R * Bel * 8 + (size_t)NW * R * 8 + 128;
  if (smem > 227 * 1024) { printf("NW=%d R=%d B=%d tbl=%d: smem %zu too large\n", NW, R, Bel * 8, tbl_kb, smem); return; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const long long n = g_n / tile_el * tile_el;
  for (int i = 0; i < 3; ++i) ring_kernel<<<148, NW * 32, smem>>>(g_a, n, g_out, R, Bel, tile_el, tbl_n, dot);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) ring_kernel<<<148, NW * 32, smem>>>(g_a, n, g_out, R, Bel, tile_el, tbl_n, dot);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("ring NW=%2d R=%d B=%5d tile=%3dKB tbl=%3dKB ring=%3zuKB dot=%d: %7.1f us  %6.0f GB/s %s\n", NW, R, Bel * 8, tile_kb, tbl_kb,
         (size_t)NW * R * Bel * 8 / 1024, dot, ms / reps * 1e3, n * 8.0 / (ms / reps * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
  fflush(stdout);
}

template <int U>
static void run_ldg(int ctas_per_sm, int threads) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const long long nvec = g_n / 2;
  for (int i = 0; i < 3; ++i) ldg_kernel<U><<<148 * ctas_per_sm, threads>>>((const double2*)g_a, nvec, g_out);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) ldg_kernel<U><<<148 * ctas_per_sm, threads>>>((const double2*)g_a, nvec, g_out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("ldg  U=%d ctas/sm=%d threads=%4d (%3d KB/SM in flight): %7.1f us  %6.0f GB/s %s\n", U, ctas_per_sm, threads, U * 16 * ctas_per_sm * threads / 1024,
         ms / reps * 1e3, nvec * 16.0 / (ms / reps * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
  fflush(stdout);
}

int main() {
  g_n = 68685056;
  cudaMalloc(&g_a, g_n * 8);
  cudaMalloc(&g_out, 8);
  cudaMemset(g_a, 0, g_n * 8);
  cudaMemset(g_out, 0, 8);
  run_ldg<1>(2, 1024);
  run_ldg<2>(2, 1024);
  run_ldg<4>(2, 1024);
  run_ldg<4>(1, 1024);
  run_ldg<4>(1, 512);
  run_ldg<8>(1, 512);
  run_ldg<8>(1, 1024);
  // no table: how much ring does the memory system want?
  for (int dot = 0; dot < 2; ++dot) {
    run_ring(16, 2, 4, 16, dot ? 8 : 0, dot);
    run_ring(16, 2, 8, 16, dot ? 8 : 0, dot);
    run_ring(16, 3, 8, 16, dot ? 8 : 0, dot);
    run_ring(16, 4, 8, 16, dot ? 8 : 0, dot);
    run_ring(16, 2, 16, 16, dot ? 8 : 0, dot);
    run_ring(16, 3, 16, 16, dot ? 8 : 0, dot);
    run_ring(8, 4, 16, 16, dot ? 8 : 0, dot);
    run_ring(32, 2, 8, 16, dot ? 8 : 0, dot);
    run_ring(32, 3, 8, 16, dot ? 8 : 0, dot);
    run_ring(32, 2, 16, 16, dot ? 8 : 0, dot);
  }
  // with the 160 KB table of rank 4 dim 200 fp64: ring <= ~56 KB
  run_ring(16, 3, 4, 16, 160, 1);
  run_ring(16, 2, 4, 16, 160, 1);
  run_ring(16, 2, 6, 12, 160, 1);
  run_ring(16, 3, 4, 8, 160, 1);
  run_ring(16, 3, 4, 32, 160, 1);
  run_ring(8, 3, 8, 16, 160, 1);
  run_ring(8, 2, 8, 16, 160, 1);
  run_ring(8, 6, 4, 16, 160, 1);
  run_ring(24, 2, 4, 16, 160, 1);
  run_ring(12, 2, 8, 16, 160, 1);
  run_ring(12, 4, 4, 16, 160, 1);
  run_ring(32, 3, 2, 16, 160, 1);
  run_ring(32, 2, 3, 12, 160, 1);
  printf("done\n");
  return 0;
}
