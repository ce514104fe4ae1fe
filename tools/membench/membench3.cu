// Throw-away probe 2: does a large shared-memory carve-out (small L1) throttle streaming loads, and which load
// flavour avoids it?  Mimics vec_tail_kernel: 148 CTAs x 512 threads, U vectors per lane per batch, table dot.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ double2 ldv(const double2* p) {
  double2 r;
  if (MODE == 0) r = __ldcs(p);
  else if (MODE == 1) r = *p;
  else if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  else if (MODE == 3) r = __ldcg(p);
  else if (MODE == 4) asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  else asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

template <int U, int MODE, int DOT>
__global__ void __launch_bounds__(512, 1) read_kernel(const double2* __restrict__ a, long long nvec, double* out, int range_batches, int tbl_n) {
  extern __shared__ double tbl[];
  if (DOT) for (int i = threadIdx.x; i < tbl_n; i += blockDim.x) tbl[i] = 1.0 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long batch_vecs = 32LL * U;
  const long long range_vecs = batch_vecs * range_batches;
  double s0 = 0, s1 = 0;
  for (long long r0 = warp * range_vecs; r0 < nvec; r0 += nwarps * range_vecs) {
    const long long r1 = (r0 + range_vecs < nvec) ? r0 + range_vecs : nvec;
    int toff = (int)(r0 % 61);
    for (long long b = r0; b + batch_vecs <= r1; b += batch_vecs) {
      double2 v[U];
#pragma unroll
      for (int s = 0; s < U; ++s) v[s] = ldv<MODE>(a + b + s * 32 + lane);
      if (DOT) {
        if (toff + U * 64 + 64 > tbl_n) toff = 1;
        const double* tp = tbl + toff + lane * 2;
#pragma unroll
        for (int s = 0; s < U; ++s) { s0 += v[s].x * tp[s * 64]; s1 += v[s].y * tp[s * 64 + 1]; }
        toff += U * 64;
      } else {
#pragma unroll
        for (int s = 0; s < U; ++s) { s0 += v[s].x; s1 += v[s].y; }
      }
    }
  }
  double s = s0 + s1;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) atomicAdd(out, s);
}

template <int U, int MODE, int DOT>
static void run(const double2* a, long long nvec, double* out, int smem_kb, int rb) {
  cudaFuncSetAttribute(read_kernel<U, MODE, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int tbl_n = smem_kb * 1024 / 8;
  for (int i = 0; i < 3; ++i) read_kernel<U, MODE, DOT><<<148, 512, smem_kb * 1024>>>(a, nvec, out, rb, tbl_n);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) read_kernel<U, MODE, DOT><<<148, 512, smem_kb * 1024>>>(a, nvec, out, rb, tbl_n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("U=%2d mode=%d dot=%d smem=%3d KB range_batches=%3d: %7.1f us  %6.0f GB/s %s\n", U, MODE, DOT, smem_kb, rb, ms / reps * 1e3,
         nvec * 16.0 / (ms / reps * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  const long long nvec = 34342525;
  double2* a;
  double* out;
  cudaMalloc(&a, nvec * 16);
  cudaMalloc(&out, 8);
  cudaMemset(a, 0, nvec * 16);
  cudaMemset(out, 0, 8);
  for (int rb : {1, 2, 4, 8, 13, 26, 52, 104, 208, 416, 1000}) run<8, 0, 1>(a, nvec, out, 172, rb);
  printf("done\n");
  return 0;
}
