// Throw-away bandwidth experiments behind the vector-contraction kernel design (not part of the library):
// how many bytes must be in flight per SM, and does a TMA L2 prefetch replace register staging?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu && ./membench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int U, int PF>
__global__ void __launch_bounds__(512, 1) read_kernel(const double2* __restrict__ a, long long nvec, double* out, int pf_dist, int range_batches) {
  // each warp owns contiguous ranges of `range_batches` batches (like the product kernel), claimed round-robin
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long batch_vecs = 32LL * U;
  const long long range_vecs = batch_vecs * range_batches;
  double s0 = 0, s1 = 0;
  for (long long r0 = warp * range_vecs; r0 < nvec; r0 += nwarps * range_vecs) {
    const long long r1 = (r0 + range_vecs < nvec) ? r0 + range_vecs : nvec;
    if (PF) {
      // lane l prefetches batch 1 + l of the range (up to pf_dist batches ahead)
      const long long p = r0 + (1 + lane) * batch_vecs;
      if (lane < pf_dist && p + batch_vecs <= r1)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a + p), "r"((int)(batch_vecs * 16)) : "memory");
    }
    for (long long b = r0; b + batch_vecs <= r1; b += batch_vecs) {
      if (PF) {
        const long long p = b + (long long)(1 + pf_dist) * batch_vecs;
        if (lane == 0 && p + batch_vecs <= r1)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a + p), "r"((int)(batch_vecs * 16)) : "memory");
      }
      double2 v[U];
#pragma unroll
      for (int s = 0; s < U; ++s) v[s] = __ldcs(a + b + s * 32 + lane);
#pragma unroll
      for (int s = 0; s < U; ++s) { s0 += v[s].x; s1 += v[s].y; }
    }
  }
  double s = s0 + s1;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) atomicAdd(out, s);
}

template <int U, int PF>
static void run(const char* name, const double2* a, long long nvec, double* out, int grid, int pf, int rb) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) read_kernel<U, PF><<<grid, 512>>>(a, nvec, out, pf, rb);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) read_kernel<U, PF><<<grid, 512>>>(a, nvec, out, pf, rb);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("%-28s U=%2d pf=%2d range_batches=%3d grid=%4d: %7.1f us  %6.0f GB/s %s\n", name, U, pf, rb, grid, ms / reps * 1e3,
         nvec * 16.0 / (ms / reps * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  const long long nvec = 34342525;  // 549.5 MB, the rank-4 dim-200 fp64 tensor
  double2* a;
  double* out;
  cudaMalloc(&a, nvec * 16);
  cudaMalloc(&out, 8);
  cudaMemset(a, 0, nvec * 16);
  cudaMemset(out, 0, 8);
  for (int rb : {13, 52, 1000000}) {
    run<8, 0>("plain", a, nvec, out, 148, 0, rb);
    run<16, 0>("plain", a, nvec, out, 148, 0, rb / 2 > 0 ? rb / 2 : 1);
    run<4, 0>("plain", a, nvec, out, 148, 0, rb * 2);
    for (int pf : {2, 4, 8, 16}) run<8, 1>("l2-prefetch", a, nvec, out, 148, pf, rb);
    for (int pf : {4, 8, 16}) run<4, 1>("l2-prefetch", a, nvec, out, 148, pf, rb * 2);
  }
  run<8, 0>("plain 2 CTAs/SM-equivalent", a, nvec, out, 296, 0, 13);
  printf("done\n");
  return 0;
}
