// Probe: how long does a burst of same-address atomicAdd-with-return take when every warp of a persistent grid
// (148 CTAs x 16 warps, as vec_ring_kernel) claims at once -- against the same claims spread over S counters 256 B apart.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_probe atomic_probe.cu && ./atomic_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>

__global__ void __launch_bounds__(512, 1) probe(unsigned long long* ctr, int stripes, int rounds, unsigned long long* t_out, unsigned long long* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  unsigned long long acc = 0;
  __syncthreads();
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int r = 0; r < rounds; ++r) {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(ctr + (size_t)(gw % stripes) * 32, 1ULL);
    v = __shfl_sync(0xffffffffu, v, 0);
    acc += v;
  }
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (lane == 0) { t_out[2 * gw] = t0; t_out[2 * gw + 1] = t1; sink[gw] = acc; }
}

int main() {
  const int G = 148, NW = 16, W = G * NW;
  unsigned long long *ctr, *t, *sink;
  cudaMalloc(&ctr, 64 * 32 * 8);
  cudaMalloc(&t, W * 2 * 8);
  cudaMalloc(&sink, W * 8);
  std::vector<unsigned long long> h(W * 2);
  for (int stripes : {1, 2, 4, 8, 16, 64}) {
    for (int rounds : {1, 4}) {
      double best = 1e30, avg_lat = 0;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(ctr, 0, 64 * 32 * 8);
        probe<<<G, NW * 32>>>(ctr, stripes, rounds, t, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), t, W * 2 * 8, cudaMemcpyDeviceToHost);
        unsigned long long lo = ~0ULL, hi = 0;
        double lat = 0;
        for (int w = 0; w < W; ++w) { lo = std::min(lo, h[2 * w]); hi = std::max(hi, h[2 * w + 1]); lat += (double)(h[2 * w + 1] - h[2 * w]); }
        if ((double)(hi - lo) < best) { best = (double)(hi - lo); avg_lat = lat / W; }
      }
      printf("stripes %2d rounds %d: burst of %d claims/round drained in %8.2f us (first start to last return), mean per-warp time %7.2f us, %6.2f ns per claim\n",
             stripes, rounds, W, best / 1e3, avg_lat / 1e3, best / (double)(W * rounds));
    }
  }
  return 0;
}
