// Throw-away probe: FP64 FMA (DFMA) and FP64 tensor (DMMA m8n8k4) throughput per SM on this part.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(512, 1) ffma_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(512, 1) dmma_kernel(double* out, int iters, double a, double b) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double av = a + threadIdx.x, bv = b;
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(av), "d"(bv));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(av), "d"(bv));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(av), "d"(bv));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(av), "d"(bv));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 512 * 8 * 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  const int iters = 4096;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<<<148, 512>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", ms, 148.0 * 512 * iters * 8 * 2 / (ms * 1e-3) / 1e12,
           512.0 * iters * 8 / (ms * 1e-3 * 1.9e9));
    cudaEventRecord(e0);
    ffma_kernel<<<148, 512>>>((float*)out, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", ms, 148.0 * 512 * iters * 8 * 2 / (ms * 1e-3) / 1e12,
           512.0 * iters * 8 / (ms * 1e-3 * 1.9e9));
    cudaEventRecord(e0);
    dmma_kernel<<<148, 512>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // one m8n8k4 = 256 FMA per warp
    printf("DMMA m8n8k4: %.3f ms  %.2f TFLOP/s\n", ms, 148.0 * 16 * iters * 4 * 256 * 2 / (ms * 1e-3) / 1e12);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
