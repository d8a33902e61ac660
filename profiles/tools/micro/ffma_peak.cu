// FP32 FFMA peak of the box (not in MEASURED_PEAKS.json): 148 x k CTAs of 256 threads, 8 independent
// FMA chains per thread, timed with CUDA events.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) ffma(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  const int iters = 20000;
  for (int ctas_per_sm : {4, 8}) {
    ffma<<<148 * ctas_per_sm, 256>>>(out, 100, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(s);
      ffma<<<148 * ctas_per_sm, 256>>>(out, iters, 1.0001f, 0.5f);
      cudaEventRecord(e); cudaEventSynchronize(e);
      float ms; cudaEventElapsedTime(&ms, s, e); best = ms < best ? ms : best;
    }
    const double flops = 2.0 * 148 * ctas_per_sm * 256 * 8.0 * 16 * iters;
    printf("ffma peak: %d CTAs/SM x 256 thr: %.3f ms -> %.1f TFLOP/s fp32\n", ctas_per_sm, best, flops / (best * 1e-3) / 1e12);
  }
  return 0;
}
