// FP32 FFMA rate of a register-blocked outer product (the shape of an SGEMM inner loop: every FFMA reads two operand
// registers and an accumulator), without any shared-memory traffic: the practical ceiling of a GEMM-shaped FFMA stream,
// next to the constant-operand peak of ffma_peak.cu.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int R, int C>
__global__ void outer(float* out, const float* in, int iters) {
  float a[R], b[C], acc[R][C];
  for (int r = 0; r < R; ++r) a[r] = in[threadIdx.x + r];
  for (int c = 0; c < C; ++c) b[c] = in[64 + threadIdx.x + c];
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) acc[r][c] = 0.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
      // rotate the operands (register moves would be eliminated; a dependent cheap op keeps them live and changing)
      a[u % R] = __int_as_float(__float_as_int(a[u % R]) ^ i);
    }
  }
  float s = 0.f;
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) s += acc[r][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int R, int C>
void run(const char* name, int threads, int ctas_per_sm) {
  float *out, *in; cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  const int iters = 5000;
  outer<R, C><<<148 * ctas_per_sm, threads>>>(out, in, 10);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int k = 0; k < 5; ++k) {
    cudaEventRecord(s);
    outer<R, C><<<148 * ctas_per_sm, threads>>>(out, in, iters);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e); best = ms < best ? ms : best;
  }
  const double flops = 2.0 * 148 * ctas_per_sm * threads * double(R * C) * 4 * iters;
  printf("%s: %d thr x %d CTA/SM: %.3f ms -> %.1f TFLOP/s\n", name, threads, ctas_per_sm, best, flops / (best * 1e-3) / 1e12);
  cudaFree(out); cudaFree(in);
}
int main() {
  run<8, 8>("outer 8x8", 256, 1);
  run<8, 8>("outer 8x8", 512, 1);
  run<8, 8>("outer 8x8", 256, 3);
  run<4, 8>("outer 4x8", 256, 1);
  run<4, 8>("outer 4x8", 512, 1);
  run<4, 8>("outer 4x8", 256, 4);
  return 0;
}
