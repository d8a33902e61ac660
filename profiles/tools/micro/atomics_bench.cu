// Microbenchmark (diagnostic): latency of a burst of scattered L2 reductions from one / several CTAs.
// f64 RED vs u64 RED vs u32 RED vs plain load+store; 256 threads x 12 ops, addresses like sum-tree ancestors.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int MODE>
__global__ void k(double* tree, unsigned long long* out, long long cap, int per_thread_levels, unsigned seed) {
  __shared__ unsigned long long t0;
  const int tid = threadIdx.x;
  unsigned x = seed + 7919u * (blockIdx.x * blockDim.x + tid);
  x ^= x << 13; x ^= x >> 17; x ^= x << 5;
  long long leaf = cap - 1 + (x % cap);
  __syncthreads();
  if (tid == 0) t0 = gt();
  __syncthreads();
  long long n = leaf;
  for (int l = 0; l < per_thread_levels; ++l) {
    n = (n - 1) >> 1;
    if (MODE == 0) atomicAdd(tree + n, 0.25);
    if (MODE == 1) atomicAdd(reinterpret_cast<unsigned long long*>(tree + n), 3ull);
    if (MODE == 2) atomicAdd(reinterpret_cast<unsigned*>(tree + n), 3u);
    if (MODE == 3) tree[n] = tree[n] + 0.25;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) out[blockIdx.x] = gt() - t0;
}
int main() {
  const long long cap = 1000000;
  double* tree; unsigned long long* out;
  cudaMalloc(&tree, sizeof(double) * 2 * cap); cudaMemset(tree, 0, sizeof(double) * 2 * cap);
  cudaMalloc(&out, 8 * 64);
  unsigned long long h[64];
  const char* names[] = {"red.f64", "red.u64", "red.u32", "ld+st f64"};
  for (int ctas : {1, 8}) for (int mode = 0; mode < 4; ++mode) for (int levels : {12, 20}) {
    unsigned long long best = ~0ull, sum = 0;
    for (int it = 0; it < 20; ++it) {
      if (mode == 0) k<0><<<ctas, 256 / (ctas == 8 ? 8 : 1)>>>(tree, out, cap, levels, it);
      if (mode == 1) k<1><<<ctas, 256 / (ctas == 8 ? 8 : 1)>>>(tree, out, cap, levels, it);
      if (mode == 2) k<2><<<ctas, 256 / (ctas == 8 ? 8 : 1)>>>(tree, out, cap, levels, it);
      if (mode == 3) k<3><<<ctas, 256 / (ctas == 8 ? 8 : 1)>>>(tree, out, cap, levels, it);
      cudaMemcpy(h, out, 8 * ctas, cudaMemcpyDeviceToHost);
      unsigned long long mx = 0; for (int c = 0; c < ctas; ++c) mx = h[c] > mx ? h[c] : mx;
      if (it >= 5) { best = mx < best ? mx : best; sum += mx; }
    }
    printf("%-10s ctas=%d threads/cta=%d levels=%d : min %llu ns, mean %llu ns\n", names[mode], ctas, 256 / (ctas == 8 ? 8 : 1), levels, best, sum / 15);
  }
  return 0;
}
