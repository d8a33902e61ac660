// rmc_comm.cuh -- the exchange step of the minibatch-sharded large-batch learner (BASELINE configs[4]) as
// kernels over NVLink peer memory instead of library collectives:
//
//   k_comm_publish      local gradient blob, loss partial and (leaf, |td|) slice -> this rank's exchange buffer (one of two
//                       parity slots), then ONE release-store per peer of this step's epoch into the peer's flag array
//   k_comm_reduce_adam  waits for every rank's flag (local polling), then each thread sums ITS parameter's gradient over the
//                       ranks' buffers in rank order (peer loads over NVLink; fixed order -> every replica computes the same
//                       bits), applies Adam (+ Polyak) to that parameter in the same kernel, and the spare blocks gather the
//                       ranks' (leaf, |td|) slices into local arrays for the replicated priority write-back.
//
// One-shot all-reduce: every rank reads (W-1) x 152 KB; at W = 8 that is ~1 MB per rank per step, far below the
// 900 GB/s/direction of NVLink 5, so the exchange costs one flag round trip instead of a multi-step ring.
// Double buffering: rank r rewrites slot (t & 1) at step t+2 only after it has seen every peer's flag of step t+1,
// which a peer sends after (stream order) its step-t reads completed.
#pragma once
#include "rmc_mlp.cuh"
#include "rmc_tc_train.cuh"

namespace rmc {

constexpr int kCommMaxWorld = 8;
constexpr int kCommHeaderBytes = 4096;     // flags[2][kCommMaxWorld] + error word | (leaf,|td|) flags[2][kCommMaxWorld] at word 32, padded
constexpr int kCommTdFlagWord = 32;

struct CommView {
  unsigned char* base[kCommMaxWorld];      // exchange buffer of every rank (own rank: local memory)
  int rank, world;
  long long slot_bytes;                    // bytes of one parity slot
  long long off_loss, off_nodes, off_td;   // byte offsets inside a slot (gradients sit at 0)
  long long shard_lo[kCommMaxWorld + 1];   // global sample range of every rank
  unsigned long long timeout_ns;           // how long a waiting kernel polls for the peers' flags (RMC_COMM_TIMEOUT_MS, default 10 s)
  unsigned* verdict;                       // [2] local device words: {gradient exchange, (leaf,|td|) exchange}: (epoch << 1) | arrived
};

__device__ __forceinline__ unsigned char* comm_slot(const CommView& V, int r, int parity) {
  return V.base[r] + kCommHeaderBytes + static_cast<long long>(parity) * V.slot_bytes;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long ld_sys_s64(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Grid-uniform wait for the flags of all ranks (this rank's local flag array, one word per rank) to reach `epoch`.
// EVERY block of the calling kernel returns the same answer, so an exchange is applied by all blocks or by none (a
// per-block timer could let some blocks apply Adam while others give up).  Warp 0 of each block polls (lane r <-> rank r);
// the first block to conclude -- all flags arrived, or `timeout_ns` elapsed -- publishes the verdict of this epoch in a
// device word with a compare-and-swap from the previous epoch's value, the others adopt whatever verdict they find there.
// On a timeout the epoch is also left in the error word of the header (rmc_comm_status_sync).
__device__ __forceinline__ bool comm_wait_all(const CommView& V, const unsigned* flags, unsigned epoch, unsigned* verdict) {
  __shared__ int s_ok;
  if (threadIdx.x < 32) {
    const unsigned lane = threadIdx.x;
    unsigned d = __shfl_sync(0xffffffffu, ld_acquire_u32(verdict), 0);      // one view of the verdict per warp: the loop stays convergent
    const unsigned long long t0 = global_timer_ns();
    while ((d >> 1) != epoch) {
      const bool here = (static_cast<int>(lane) >= V.world) || (ld_acquire_sys_u32(flags + lane) == epoch);
      const bool all = __all_sync(0xffffffffu, here);
      unsigned want = 0u;
      if (all) want = (epoch << 1) | 1u;
      else if (global_timer_ns() - t0 > V.timeout_ns) want = epoch << 1;
      want = __shfl_sync(0xffffffffu, want, 0);          // lane 0's clock decides for the warp
      if (want == 0u) {
        __nanosleep(64);
        d = __shfl_sync(0xffffffffu, ld_acquire_u32(verdict), 0);
        continue;
      }
      if (lane == 0) {
        __threadfence();                                  // the flags this warp acquired are visible to whoever adopts the verdict
        const unsigned seen = atomicCAS(verdict, d, want);
        d = (seen == d) ? want : seen;
      }
      d = __shfl_sync(0xffffffffu, d, 0);
    }
    if (lane == 0) {
      s_ok = static_cast<int>(d & 1u);
      if (!(d & 1u)) reinterpret_cast<unsigned*>(V.base[V.rank])[2 * kCommMaxWorld] = epoch;
      __threadfence();
    }
  }
  __syncthreads();
  return s_ok != 0;
}

// Publish protocol (both publish kernels): every block copies its part with the loads of a thread batched, then ONE thread
// per block makes the block's stores visible at GPU scope (bar.sync + fence: cumulative over the block's threads) and arrives
// at a counter; the last block to arrive executes the single system-scope fence of the kernel and release-stores the epoch
// into every rank's flag array.  (The first version fenced at system scope in every thread of every block and copied with a
// load -> store loop: 13-14 us for 152 KB, on the critical path of the sharded step.)
__device__ __forceinline__ bool comm_block_arrive_last(unsigned* arrive) {
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(arrive, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  return s_last;
}
__device__ __forceinline__ void comm_raise_flags(const CommView& V, int flag_word0, int parity, unsigned epoch, unsigned* arrive) {
  if (threadIdx.x == 0) {
    *arrive = 0u;
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x < V.world) {     // flag of (this rank, this parity) in every rank's header, own included
    unsigned* flags = reinterpret_cast<unsigned*>(V.base[threadIdx.x]) + flag_word0;
    st_release_sys_u32(flags + parity * kCommMaxWorld + V.rank, epoch);
  }
}

__global__ void __launch_bounds__(256) k_comm_publish(CommView V, int parity, unsigned epoch, const float* __restrict__ grads, int total,
                                                      const float* __restrict__ loss, const long long* __restrict__ nodes,
                                                      const float* __restrict__ abs_td, long long n_local, unsigned* arrive) {
  pdl_enter();
  const SpanScope span_(SPAN_PUBLISH);
  unsigned char* slot = comm_slot(V, V.rank, parity);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  {   // gradient blob as float4 (total is a multiple of 4), 4 loads in flight per thread
    const float4* src = reinterpret_cast<const float4*>(grads);
    float4* dst = reinterpret_cast<float4*>(slot);
    const long long n4 = total >> 2;
    for (long long k0 = t0; k0 < n4; k0 += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) if (k0 + q * stride < n4) v[q] = __ldcg(src + k0 + q * stride);
#pragma unroll
      for (int q = 0; q < 4; ++q) if (k0 + q * stride < n4) dst[k0 + q * stride] = v[q];
    }
  }
  if (t0 == 0) *reinterpret_cast<float*>(slot + V.off_loss) = __ldcg(loss);
  if (nodes != nullptr) {
    long long* dn = reinterpret_cast<long long*>(slot + V.off_nodes);
    float* dt = reinterpret_cast<float*>(slot + V.off_td);
    for (long long k0 = t0; k0 < n_local; k0 += 4 * stride) {
      long long a[4]; float b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) if (k0 + q * stride < n_local) { a[q] = __ldcg(nodes + k0 + q * stride); b[q] = __ldcg(abs_td + k0 + q * stride); }
#pragma unroll
      for (int q = 0; q < 4; ++q) if (k0 + q * stride < n_local) { dn[k0 + q * stride] = a[q]; dt[k0 + q * stride] = b[q]; }
    }
  }
  if (!comm_block_arrive_last(arrive)) return;
  comm_raise_flags(V, 0, parity, epoch, arrive);
}

// Early exchange of the (leaf, |td|) slices (tensor-core mode: |td| exists right after the TD kernel, long before the
// gradients): publish + flag, then a gather kernel on a side stream feeds the replicated priority write-back, which so
// overlaps the backward pass and the gradient exchange.
__global__ void __launch_bounds__(256) k_comm_publish_td(CommView V, int parity, unsigned epoch, const long long* __restrict__ nodes,
                                                         const float* __restrict__ abs_td, long long n_local, unsigned* arrive) {
  pdl_enter();
  const SpanScope span_(SPAN_PUBLISH_TD);
  unsigned char* slot = comm_slot(V, V.rank, parity);
  long long* dn = reinterpret_cast<long long*>(slot + V.off_nodes);
  float* dt = reinterpret_cast<float*>(slot + V.off_td);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long k0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k0 < n_local; k0 += 4 * stride) {
    long long a[4]; float b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) if (k0 + q * stride < n_local) { a[q] = __ldcg(nodes + k0 + q * stride); b[q] = __ldcg(abs_td + k0 + q * stride); }
#pragma unroll
    for (int q = 0; q < 4; ++q) if (k0 + q * stride < n_local) { dn[k0 + q * stride] = a[q]; dt[k0 + q * stride] = b[q]; }
  }
  if (!comm_block_arrive_last(arrive)) return;
  comm_raise_flags(V, kCommTdFlagWord, parity, epoch, arrive);
}

// (leaf, |td|) of global sample i from the slot of the rank that owns it: one thread per sample (both loads of a sample in
// flight together; a per-rank loop would chain one NVLink round trip per rank)
__device__ __forceinline__ void comm_gather_rows(const CommView& V, int parity, long long* __restrict__ g_nodes, float* __restrict__ g_td, long long first,
                                                 long long stride) {
  const long long total = V.shard_lo[V.world];
  for (long long i = first; i < total; i += stride) {
    int r = 0;
    while (r + 1 < V.world && i >= V.shard_lo[r + 1]) ++r;
    const long long k = i - V.shard_lo[r];
    const long long node = ld_sys_s64(reinterpret_cast<const long long*>(comm_slot(V, r, parity) + V.off_nodes) + k);
    const float td = ld_sys_f32(reinterpret_cast<const float*>(comm_slot(V, r, parity) + V.off_td) + k);
    g_nodes[i] = node;
    g_td[i] = td;
  }
}

__global__ void __launch_bounds__(256) k_comm_gather_td(CommView V, int parity, unsigned epoch, long long* __restrict__ g_nodes, float* __restrict__ g_td) {
  pdl_enter();
  const SpanScope span_(SPAN_GATHER_TD);
  const unsigned* my_flags = reinterpret_cast<const unsigned*>(V.base[V.rank]);
  if (!comm_wait_all(V, my_flags + kCommTdFlagWord + parity * kCommMaxWorld, epoch, V.verdict + 1)) {
    // a peer never published its slice: hand the write-back that follows a list of out-of-range nodes (the tree kernels
    // skip those), so that nothing of a half-arrived batch reaches the tree
    for (long long k = blockIdx.x * 256ll + threadIdx.x; k < V.shard_lo[V.world]; k += static_cast<long long>(gridDim.x) * 256) { g_nodes[k] = -1; g_td[k] = 0.f; }
    return;
  }
  comm_gather_rows(V, parity, g_nodes, g_td, blockIdx.x * 256ll + threadIdx.x, static_cast<long long>(gridDim.x) * 256);
}

// blocks [0, param_blocks): parameters; the remaining blocks: (leaf, |td|) gather.
__global__ void __launch_bounds__(256) k_comm_reduce_adam(AgentCtx C, StepScalars S, CommView V, int parity, unsigned epoch, int param_blocks,
                                                          long long* __restrict__ g_nodes, float* __restrict__ g_td, int want_gather, TcPackOut P) {
  pdl_enter();
  const SpanScope span_(SPAN_COMM_REDUCE);
  const unsigned* my_flags = reinterpret_cast<const unsigned*>(V.base[V.rank]);
  if (!comm_wait_all(V, my_flags + parity * kCommMaxWorld, epoch, V.verdict)) {
    // all-or-nothing: no block applies Adam.  The (leaf, |td|) gather part hands the write-back out-of-range nodes.
    if (want_gather && static_cast<int>(blockIdx.x) >= param_blocks)
      for (long long k = (blockIdx.x - param_blocks) * 256ll + threadIdx.x; k < V.shard_lo[V.world]; k += static_cast<long long>(gridDim.x - param_blocks) * 256) { g_nodes[k] = -1; g_td[k] = 0.f; }
    return;
  }
  const NetLayout& L = C.L;
  if (static_cast<int>(blockIdx.x) < param_blocks) {
    const int pi = blockIdx.x * 256 + threadIdx.x;
    if (pi < L.total) {
      // all ranks' values in flight together, THEN the adds in rank order (a load-add loop is one NVLink round trip per rank:
      // ~1.5 us each, 12 us at 8 ranks)
      float v[kCommMaxWorld];
#pragma unroll
      for (int r = 0; r < kCommMaxWorld; ++r) v[r] = (r < V.world) ? ld_sys_f32(reinterpret_cast<const float*>(comm_slot(V, r, parity)) + pi) : 0.f;
      float g = 0.f;
#pragma unroll
      for (int r = 0; r < kCommMaxWorld; ++r) g += (r < V.world) ? v[r] : 0.f;
      C.grads[pi] = g;
      const float2 pt = adam_polyak_element(C, S, pi, g);
      tc_pack_updated(L, S, pi, pt, P);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      float lv[kCommMaxWorld];
#pragma unroll
      for (int r = 0; r < kCommMaxWorld; ++r) lv[r] = (r < V.world) ? ld_sys_f32(reinterpret_cast<const float*>(comm_slot(V, r, parity) + V.off_loss)) : 0.f;
      float loss = 0.f;
#pragma unroll
      for (int r = 0; r < kCommMaxWorld; ++r) loss += (r < V.world) ? lv[r] : 0.f;
      C.loss[0] = loss;
      host_loss_store(C.host_loss, loss, S.epoch);
    }
  } else if (want_gather) {
    comm_gather_rows(V, parity, g_nodes, g_td, (blockIdx.x - param_blocks) * 256ll + threadIdx.x, static_cast<long long>(gridDim.x - param_blocks) * 256);
  }
}

}  // namespace rmc
