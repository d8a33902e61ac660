// rmc_tree.cuh -- GPU-resident sum tree (reference heap layout, float64 nodes) + min/max
// summaries, stratified prefix-search sampling and batched priority write-back.
//
// Reference semantics restated here:
//   dqn/utils/sum_tree.py:15-32  update  -> leaf store + exact ancestor fix-up (float64 adds of
//                                           float32-exact values are exact in any order, SURVEY
//                                           finding 6, so atomics are bit-identical to the
//                                           reference's sequential propagation)
//   dqn/utils/sum_tree.py:42-61  get_leaf-> per_descend_warp (same `v <= left` / `v -= left`)
//   dqn/utils/sum_tree.py:67-73  max/min -> values of max/min(leaves[:size]) kept by two summary
//                                           levels (fan-in kBlk) instead of index tracking + rescans
#pragma once
#include "rmc_device.cuh"

namespace rmc {

__device__ __forceinline__ float finf() { return __int_as_float(0x7f800000); }

// ---- stratified prefix search: one warp per sample, 4 tree levels per L2 round trip -------------
// All 32 lanes call with identical (v); returns the leaf's tree index on every lane.
__device__ __forceinline__ long long per_descend_warp(const double* __restrict__ tree, long long n_nodes,
                                                      double v, double* leaf_val) {
  const int lane = threadIdx.x & 31;
  int d = 0, o = 0;
  if (lane < 30) {  // lanes 0..29 <-> the 2+4+8+16 descendants at relative depth 1..4
    d = 31 - __clz(lane + 2);
    o = lane + 2 - (1 << d);
  }
  long long p = 0;
  while (2 * p + 1 < n_nodes) {
    double val = 0.0;
    if (lane < 30) {
      const long long idx = ((p + 1) << d) - 1 + o;
      if (idx < n_nodes) val = __ldcg(tree + idx);
    }
    long long cur = p;
    int oc = 0;
    bool leaf = false;
#pragma unroll
    for (int dd = 1; dd <= 4; ++dd) {
      const long long left = 2 * cur + 1;
      if (left >= n_nodes) { leaf = true; break; }
      const double lv = __shfl_sync(0xffffffffu, val, (1 << dd) - 2 + 2 * oc);
      if (v <= lv) { cur = left; oc = 2 * oc; }
      else { v = v - lv; cur = left + 1; oc = 2 * oc + 1; }
    }
    p = cur;
    if (leaf) break;
  }
  *leaf_val = __ldcg(tree + p);
  return p;
}

// value drawn in stratum i (dqn/replay_memory.py:72,80; np.random.uniform(lo,hi) == lo+(hi-lo)*u)
__device__ __forceinline__ double stratum_value(double total, long long Bglobal, long long i, double u) {
  const double seg = total / static_cast<double>(Bglobal);
  const double lo = seg * static_cast<double>(i);
  const double hi = seg * static_cast<double>(i + 1);
  return lo + (hi - lo) * u;   // -fmad=false: mul and add round separately, like numpy
}

// importance weight (dqn/replay_memory.py:76-77,84-86), float64 then cast by the caller
__device__ __forceinline__ double is_weight(double size, double p, double total, double min_p, double beta) {
  const double max_w = pow(size * (min_p / total), -beta);
  return pow(size * (p / total), -beta) / max_w;
}

// |td| -> priority (dqn/replay_memory.py:95): float32 min/add, pow evaluated in float64 and rounded
// once to float32 (= correctly rounded powf; numpy's SIMD powf is within 1 ulp of this).
__device__ __forceinline__ float td_to_priority(float abs_td, float eps, float alpha, float pmax) {
  const float x = fminf(abs_td + eps, pmax);
  return static_cast<float>(pow(static_cast<double>(x), static_cast<double>(alpha)));
}

// ---- min/max summaries ----------------------------------------------------------------------
__device__ __forceinline__ void minmax_l0(const ReplayDev& R, long long size, long long b) {
  const double* leaves = R.tree + (R.cap - 1);
  const long long lo = b * kBlk;
  float mn = finf(), mx = 0.f;
#pragma unroll 8
  for (int j = 0; j < kBlk; ++j) {
    const long long k = lo + j;
    if (k < size) {
      const float p = static_cast<float>(__ldcg(leaves + k));
      mn = fminf(mn, p);
      mx = fmaxf(mx, p);
    }
  }
  R.b0min[b] = mn;
  R.b0max[b] = mx;
}
__device__ __forceinline__ void minmax_l1(const ReplayDev& R, long long c) {
  const long long lo = c * kBlk;
  float mn = finf(), mx = 0.f;
#pragma unroll 8
  for (int j = 0; j < kBlk; ++j) {
    const long long k = lo + j;
    if (k < R.n0) {
      mn = fminf(mn, __ldcg(R.b0min + k));
      mx = fmaxf(mx, __ldcg(R.b0max + k));
    }
  }
  R.b1min[c] = mn;
  R.b1max[c] = mx;
}
// one warp
__device__ __forceinline__ void minmax_global_warp(const ReplayDev& R) {
  const int lane = threadIdx.x & 31;
  float mn = finf(), mx = 0.f;
  for (long long k = lane; k < R.n1; k += 32) {
    mn = fminf(mn, __ldcg(R.b1min + k));
    mx = fmaxf(mx, __ldcg(R.b1max + k));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  }
  if (lane == 0) {
    R.st->min_p = mn;
    R.st->max_p = mx;
  }
}

// leaf store + exact ancestor fix-up for the elected writer of a leaf
__device__ __forceinline__ void tree_set_leaf(const ReplayDev& R, long long leaf, float p) {
  const double np = static_cast<double>(p);
  const double old = __ldcg(R.tree + leaf);
  R.tree[leaf] = np;
  const double delta = np - old;
  if (delta != 0.0) {
    long long n = leaf;
    while (n != 0) {
      n = (n - 1) >> 1;
      atomicAdd(R.tree + n, delta);
    }
  }
}

// Whole write-back by ONE CTA (n <= kTreeCtaMax): duplicates -> last in batch order wins
// (dqn/replay_memory.py:97-98 applies updates sequentially).
__device__ void tree_update_cta(const ReplayDev& R, const long long* __restrict__ nodes,
                                const float* __restrict__ pri, long long n, long long size, bool stamps_done) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const long long first_leaf = R.cap - 1;
  if (!stamps_done) {
    for (long long i = tid; i < n; i += nt) atomicMax(R.stamps + (nodes[i] - first_leaf), static_cast<int>(i + 1));
    __syncthreads();
  }
  for (long long i = tid; i < n; i += nt) {
    const long long leaf = nodes[i];
    int* st = R.stamps + (leaf - first_leaf);
    if (__ldcg(st) == static_cast<int>(i + 1)) {
      *st = 0;
      tree_set_leaf(R, leaf, pri[i]);
    }
  }
  __syncthreads();
  for (long long i = tid; i < n; i += nt) minmax_l0(R, size, (nodes[i] - first_leaf) / kBlk);
  __syncthreads();
  for (long long i = tid; i < n; i += nt) minmax_l1(R, ((nodes[i] - first_leaf) / kBlk) / kBlk);
  __syncthreads();
  if (tid < 32) minmax_global_warp(R);
}

// ---- standalone kernels -------------------------------------------------------------------
// one CTA: rmc_per_update for n <= kTreeCtaMax; optional |td| -> priority conversion first
__global__ void __launch_bounds__(kThreads) k_tree_update_small(ReplayDev R, const long long* nodes, const float* pri_in,
                                                                const float* abs_td, float* pri_out, long long n,
                                                                float eps, float alpha, float pmax) {
  const float* pri = pri_in;
  if (abs_td != nullptr) {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) pri_out[i] = td_to_priority(abs_td[i], eps, alpha, pmax);
    __syncthreads();
    pri = pri_out;
  }
  tree_update_cta(R, nodes, pri, n, R.st->size, false);
}

// grid-wide variants for large batches
__global__ void k_td_to_pri(const float* abs_td, float* pri, long long n, float eps, float alpha, float pmax) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) pri[i] = td_to_priority(abs_td[i], eps, alpha, pmax);
}
__global__ void k_tree_stamp(ReplayDev R, const long long* nodes, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) atomicMax(R.stamps + (nodes[i] - (R.cap - 1)), static_cast<int>(i + 1));
}
__global__ void k_tree_apply(ReplayDev R, const long long* nodes, const float* pri, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) {
    const long long leaf = nodes[i];
    int* st = R.stamps + (leaf - (R.cap - 1));
    if (__ldcg(st) == static_cast<int>(i + 1)) {
      *st = 0;
      tree_set_leaf(R, leaf, pri[i]);
    }
  }
}
__global__ void k_minmax_l0_all(ReplayDev R) {
  const long long b = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (b < R.n0) minmax_l0(R, R.st->size, b);
}
__global__ void k_minmax_l1_all(ReplayDev R) {
  const long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (c < R.n1) minmax_l1(R, c);
}
__global__ void k_minmax_global(ReplayDev R) { minmax_global_warp(R); }

// bottom-up rebuild of one heap level: nodes [first, first+count)
__global__ void k_tree_rebuild_level(double* tree, long long first, long long count) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < count) {
    const long long n = first + i;
    tree[n] = tree[2 * n + 1] + tree[2 * n + 2];
  }
}
__global__ void k_set_leaves(ReplayDev R, const float* pri, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) R.tree[R.cap - 1 + i] = static_cast<double>(pri[i]);
}

// ---- push (store_transitions) ---------------------------------------------------------------
// pack separate device arrays into AoS rows
__global__ void k_pack_rows(float* dst, const float* obs, const long long* act, const float* rew, const float* done,
                            const float* nxt, long long n, int D, int row_floats) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long row = t / row_floats;
  const int c = static_cast<int>(t % row_floats);
  if (row >= n) return;
  float v = 0.f;
  if (c < D) v = obs[row * D + c];
  else if (c < 2 * D) v = nxt[row * D + (c - D)];
  else if (c == 2 * D) v = __int_as_float(static_cast<int>(act[row]));
  else if (c == 2 * D + 1) v = rew[row];
  else if (c == 2 * D + 2) v = done[row];
  dst[t] = v;
}

// n <= kTreeCtaMax rows, ONE CTA: ring write at data_pointer, leaves <- max_priority (1.0 if 0, read
// once per call: dqn/replay_memory.py:57-60), size bumped before the update (sum_tree.py:34-40).
__global__ void __launch_bounds__(kThreads) k_push_small(ReplayDev R, const float* __restrict__ rows, long long n,
                                                         long long* scratch_nodes, float* scratch_pri, float pmax) {
  __shared__ long long s_dp, s_size;
  __shared__ float s_p;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    s_dp = R.st->dp;
    s_size = R.st->size;
    const float mp = R.st->max_p;
    s_p = (mp == 0.f) ? pmax : mp;
  }
  __syncthreads();
  const long long dp = s_dp;
  const long long new_size = min(s_size + n, R.cap);
  const float p = s_p;
  const int rf = R.row_floats;
  for (long long t = tid; t < n * rf; t += nt) {
    const long long j = t / rf;
    const int c = static_cast<int>(t % rf);
    R.ring[((dp + j) % R.cap) * rf + c] = rows[t];
  }
  if (R.prioritized) {
    for (long long j = tid; j < n; j += nt) {
      scratch_nodes[j] = (dp + j) % R.cap + (R.cap - 1);
      scratch_pri[j] = p;
    }
    __syncthreads();
    tree_update_cta(R, scratch_nodes, scratch_pri, n, new_size, false);
  }
  __syncthreads();
  if (tid == 0) {
    R.st->dp = (dp + n) % R.cap;
    R.st->size = new_size;
  }
}

// bulk path pieces
__global__ void k_push_begin(ReplayDev R, float pmax) {
  const float mp = R.st->max_p;
  R.st->push_p = (mp == 0.f) ? pmax : mp;
}
__global__ void k_push_rows_bulk(ReplayDev R, const float* __restrict__ rows, long long n, long long dp) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int rf = R.row_floats;
  if (t >= n * rf) return;
  const long long j = t / rf;
  const int c = static_cast<int>(t % rf);
  const long long slot = (dp + j) % R.cap;
  R.ring[slot * rf + c] = rows[t];
  if (c == 0 && R.prioritized) R.tree[R.cap - 1 + slot] = static_cast<double>(R.st->push_p);
}
__global__ void k_push_end(ReplayDev R, long long dp, long long size) {
  R.st->dp = dp;
  R.st->size = size;
}

// ---- standalone samplers (ReplayMemory*.sample_transitions) ---------------------------------
__device__ __forceinline__ void gather_row_warp(const ReplayDev& R, long long slot, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const float* src = R.ring + slot * R.row_floats;
  for (int c = lane; c < R.row_floats; c += 32) dst[c] = __ldcg(src + c);
}

__global__ void __launch_bounds__(kThreads) k_per_sample(ReplayDev R, long long B, long long Bglobal, long long shard_off,
                                                         double beta, const double* u, unsigned long long seed,
                                                         unsigned long long counter, unsigned agent, long long* out_nodes,
                                                         float* out_w, float* out_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = blockIdx.x * static_cast<long long>(kWarps) + warp;
  if (i >= B) return;
  const long long n_nodes = 2 * R.cap - 1;
  const double total = __ldcg(R.tree);
  const long long size = R.st->size;
  const double ui = (u != nullptr) ? u[i] : philox_uniform(seed, counter, agent, static_cast<uint32_t>(shard_off + i));
  const double v = stratum_value(total, Bglobal, shard_off + i, ui);
  double p;
  const long long leaf = per_descend_warp(R.tree, n_nodes, v, &p);
  if (lane == 0) {
    out_nodes[i] = leaf;
    if (out_w != nullptr)
      out_w[i] = static_cast<float>(is_weight(static_cast<double>(size), p, total, static_cast<double>(R.st->min_p), beta));
  }
  if (out_rows != nullptr) gather_row_warp(R, leaf - (R.cap - 1), out_rows + i * R.row_floats);
}

// SumTree.get_leaf for explicit prefix values
__global__ void __launch_bounds__(kThreads) k_tree_get_leaf(ReplayDev R, const double* v, long long n, long long* out_nodes, double* out_pri) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = blockIdx.x * static_cast<long long>(kWarps) + warp;
  if (i >= n) return;
  double p;
  const long long leaf = per_descend_warp(R.tree, 2 * R.cap - 1, v[i], &p);
  if (lane == 0) {
    out_nodes[i] = leaf;
    if (out_pri != nullptr) out_pri[i] = p;
  }
}

// deque position (0 = oldest) -> ring slot
__device__ __forceinline__ long long deque_pos_to_slot(long long pos, long long size, long long dp, long long cap) {
  return (size == cap) ? (dp + pos) % cap : pos;
}

__global__ void __launch_bounds__(kThreads) k_uniform_sample(ReplayDev R, long long B, const long long* idx,
                                                             unsigned long long seed, unsigned long long counter,
                                                             unsigned agent, long long* out_slots, float* out_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = blockIdx.x * static_cast<long long>(kWarps) + warp;
  if (i >= B) return;
  const long long size = R.st->size, dp = R.st->dp;
  const long long pos = (idx != nullptr) ? idx[i] : static_cast<long long>(feistel_perm(i, size, seed, counter, agent));
  const long long slot = deque_pos_to_slot(pos, size, dp, R.cap);
  if (lane == 0) out_slots[i] = slot;
  if (out_rows != nullptr) gather_row_warp(R, slot, out_rows + i * R.row_floats);
}

}  // namespace rmc
