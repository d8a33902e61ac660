// rmc_tree.cuh -- GPU-resident sum tree (reference heap layout, float64 nodes) + min/max
// summaries, stratified prefix-search sampling and batched priority write-back.
//
// Reference semantics restated here:
//   dqn/utils/sum_tree.py:15-32  update  -> leaf store + exact ancestor fix-up (float64 adds of
//                                           float32-exact values are exact in any order, SURVEY
//                                           finding 6, so atomics are bit-identical to the
//                                           reference's sequential propagation)
//   dqn/utils/sum_tree.py:42-61  get_leaf-> per_descend_warp (same `v <= left` / `v -= left`)
//   dqn/utils/sum_tree.py:67-73  max/min -> values of max/min(leaves[:size]) tracked as (value,
//                                           multiplicity) instead of index tracking + O(N) rescans
#pragma once
#include "rmc_device.cuh"

namespace rmc {

__device__ __forceinline__ float finf() { return __int_as_float(0x7f800000); }
// fire-and-forget float64 reduction on a GLOBAL address (the generic atomicAdd carries address-space dispatch)
__device__ __forceinline__ void red_add_f64(double* gptr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(gptr), "d"(v) : "memory");
}

// ---- stratified prefix search: one warp per sample, 4 tree levels per L2 round trip -------------
// All 32 lanes call with identical (v); returns the leaf's tree index on every lane.
// `p`/`pv`: start node and its value.  Lane l holds the look-ahead nodes at linear positions l, l+32, l+64,
// l+96 of the 2+4+...+64 = 126 descendants of p at relative depth 1..6 (depth d starts at position 2^d-2).
__device__ __forceinline__ double lookahead_pick(const double (&val)[4], int pos) {
  const int reg = pos >> 5;   // warp-uniform
  const double x = (reg == 0) ? val[0] : (reg == 1) ? val[1] : (reg == 2) ? val[2] : val[3];
  return __shfl_sync(0xffffffffu, x, pos & 31);
}
__device__ __forceinline__ long long per_descend_from(const double* __restrict__ tree, long long n_nodes, long long p,
                                                      double pv, double v, double* leaf_val) {
  const int lane = threadIdx.x & 31;
  while (2 * p + 1 < n_nodes) {
    double val[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int pos = lane + 32 * q;
      val[q] = 0.0;
      if (pos < 126) {
        const int d = 31 - __clz(pos + 2);
        const long long idx = ((p + 1) << d) - 1 + (pos + 2 - (1 << d));
        if (idx < n_nodes) val[q] = __ldcg(tree + idx);
      }
    }
    long long cur = p;
    int oc = 0;
    bool leaf = false;
#pragma unroll
    for (int dd = 1; dd <= 6; ++dd) {
      const long long left = 2 * cur + 1;
      if (left >= n_nodes) { leaf = true; break; }
      const int lpos = (1 << dd) - 2 + 2 * oc;
      const double lv = lookahead_pick(val, lpos);
      if (v <= lv) { cur = left; oc = 2 * oc; pv = lv; }
      else { v = v - lv; cur = left + 1; oc = 2 * oc + 1; pv = lookahead_pick(val, lpos + 1); }
    }
    p = cur;
    if (leaf) break;
  }
  *leaf_val = pv;
  return p;
}
// TWO descents per warp sharing every memory round trip (the multi-tile launches give a warp two samples: their descents are
// independent, so both samples' 126-node look-ahead windows are requested together).  Per sample exactly the walk above.
__device__ __forceinline__ void per_descend_load(const double* __restrict__ tree, long long n_nodes, long long p, double (&val)[4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int pos = lane + 32 * q;
    val[q] = 0.0;
    if (pos < 126) {
      const int d = 31 - __clz(pos + 2);
      const long long idx = ((p + 1) << d) - 1 + (pos + 2 - (1 << d));
      if (idx < n_nodes) val[q] = __ldcg(tree + idx);
    }
  }
}
__device__ __forceinline__ bool per_descend_walk6(const double (&val)[4], long long n_nodes, long long& p, double& v, double& pv) {
  long long cur = p;
  int oc = 0;
  bool leaf = false;
#pragma unroll
  for (int dd = 1; dd <= 6; ++dd) {
    const long long left = 2 * cur + 1;
    if (left >= n_nodes) { leaf = true; break; }
    const int lpos = (1 << dd) - 2 + 2 * oc;
    const double lv = lookahead_pick(val, lpos);
    if (v <= lv) { cur = left; oc = 2 * oc; pv = lv; }
    else { v = v - lv; cur = left + 1; oc = 2 * oc + 1; pv = lookahead_pick(val, lpos + 1); }
  }
  p = cur;
  return leaf;
}
__device__ __forceinline__ void per_descend_cached2(const double* __restrict__ s_top, int n_top, const double* __restrict__ tree, long long n_nodes,
                                                    double va, double vb, bool has_b, long long* node_a, double* leaf_a, long long* node_b,
                                                    double* leaf_b) {
  long long pa = 0, pb = 0;
  double pva = s_top[0], pvb = s_top[0];
  while (2 * pa + 2 < n_top) {
    const double lv = s_top[2 * pa + 1];
    if (va <= lv) { pa = 2 * pa + 1; pva = lv; }
    else { va = va - lv; pa = 2 * pa + 2; pva = s_top[pa]; }
  }
  while (has_b && 2 * pb + 2 < n_top) {
    const double lv = s_top[2 * pb + 1];
    if (vb <= lv) { pb = 2 * pb + 1; pvb = lv; }
    else { vb = vb - lv; pb = 2 * pb + 2; pvb = s_top[pb]; }
  }
  bool done_a = !(2 * pa + 1 < n_nodes), done_b = !has_b || !(2 * pb + 1 < n_nodes);
  while (!(done_a && done_b)) {
    double xa[4], xb[4];
    if (!done_a) per_descend_load(tree, n_nodes, pa, xa);
    if (!done_b) per_descend_load(tree, n_nodes, pb, xb);
    if (!done_a) { const bool leaf = per_descend_walk6(xa, n_nodes, pa, va, pva); done_a = leaf || !(2 * pa + 1 < n_nodes); }
    if (!done_b) { const bool leaf = per_descend_walk6(xb, n_nodes, pb, vb, pvb); done_b = leaf || !(2 * pb + 1 < n_nodes); }
  }
  *node_a = pa; *leaf_a = pva;
  *node_b = pb; *leaf_b = pvb;
}
__device__ __forceinline__ long long per_descend_warp(const double* __restrict__ tree, long long n_nodes, double v,
                                                      double* leaf_val) {
  return per_descend_from(tree, n_nodes, 0, __ldcg(tree), v, leaf_val);
}
// same walk, the first levels served from a shared-memory copy of tree[0 .. n_top)
__device__ __forceinline__ long long per_descend_cached(const double* __restrict__ s_top, int n_top,
                                                        const double* __restrict__ tree, long long n_nodes, double v,
                                                        double* leaf_val) {
  long long p = 0;
  double pv = s_top[0];
  while (2 * p + 2 < n_top) {          // both children cached
    const double lv = s_top[2 * p + 1];
    if (v <= lv) { p = 2 * p + 1; pv = lv; }
    else { v = v - lv; p = 2 * p + 2; pv = s_top[p]; }
  }
  return per_descend_from(tree, n_nodes, p, pv, v, leaf_val);
}

// value drawn in stratum i (dqn/replay_memory.py:72,80; np.random.uniform(lo,hi) == lo+(hi-lo)*u)
__device__ __forceinline__ double stratum_value(double total, long long Bglobal, long long i, double u) {
  const double seg = total / static_cast<double>(Bglobal);
  const double lo = seg * static_cast<double>(i);
  const double hi = seg * static_cast<double>(i + 1);
  return lo + (hi - lo) * u;   // -fmad=false: mul and add round separately, like numpy
}

// importance weight (dqn/replay_memory.py:76-77,84-86), float64 then cast by the caller
__device__ __forceinline__ double is_weight_max(double size, double total, double min_p, double beta) {
  return pow(size * (min_p / total), -beta);
}
__device__ __forceinline__ double is_weight(double size, double p, double total, double min_p, double beta) {
  return pow(size * (p / total), -beta) / is_weight_max(size, total, min_p, beta);
}

// |td| -> priority (dqn/replay_memory.py:95): float32 min/add, pow evaluated in float64 and rounded
// once to float32 (= correctly rounded powf; numpy's SIMD powf is within 1 ulp of this).
__device__ __forceinline__ float td_to_priority(float abs_td, float eps, float alpha, float pmax) {
  const float x = fminf(abs_td + eps, pmax);
  return static_cast<float>(pow(static_cast<double>(x), static_cast<double>(alpha)));
}

// ---- extreme tracking -------------------------------------------------------------------------
// The reference keeps arg-max / arg-min leaf indices and rescans all leaves when the extreme leaf is
// overwritten (dqn/utils/sum_tree.py:16-28); only the VALUES max/min(leaves[:size]) are observable
// (replay_memory.py:57,76).  Here the state keeps (value, number of leaves holding it) for both
// extremes: a batch update adjusts the counts exactly, and only when a count drops to zero (the last
// leaf holding the extreme was overwritten by a non-extreme value -- probability ~ B*p_min/total per
// step) are the leaves rescanned.

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// leaf store + exact ancestor fix-up for the elected writer of a leaf; returns the old leaf value.
// Nodes below `first_fixed` are NOT touched here: the caller rebuilds those (heavily shared) top levels
// from their children afterwards (sums are exact, SURVEY finding 6, so a rebuild equals propagation).
constexpr int kTopRebuild = 511;   // nodes 0..510 = top 9 levels, children of the last of them: 511..1022
__device__ __forceinline__ double tree_set_leaf(const ReplayDev& R, long long leaf, float p, const double* old_known,
                                                long long first_fixed) {
  const double np = static_cast<double>(p);
  const double old = (old_known != nullptr) ? *old_known : __ldcg(R.tree + leaf);
  double* const tree = R.tree;
  tree[leaf] = np;
  const double delta = np - old;
  if (delta != 0.0) {
    long long n = leaf;
    while (n != 0) {
      n = (n - 1) >> 1;
      if (n < first_fixed) break;
      red_add_f64(tree + n, delta);
    }
  }
  return old;
}

// Rebuild tree[0 .. kTopRebuild) bottom-up in shared memory from tree[kTopRebuild .. 2*kTopRebuild+1)
// (requires 2*kTopRebuild+1 <= n_nodes).  s_buf: 2*kTopRebuild+1 doubles.  All threads of the CTA call.
__device__ void tree_rebuild_top_cta(const ReplayDev& R, double* s_buf) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = kTopRebuild + tid; k < 2 * kTopRebuild + 1; k += nt) s_buf[k] = __ldcg(R.tree + k);
  __syncthreads();
  for (int first = (kTopRebuild - 1) / 2; ; first = (first - 1) / 2) {   // levels 255..510, 127..254, ..., 0
    const int count = first + 1;
    for (int k = first + tid; k < first + count; k += nt) s_buf[k] = s_buf[2 * k + 1] + s_buf[2 * k + 2];
    __syncthreads();
    if (first == 0) break;
  }
  for (int k = tid; k < kTopRebuild; k += nt) R.tree[k] = s_buf[k];
}

// full rescan of leaves[:size] by one CTA (rare path): exact extremes and their multiplicities
__device__ void extremes_rescan_cta(const ReplayDev& R, long long size, float* s_f, int* s_i) {
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const double* leaves = R.tree + (R.cap - 1);
  float mx = 0.f, mn = finf();
  for (long long k = tid; k < size; k += nt) {
    const float p = static_cast<float>(__ldcg(leaves + k));
    mx = fmaxf(mx, p);
    mn = fminf(mn, p);
  }
  mx = warp_max(mx);
  mn = warp_min(mn);
  __syncthreads();
  if (lane == 0) { s_f[warp] = mx; s_f[32 + warp] = mn; }
  __syncthreads();
  mx = 0.f; mn = finf();
  for (int w = 0; w < nw; ++w) { mx = fmaxf(mx, s_f[w]); mn = fminf(mn, s_f[32 + w]); }
  int cx = 0, cn = 0;
  for (long long k = tid; k < size; k += nt) {
    const float p = static_cast<float>(__ldcg(leaves + k));
    cx += (p == mx);
    cn += (p == mn);
  }
  cx = warp_sum(cx);
  cn = warp_sum(cn);
  __syncthreads();
  if (lane == 0) { s_i[warp] = cx; s_i[32 + warp] = cn; }
  __syncthreads();
  if (tid == 0) {
    long long tx = 0, tn = 0;
    for (int w = 0; w < nw; ++w) { tx += s_i[w]; tn += s_i[32 + w]; }
    R.st->max_p = mx; R.st->min_p = mn; R.st->cnt_max = tx; R.st->cnt_min = tn;
  }
}

// Whole write-back by ONE CTA (n <= kTreeCtaMax): duplicates -> last in batch order wins
// (dqn/replay_memory.py:97-98 applies updates sequentially).  old_size/new_size: number of valid
// leaves before / after (they differ for pushes).
__device__ void tree_update_cta(const ReplayDev& R, const long long* __restrict__ nodes,
                                const float* __restrict__ pri, long long n, long long old_size, long long new_size,
                                bool stamps_done, const double* __restrict__ old_vals, double* s_top_buf,
                                unsigned long long* dbg = nullptr) {
#define RMC_TSTAMP(k) do { if (dbg != nullptr && threadIdx.x == 0) dbg[k] = global_timer_ns(); } while (0)
  __shared__ float s_f[64];
  __shared__ int s_i[64];
  __shared__ float s_ext[4];   // M0, m0, M1, m1
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const long long first_leaf = R.cap - 1;
  // big trees: per-leaf atomics only below the top 9 levels, which are rebuilt afterwards
  const bool rebuild_top = (s_top_buf != nullptr) && (2 * R.cap - 1 >= 2 * kTopRebuild + 1);
  const long long first_fixed = rebuild_top ? kTopRebuild : 0;
  // (index lists that come from outside the launch may hold anything: entries outside the leaf range are skipped and counted)
  if (!stamps_done)
    for (long long i = tid; i < n; i += nt) {
      const long long di = nodes[i] - first_leaf;
      if (di >= 0 && di < R.cap) atomicMax(R.stamps + di, static_cast<int>(i + 1));
      else atomicAdd(&R.st->bad_nodes, 1);
    }
  __syncthreads();
  RMC_TSTAMP(9);
  // pass 1: elected writers apply; remember the overwritten value (-1: leaf was outside the old domain,
  // -2: not the elected writer); batch extremes of the NEW values
  float bmax = 0.f, bmin = finf();
  for (long long i = tid; i < n; i += nt) {
    const long long leaf = nodes[i];
    const long long di = leaf - first_leaf;
    int* st = R.stamps + di;
    float oldv = -2.f;
    if (di >= 0 && di < R.cap && __ldcg(st) == static_cast<int>(i + 1)) {
      *st = 0;
      const float p = pri[i];
      const double old = tree_set_leaf(R, leaf, p, old_vals ? old_vals + i : nullptr, first_fixed);
      oldv = (di < old_size) ? static_cast<float>(old) : -1.f;
      bmax = fmaxf(bmax, p);
      bmin = fminf(bmin, p);
    }
    R.scratch_old[i] = oldv;
  }
  bmax = warp_max(bmax);
  bmin = warp_min(bmin);
  if (lane == 0) { s_f[warp] = bmax; s_f[32 + warp] = bmin; }
  __syncthreads();
  RMC_TSTAMP(10);
  if (tid == 0) {
    float M = 0.f, m = finf();
    for (int w = 0; w < nw; ++w) { M = fmaxf(M, s_f[w]); m = fminf(m, s_f[32 + w]); }
    const float M0 = (old_size > 0) ? R.st->max_p : 0.f;
    const float m0 = (old_size > 0) ? R.st->min_p : finf();
    s_ext[0] = M0; s_ext[1] = m0; s_ext[2] = fmaxf(M0, M); s_ext[3] = fminf(m0, m);
  }
  __syncthreads();
  const float M0 = s_ext[0], m0 = s_ext[1], M1 = s_ext[2], m1 = s_ext[3];
  // pass 2: multiplicities
  int a = 0, b = 0, c = 0, d = 0;   // new==M1, old==M0, new==m1, old==m0
  for (long long i = tid; i < n; i += nt) {
    const float oldv = R.scratch_old[i];
    if (oldv != -2.f) {
      const float p = pri[i];
      a += (p == M1);
      c += (p == m1);
      if (oldv >= 0.f) { b += (oldv == M0); d += (oldv == m0); }
    }
  }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
  __syncthreads();
  if (lane == 0) { s_i[warp] = a; s_i[8 + warp] = b; s_i[16 + warp] = c; s_i[24 + warp] = d; }
  __syncthreads();
  if (tid == 0) {
    long long ta = 0, tb = 0, tc = 0, td = 0;
    for (int w = 0; w < nw; ++w) { ta += s_i[w]; tb += s_i[8 + w]; tc += s_i[16 + w]; td += s_i[24 + w]; }
    const long long c0M = (old_size > 0) ? R.st->cnt_max : 0, c0m = (old_size > 0) ? R.st->cnt_min : 0;
    const long long cM = (M1 > M0) ? ta : c0M - tb + ta;
    const long long cm = (m1 < m0) ? tc : c0m - td + tc;
    R.st->max_p = M1; R.st->min_p = m1; R.st->cnt_max = cM; R.st->cnt_min = cm;
    s_i[63] = (new_size > 0 && (cM <= 0 || cm <= 0)) ? 1 : 0;
  }
  __syncthreads();
  if (s_i[63]) extremes_rescan_cta(R, new_size, s_f, s_i);
  RMC_TSTAMP(11);
  if (rebuild_top) {
    __threadfence();     // this thread's reductions on the lower levels have been performed
    __syncthreads();
    RMC_TSTAMP(12);
    tree_rebuild_top_cta(R, s_top_buf);
  }
#undef RMC_TSTAMP
}

// ---- write-back by a TEAM of kTreeTeam CTAs (learner step, big trees) ---------------------------------
// Spatial partition: member t owns the leaves below the t-th node of depth 3 (heap indices 7..14), so
// writer elections and all fix-up atomics below depth 3 are private to one member; every member rebuilds
// levels 8..3 under its own depth-3 node, and whichever member arrives last (no spinning) finishes
// levels 2..0 and combines the extremes.
__device__ __forceinline__ int depth3_owner(long long node) {
  const unsigned long long n1 = static_cast<unsigned long long>(node) + 1ull;
  const int depth = 63 - __clzll(n1);            // root = depth 0
  return static_cast<int>((n1 >> (depth - 3)) - 8ull);   // valid for depth >= 3
}

__device__ void tree_update_team(const ReplayDev& R, const long long* __restrict__ nodes, const float* __restrict__ abs_td, int td_stride,
                                 float* __restrict__ pri_out, long long n, long long size, const double* __restrict__ old_vals,
                                 float eps, float alpha, float pmax, int member, bool sorted, double* s_top_buf, unsigned long long* dbg) {
  __shared__ float s_f[64];
  __shared__ int s_i[64];
  __shared__ float s_loc[2];
  __shared__ int s_last;
#define RMC_TSTAMP(k) do { if (dbg != nullptr && threadIdx.x == 0) dbg[k] = global_timer_ns(); } while (0)
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const long long first_leaf = R.cap - 1;
  const float M0 = R.st->max_p, m0 = R.st->min_p;      // extremes before this batch (size > 0 here)
  // Election of the writer of a leaf that occurs several times in the batch (the reference applies the updates in batch
  // order, so the last one wins).  `sorted`: the leaves were drawn by THIS launch's stratified sampler -- stratum values
  // are non-decreasing in the sample index (v_i <= seg*(i+1) <= v_{i+1}: hi - lo is exact by Sterbenz, rounding is monotone)
  // and the descent is monotone in v (the drawn leaves move left to right through the tree; their heap INDEX need not grow
  // when the capacity is not a power of two), so equal leaves are adjacent and sample i is the writer iff leaf[i+1] != leaf[i].
  // No stamps, no atomics, no extra round trip.  Otherwise (leaves from an earlier call): atomicMax stamps, two passes.
  if (!sorted) {
    for (long long i = tid; i < n; i += nt) {
      const long long leaf = __ldcg(nodes + i);
      if (leaf >= first_leaf && leaf - first_leaf < R.cap && depth3_owner(leaf) == member) {
        pri_out[i] = td_to_priority(__ldcg(abs_td + i * td_stride), eps, alpha, pmax);
        atomicMax(R.stamps + (leaf - first_leaf), static_cast<int>(i + 1));
      }
    }
    __syncthreads();
  }
  RMC_TSTAMP(9);
  // pass 1: elected writers apply (atomics only below the top 9 levels); local extremes of the new values.
  // The thread's first sample stays in registers for pass 2 (the whole batch when n <= blockDim); later ones go through
  // the L2-resident scratch.
  float bmax = 0.f, bmin = finf();
  float r_old = -3.f, r_p = 0.f;                       // -3: not mine, -2: mine but not the writer, else the overwritten value
  for (long long i = tid; i < n; i += nt) {
    const long long leaf = __ldcg(nodes + i);
    float oldv = -3.f, p = 0.f;
    if (leaf >= first_leaf && leaf - first_leaf < R.cap && depth3_owner(leaf) == member) {
      oldv = -2.f;
      bool writer;
      if (sorted) {
        const long long next = (i + 1 < n) ? __ldcg(nodes + i + 1) : -1;
        p = td_to_priority(__ldcg(abs_td + i * td_stride), eps, alpha, pmax);
        pri_out[i] = p;
        writer = next != leaf;
      } else {
        int* st = R.stamps + (leaf - first_leaf);
        writer = __ldcg(st) == static_cast<int>(i + 1);
        if (writer) *st = 0;
        p = pri_out[i];
      }
      if (writer) {
        oldv = static_cast<float>(tree_set_leaf(R, leaf, p, old_vals ? old_vals + i : nullptr, kTopRebuild));
        bmax = fmaxf(bmax, p);
        bmin = fminf(bmin, p);
      }
      if (i >= nt) R.scratch_old[i] = oldv;
    }
    if (i < nt) { r_old = oldv; r_p = p; }
  }
  bmax = warp_max(bmax);
  bmin = warp_min(bmin);
  if (lane == 0) { s_f[warp] = bmax; s_f[32 + warp] = bmin; }
  __syncthreads();
  if (tid == 0) {
    float M = 0.f, m = finf();
    for (int w = 0; w < nw; ++w) { M = fmaxf(M, s_f[w]); m = fminf(m, s_f[32 + w]); }
    s_loc[0] = M; s_loc[1] = m;
  }
  __syncthreads();
  RMC_TSTAMP(10);
  const float Mt = s_loc[0], mt = s_loc[1];
  int a = 0, b = 0, c = 0, d = 0;   // new==Mt, old==M0, new==mt, old==m0
  if (r_old > -2.f) {
    a += (r_p == Mt); c += (r_p == mt);
    b += (r_old == M0); d += (r_old == m0);
  }
  for (long long i = tid + nt; i < n; i += nt) {
    const long long leaf = nodes[i];
    if (leaf >= first_leaf && leaf - first_leaf < R.cap && depth3_owner(leaf) == member) {
      const float oldv = R.scratch_old[i];
      if (oldv != -2.f) {
        const float p = pri_out[i];
        a += (p == Mt); c += (p == mt);
        b += (oldv == M0); d += (oldv == m0);
      }
    }
  }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
  if (lane == 0) { s_i[warp] = a; s_i[8 + warp] = b; s_i[16 + warp] = c; s_i[24 + warp] = d; }
  __threadfence();          // this thread's leaf stores / reductions are performed before the local rebuild reads them
  __syncthreads();
  // Local part of the top rebuild: levels 8..3 under this member's depth-3 node depend only on its own level-9
  // nodes (64 of them), which no other member touches -- done here, off the last arriver's tail.  Warp 1 meanwhile
  // posts this member's share of the extremes.
  if (warp == 0) {
    const long long d3 = 7 + member;                              // heap index of the member's depth-3 node
    const long long first9 = ((d3 + 1) << 6) - 1;                 // its 64 descendants at depth 9
    double v0 = __ldcg(R.tree + first9 + 2 * lane), v1 = __ldcg(R.tree + first9 + 2 * lane + 1);
    double sum = v0 + v1;                                         // depth 8: 32 nodes, lane l <-> node l
    long long first = ((d3 + 1) << 5) - 1;
    R.tree[first + lane] = sum;
#pragma unroll
    for (int lvl = 7, width = 16; lvl >= 3; --lvl, width >>= 1) {  // depth 7 (16 nodes) ... depth 3 (1 node)
      const double other = __shfl_down_sync(0xffffffffu, sum, 1);   // partner = next lane's value at the level below
      const double pair = sum + other;                            // valid on even lanes of the previous level
      // compact: node j of this level = lanes 2j,2j+1 of the level below -> gather into lane j
      sum = __shfl_sync(0xffffffffu, pair, 2 * lane);
      first = ((d3 + 1) << (lvl - 3)) - 1;
      if (lane < width) R.tree[first + lane] = sum;
    }
  } else if (tid == 32) {
    int ta = 0, tb = 0, tc = 0, td = 0;
    for (int w = 0; w < nw; ++w) { ta += s_i[w]; tb += s_i[8 + w]; tc += s_i[16 + w]; td += s_i[24 + w]; }
    TeamPart tp;
    tp.bmax = Mt; tp.bmin = mt; tp.cnt_bmax = ta; tp.cnt_bmin = tc; tp.old_eq_max = tb; tp.old_eq_min = td; tp.pad[0] = tp.pad[1] = 0;
    R.team_part[member] = tp;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();        // the rebuilt levels and the TeamPart (ordered before this by the barrier) precede the arrival
    const unsigned prev = atomicAdd(R.team_ctr, 1u);
    s_last = (prev == kTreeTeam - 1) ? 1 : 0;
    if (s_last) { *R.team_ctr = 0u; __threadfence(); }
  }
  __syncthreads();
  RMC_TSTAMP(11);
  if (!s_last) return;
  // ---- last member: combine the extremes, rebuild the top of the tree
  if (tid == 0) {
    float Mb = 0.f, mb = finf();
    TeamPart tp[kTreeTeam];
    for (int t = 0; t < kTreeTeam; ++t) {
      const int4 lo = __ldcg(reinterpret_cast<const int4*>(R.team_part + t));
      const int4 hi = __ldcg(reinterpret_cast<const int4*>(R.team_part + t) + 1);
      tp[t].bmax = __int_as_float(lo.x); tp[t].bmin = __int_as_float(lo.y); tp[t].cnt_bmax = lo.z; tp[t].cnt_bmin = lo.w;
      tp[t].old_eq_max = hi.x; tp[t].old_eq_min = hi.y;
      Mb = fmaxf(Mb, tp[t].bmax); mb = fminf(mb, tp[t].bmin);
    }
    long long ta = 0, tb = 0, tc = 0, td = 0;
    for (int t = 0; t < kTreeTeam; ++t) {
      if (tp[t].bmax == Mb) ta += tp[t].cnt_bmax;
      if (tp[t].bmin == mb) tc += tp[t].cnt_bmin;
      tb += tp[t].old_eq_max; td += tp[t].old_eq_min;
    }
    const float M1 = fmaxf(M0, Mb), m1 = fminf(m0, mb);
    const long long cM = (Mb > M0) ? ta : R.st->cnt_max - tb + ((Mb == M0) ? ta : 0);
    const long long cm = (mb < m0) ? tc : R.st->cnt_min - td + ((mb == m0) ? tc : 0);
    R.st->max_p = M1; R.st->min_p = m1; R.st->cnt_max = cM; R.st->cnt_min = cm;
    s_i[63] = (cM <= 0 || cm <= 0) ? 1 : 0;
  }
  if (tid == 32) {     // meanwhile: levels 2..0 from the eight depth-3 nodes the members have just rebuilt
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(R.tree + 7 + k);
    const double l2[4] = {v[0] + v[1], v[2] + v[3], v[4] + v[5], v[6] + v[7]};
    const double l1[2] = {l2[0] + l2[1], l2[2] + l2[3]};
    R.tree[3] = l2[0]; R.tree[4] = l2[1]; R.tree[5] = l2[2]; R.tree[6] = l2[3];
    R.tree[1] = l1[0]; R.tree[2] = l1[1];
    R.tree[0] = l1[0] + l1[1];
  }
  __syncthreads();
  if (s_i[63]) extremes_rescan_cta(R, size, s_f, s_i);
  RMC_TSTAMP(12);
#undef RMC_TSTAMP
}

// ---- standalone kernels -------------------------------------------------------------------
// one CTA: rmc_per_update for n <= kTreeCtaMax; optional |td| -> priority conversion first
__global__ void __launch_bounds__(kThreads) k_tree_update_small(ReplayDev R, const long long* nodes, const float* pri_in,
                                                                const float* abs_td, float* pri_out, long long n,
                                                                float eps, float alpha, float pmax) {
  pdl_enter();
  const SpanScope span_(SPAN_TREE_SMALL);
  const float* pri = pri_in;
  if (abs_td != nullptr) {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) pri_out[i] = td_to_priority(abs_td[i], eps, alpha, pmax);
    __syncthreads();
    pri = pri_out;
  }
  __shared__ double s_top_buf[2 * kTopRebuild + 1];
  const long long size = R.st->size;
  tree_update_cta(R, nodes, pri, n, size, size, false, nullptr, s_top_buf);
}

// grid-wide variants for large batches
__global__ void k_td_to_pri(const float* abs_td, float* pri, long long n, float eps, float alpha, float pmax) {
  pdl_enter();
  const SpanScope span_(SPAN_TD_TO_PRI);
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) pri[i] = td_to_priority(abs_td[i], eps, alpha, pmax);
}
__global__ void k_tree_stamp(ReplayDev R, const long long* nodes, long long n) {
  pdl_enter();
  const SpanScope span_(SPAN_TREE_STAMP);
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) {
    const long long di = nodes[i] - (R.cap - 1);
    if (di >= 0 && di < R.cap) atomicMax(R.stamps + di, static_cast<int>(i + 1));
    else atomicAdd(&R.st->bad_nodes, 1);
  }
}
// Leaf stores + float64 reductions on the ancestors at heap index >= first_fixed only (the contended top of the
// tree is rebuilt afterwards by k_tree_rebuild_top; sums of f32-exact values are exact in any order, SURVEY finding 6).
__global__ void k_tree_apply(ReplayDev R, const long long* nodes, const float* pri, long long n, long long first_fixed) {
  pdl_enter();
  const SpanScope span_(SPAN_TREE_APPLY);
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) {
    const long long leaf = nodes[i];
    const long long di = leaf - (R.cap - 1);
    int* st = R.stamps + di;
    if (di >= 0 && di < R.cap && __ldcg(st) == static_cast<int>(i + 1)) {
      *st = 0;
      tree_set_leaf(R, leaf, pri[i], nullptr, first_fixed);
    }
  }
}
// ONE CTA: nodes [0, F) rebuilt bottom-up in shared memory from nodes [F, 2F+1)  (F = 2^L - 1 <= kTopLargeMax)
constexpr int kTopLargeMax = 2047;
__global__ void __launch_bounds__(1024) k_tree_rebuild_top(ReplayDev R, int F) {
  pdl_enter();
  const SpanScope span_(SPAN_TREE_TOP);
  __shared__ double s_buf[2 * kTopLargeMax + 1];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = F + tid; k < 2 * F + 1; k += nt) s_buf[k] = __ldcg(R.tree + k);
  __syncthreads();
  for (int first = (F - 1) / 2;; first = (first - 1) / 2) {
    for (int k = first + tid; k < 2 * first + 1; k += nt) s_buf[k] = s_buf[2 * k + 1] + s_buf[2 * k + 2];
    __syncthreads();
    if (first == 0) break;
  }
  for (int k = tid; k < F; k += nt) R.tree[k] = s_buf[k];
}
// grid-wide rescan of the extremes (bulk paths) in ONE pass: every thread keeps (max, #max, min, #min) of its leaves,
// blocks publish their merged tuple, the last block to arrive merges the block tuples into the replay state.
struct __align__(16) ExtTuple { float mx, mn; int cx, cn; };
__device__ __forceinline__ void ext_merge(ExtTuple& a, float mx, int cx, float mn, int cn) {
  if (mx > a.mx) { a.mx = mx; a.cx = cx; } else if (mx == a.mx) a.cx += cx;
  if (mn < a.mn) { a.mn = mn; a.cn = cn; } else if (mn == a.mn) a.cn += cn;
}
__device__ __forceinline__ void ext_warp_merge(ExtTuple& a) {
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) {
    const float mx = __shfl_down_sync(0xffffffffu, a.mx, sh), mn = __shfl_down_sync(0xffffffffu, a.mn, sh);
    const int cx = __shfl_down_sync(0xffffffffu, a.cx, sh), cn = __shfl_down_sync(0xffffffffu, a.cn, sh);
    ext_merge(a, mx, cx, mn, cn);
  }
}
constexpr int kExtBlocks = 592;
__global__ void __launch_bounds__(256) k_extremes_scan(ReplayDev R, ExtTuple* parts, unsigned* arrive) {
  pdl_enter();
  const SpanScope span_(SPAN_EXTREMES);
  __shared__ ExtTuple s_w[8];
  __shared__ bool s_last;
  const long long size = R.st->size;
  const double* leaves = R.tree + (R.cap - 1);
  ExtTuple a{0.f, finf(), 0, 0};
  {   // 8 leaves per thread in flight at a time (a load -> merge loop is one memory round trip per leaf: 8.7 us for 1M leaves)
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long k0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k0 < size; k0 += 8 * stride) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (k0 + q * stride < size) ? __ldcg(leaves + k0 + q * stride) : -1.0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (v[q] >= 0.0) { const float p = static_cast<float>(v[q]); ext_merge(a, p, 1, p, 1); }
    }
  }
  ext_warp_merge(a);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    ExtTuple b = s_w[0];
    for (int w = 1; w < 8; ++w) ext_merge(b, s_w[w].mx, s_w[w].cx, s_w[w].mn, s_w[w].cn);
    parts[blockIdx.x] = b;
    __threadfence();
    s_last = (atomicAdd(arrive, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  ExtTuple b{0.f, finf(), 0, 0};
  for (int k = threadIdx.x; k < static_cast<int>(gridDim.x); k += blockDim.x) {
    const float4 raw = __ldcg(reinterpret_cast<const float4*>(parts + k));
    ext_merge(b, raw.x, __float_as_int(raw.z), raw.y, __float_as_int(raw.w));
  }
  ext_warp_merge(b);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    ExtTuple c = s_w[0];
    for (int w = 1; w < 8; ++w) ext_merge(c, s_w[w].mx, s_w[w].cx, s_w[w].mn, s_w[w].cn);
    R.st->max_p = c.mx; R.st->min_p = c.mn; R.st->cnt_max = c.cx; R.st->cnt_min = c.cn;
    *arrive = 0u;
  }
}

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
  __shared__ double s_top_buf[2 * kTopRebuild + 1];
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
    tree_update_cta(R, scratch_nodes, scratch_pri, n, s_size, new_size, false, nullptr, s_top_buf);
  }
  __syncthreads();
  if (tid == 0) {
    R.st->dp = (dp + n) % R.cap;
    R.st->size = new_size;
  }
}

// n <= 32 rows (the per-env-step push of the trainer), ONE warp: ring write, leaf <- max_priority with direct
// ancestor atomics on every level, exact extremes bookkeeping.  The body is a device function so that the ensemble's
// one-launch push (k_push_tiny_group, rmc_mlp.cuh) runs exactly the same code per member.
__device__ __forceinline__ void push_tiny_cta(const ReplayDev& R, const float* __restrict__ rows, int n, float pmax, float* s_f, int* s_i) {
  const int lane = threadIdx.x;       // blockDim.x == 32
  const long long dp = R.st->dp, size = R.st->size;
  const long long new_size = min(size + static_cast<long long>(n), R.cap);
  const float M0 = (size > 0) ? R.st->max_p : 0.f, m0 = (size > 0) ? R.st->min_p : finf();
  const float p = (M0 == 0.f) ? pmax : M0;
  const int rf = R.row_floats;
  for (int t = lane; t < n * rf; t += 32) R.ring[((dp + t / rf) % R.cap) * rf + (t % rf)] = rows[t];
  if (R.prioritized) {
    int b = 0, d = 0;
    if (lane < n) {
      const long long slot = (dp + lane) % R.cap;
      const double old = tree_set_leaf(R, slot + R.cap - 1, p, nullptr, 0);
      if (slot < size) { b = (static_cast<float>(old) == M0); d = (static_cast<float>(old) == m0); }
    }
    b = warp_sum(b);
    d = warp_sum(d);
    if (lane == 0) {
      const float M1 = fmaxf(M0, p), m1 = fminf(m0, p);
      const long long c0M = (size > 0) ? R.st->cnt_max : 0, c0m = (size > 0) ? R.st->cnt_min : 0;
      const long long cM = (p > M0) ? n : c0M - b + ((p == M0) ? n : 0);
      const long long cm = (p < m0) ? n : c0m - d + ((p == m0) ? n : 0);
      R.st->max_p = M1; R.st->min_p = m1; R.st->cnt_max = cM; R.st->cnt_min = cm;
      s_i[63] = (cM <= 0 || cm <= 0) ? 1 : 0;
    }
    __syncwarp();
    if (s_i[63]) { __threadfence(); extremes_rescan_cta(R, new_size, s_f, s_i); }
  }
  __syncwarp();
  if (lane == 0) {
    R.st->dp = (dp + n) % R.cap;
    R.st->size = new_size;
  }
}
template <bool kArgs>
__global__ void __launch_bounds__(32) k_push_tiny(ReplayDev R, const float* __restrict__ rows_ptr, TinyRows rows_arg, int n, float pmax) {
  __shared__ float s_f[64];
  __shared__ int s_i[64];
  push_tiny_cta(R, kArgs ? rows_arg.v : rows_ptr, n, pmax, s_f, s_i);   // kArgs: the rows travelled inside the launch itself
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
// split gather: issue the loads (<= 3 floats per lane, rows are <= 68 floats), do other work, then store
struct RowRegs { float v[3]; };
__device__ __forceinline__ RowRegs gather_row_load(const ReplayDev& R, long long slot) {
  const int lane = threadIdx.x & 31;
  const float* src = R.ring + slot * R.row_floats;
  RowRegs r;
#pragma unroll
  for (int q = 0; q < 3; ++q) r.v[q] = (lane + 32 * q < R.row_floats) ? __ldcg(src + lane + 32 * q) : 0.f;
  return r;
}
__device__ __forceinline__ void gather_row_store_smem(const ReplayDev& R, const RowRegs& r, float* __restrict__ sdst) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 3; ++q)
    if (lane + 32 * q < R.row_floats) sdst[lane + 32 * q] = r.v[q];
}
__device__ __forceinline__ void gather_row_store(const ReplayDev& R, const RowRegs& r, float* __restrict__ dst, float* __restrict__ sdst) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 3; ++q)
    if (lane + 32 * q < R.row_floats) { dst[lane + 32 * q] = r.v[q]; sdst[lane + 32 * q] = r.v[q]; }
}
__device__ __forceinline__ void gather_row_warp(const ReplayDev& R, long long slot, float* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const float* src = R.ring + slot * R.row_floats;
  for (int c = lane; c < R.row_floats; c += 32) dst[c] = __ldcg(src + c);
}

__global__ void __launch_bounds__(kThreads) k_per_sample(ReplayDev R, long long B, long long Bglobal, long long shard_off,
                                                         double beta, const double* u, unsigned long long seed,
                                                         unsigned long long counter, unsigned agent, long long* out_nodes,
                                                         float* out_w, float* out_rows, double* out_leaf_p) {
  pdl_enter();
  const SpanScope span_(SPAN_SAMPLE);
  __shared__ double s_max_w;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = blockIdx.x * static_cast<long long>(kWarps) + warp;
  const long long n_nodes = 2 * R.cap - 1;
  const double total = __ldcg(R.tree);
  const long long size = R.st->size;
  // the max IS weight is common to the batch: once per CTA (last thread), concurrently with the descents
  if (out_w != nullptr && threadIdx.x == kThreads - 1)
    s_max_w = is_weight_max(static_cast<double>(size), total, static_cast<double>(R.st->min_p), beta);
  long long leaf = 0;
  double p = 0.0, numer = 1.0;
  if (i < B) {
    const double ui = (u != nullptr) ? u[i] : philox_uniform(seed, counter, agent, static_cast<uint32_t>(shard_off + i));
    const double v = stratum_value(total, Bglobal, shard_off + i, ui);
    leaf = per_descend_warp(R.tree, n_nodes, v, &p);
    if (out_rows != nullptr) gather_row_warp(R, leaf - (R.cap - 1), out_rows + i * R.row_floats);
    if (out_w != nullptr && lane == 0) numer = pow(static_cast<double>(size) * (p / total), -beta);
  }
  __syncthreads();
  if (i < B && lane == 0) {
    out_nodes[i] = leaf;
    if (out_w != nullptr) out_w[i] = static_cast<float>(numer / s_max_w);
    if (out_leaf_p != nullptr) out_leaf_p[i] = p;
  }
}

// Large batches: ONE LANE per sample.  Stratified samples are sorted, so the 32 descents of a warp share their first
// ~log2(B) - 5 nodes (one broadcast L1 hit each) and the whole batch is in flight at once (2048 descents per SM);
// the comparison sequence per sample is still exactly sum_tree.py:53-57.  Rows are then gathered cooperatively:
// 8 lanes fetch one 128-byte row with float4 loads, the warp's 32 output rows are written contiguously.
constexpr int kLaneThreads = 128;
__global__ void __launch_bounds__(kLaneThreads) k_per_sample_lane(ReplayDev R, long long B, long long Bglobal, long long shard_off,
                                                              double beta, const double* __restrict__ u, unsigned long long seed,
                                                              unsigned long long counter, unsigned agent, long long* __restrict__ out_nodes,
                                                              float* __restrict__ out_w, float* __restrict__ out_rows,
                                                              double* __restrict__ out_leaf_p) {
  pdl_enter();
  const SpanScope span_(SPAN_SAMPLE);
  __shared__ double s_max_w;
  const int lane = threadIdx.x & 31;
  const long long i = blockIdx.x * static_cast<long long>(kLaneThreads) + threadIdx.x;
  const long long n_nodes = 2 * R.cap - 1;
  const double* __restrict__ tree = R.tree;
  const double total = __ldg(tree);
  const long long size = R.st->size;
  if (out_w != nullptr && threadIdx.x == kLaneThreads - 1)
    s_max_w = is_weight_max(static_cast<double>(size), total, static_cast<double>(R.st->min_p), beta);
  long long leaf = R.cap - 1;
  double pv = 0.0, numer = 1.0;
  if (i < B) {
    const double ui = (u != nullptr) ? u[i] : philox_uniform(seed, counter, agent, static_cast<uint32_t>(shard_off + i));
    double v = stratum_value(total, Bglobal, shard_off + i, ui);
    long long p = 0;
    while (2 * p + 1 < n_nodes) {
      const long long l = 2 * p + 1;
      const double lv = __ldg(tree + l);
      if (v <= lv) { p = l; } else { v = v - lv; p = l + 1; }
    }
    leaf = p;
    pv = __ldg(tree + p);
    if (out_w != nullptr) numer = pow(static_cast<double>(size) * (pv / total), -beta);
  }
  if (out_rows != nullptr) {
    const int rf4 = R.row_floats >> 2;
    const long long i0 = i - lane;                       // first sample of this warp
    const long long slot = leaf - (R.cap - 1);
    const float4* __restrict__ ring4 = reinterpret_cast<const float4*>(R.ring);
    float4* __restrict__ dst4 = reinterpret_cast<float4*>(out_rows + i0 * R.row_floats);
    for (int t = lane; t < 32 * rf4; t += 32) {
      const int row = t / rf4, c = t - row * rf4;
      const long long sl = __shfl_sync(0xffffffffu, slot, row);
      if (i0 + row < B) dst4[t] = __ldcg(ring4 + sl * rf4 + c);
    }
  }
  __syncthreads();
  if (i < B) {
    out_nodes[i] = leaf;
    if (out_w != nullptr) out_w[i] = static_cast<float>(numer / s_max_w);
    if (out_leaf_p != nullptr) out_leaf_p[i] = pv;
  }
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

__global__ void __launch_bounds__(kThreads) k_uniform_sample(ReplayDev R, long long B, long long shard_off, const long long* idx,
                                                             unsigned long long seed, unsigned long long counter,
                                                             unsigned agent, long long* out_slots, float* out_rows) {
  pdl_enter();
  const SpanScope span_(SPAN_UNIFORM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = blockIdx.x * static_cast<long long>(kWarps) + warp;
  if (i >= B) return;
  const long long size = R.st->size, dp = R.st->dp;
  const long long pos = (idx != nullptr) ? idx[i] : static_cast<long long>(feistel_perm(shard_off + i, size, seed, counter, agent));
  const long long slot = deque_pos_to_slot(pos, size, dp, R.cap);
  if (lane == 0) out_slots[i] = slot;
  if (out_rows != nullptr) gather_row_warp(R, slot, out_rows + i * R.row_floats);
}

}  // namespace rmc
