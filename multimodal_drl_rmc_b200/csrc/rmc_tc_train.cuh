// rmc_tc_train.cuh -- tensor-core (tcgen05 / TMEM, bf16 operands, fp32 accumulation) learner step for the DENSE
// large-batch config (BASELINE configs[4], B = 65,536).  Stated looser bound: gradients within 1e-1 max-norm
// relative of the exact fp32 path (bf16 operand rounding); the fp32 FFMA kernel k_learner_step stays the parity
// path and the default.
//
// Pipeline (all on one stream; the minibatch is drawn by the grid-wide sampler first):
//   k_tc_fwd3           Q_online(s'), Q_target(s'), Q_online(s) as raw heads in ONE launch (rmc_tc.cuh); the s pass leaves the
//                       shared-memory operand images of X, H1, H2 in HBM with bulk stores
//   k_tc_td             TD target, |td|, Huber, loss partials, head deltas DH (bf16)                    [CUDA cores]
//   k_tc_bwd_fused      per 128-row tile: bulk-load X | H1 | H2, dh2 = DH.Wh -> mask -> DZ2, dz1 = DZ2.W2 -> mask -> DZ1,
//                       dW2 += H1^T.DZ2, dW0 += DZ1^T.X, dWh += H2^T.DH (accumulators stay in TMEM across the CTA's tiles);
//                       one partial gradient blob per CTA
//   k_tc_reduce_adam    fixed-order sum of the partials -> gradients -> Adam (+ Polyak) -> refreshed bf16 operand images, loss
#pragma once
#include "rmc_mlp.cuh"
#include "rmc_tc.cuh"

namespace rmc {

// packed backward operands (bf16, canonical K-major): Wh^T as [N = 128 j][K = 16 a], W2 as [N = 256 k][K = 128 j]
constexpr int kTcBwdOffWhT = 0;
constexpr int kTcBwdOffW2 = kH2 * kTcNH;
constexpr int kTcBwdElems = kTcBwdOffW2 + kH1 * kH2;
constexpr int kTcBwdBytes = kTcBwdElems * 2;          // 69,632 B

__global__ void k_tc_pack_bwd(const float* __restrict__ blob, NetLayout L, __nv_bfloat16* __restrict__ out) {
  pdl_enter();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kH2 * kTcNH) {                                 // (j, a) -> Wh[a][j]
    const int j = t / kTcNH, a = t % kTcNH;
    out[kTcBwdOffWhT + tc_off(j, a, kTcNH)] = __float2bfloat16_rn(a < L.NH ? blob[L.off_wh + a * kH2 + j] : 0.f);
  }
  if (t < kH1 * kH2) {                                   // (k, j) -> W2^T[k][j]
    const int k = t / kH2, j = t % kH2;
    out[kTcBwdOffW2 + tc_off(k, j, kH2)] = __float2bfloat16_rn(blob[L.off_w2t + k * kW2LD + j]);
  }
}

struct TcTrainBufs {
  float *heads_n, *heads_t, *heads_s;                    // [B][16] raw heads: online(s'), target(s'), online(s)
  __nv_bfloat16 *Xb, *H1b, *H2b;                         // per 128-row tile: the forward kernel's operand images (tc_off layout)
  __nv_bfloat16* DHb;                                    // head deltas, row-major bf16 [B][16]
  float* partials;                                       // [n_ctas][L.total]
  int n_part;
};

// The Adam / Polyak kernels of this mode refresh the bf16 operand images themselves, element by element, as they write
// the fp32 master weights (no separate pack kernels on the step's critical path).  Null pointers = image not kept.
struct TcPackOut {
  unsigned char* fwd_online;       // k_tc_pack image of the online net (bf16 operands | fp32 biases)
  __nv_bfloat16* bwd_online;       // k_tc_pack_bwd image of the online net
  unsigned char* fwd_target;       // k_tc_pack image of the target net
};
__device__ __forceinline__ void tc_pack_param(const NetLayout& L, int pi, float val, unsigned char* fwd, __nv_bfloat16* bwd) {
  __nv_bfloat16* w = reinterpret_cast<__nv_bfloat16*>(fwd);
  float* bias = (fwd != nullptr) ? reinterpret_cast<float*>(fwd + kTcBf16Elems * 2) : nullptr;
  const __nv_bfloat16 hv = __float2bfloat16_rn(val);
  int r;
  if ((r = pi - L.off_w0t) >= 0 && r < L.D * kH1) {
    const int d = r / kH1, i = r - d * kH1;
    if (w) w[kTcOffW0 + tc_off(i, d, kTcK1)] = hv;
  } else if ((r = pi - L.off_b0) >= 0 && r < kH1) {
    if (bias) bias[r] = val;
    if (w && tc_bias_folded_l1(L.D)) w[kTcOffW0 + tc_off(r, kTcK1 - 1, kTcK1)] = hv;      // b0: column 15 of the W0 image
  } else if ((r = pi - L.off_w2t) >= 0 && r < kH1 * kW2LD) {
    const int k = r / kW2LD, j = r - k * kW2LD;
    if (j < kH2) {
      if (w) w[kTcOffW2 + tc_off(j, k, kH1)] = hv;
      if (bwd) bwd[kTcBwdOffW2 + tc_off(k, j, kH2)] = hv;
    }
  } else if ((r = pi - L.off_b2) >= 0 && r < kH2) {
    if (bias) bias[kH1 + r] = val;
    if (w) w[kTcOffB2T + tc_off(r, 0, kTcK1)] = hv;                                         // b2: column 0 of the bias tile
  } else if ((r = pi - L.off_wh) >= 0 && r < L.NH * kH2) {
    const int a = r / kH2, j = r - a * kH2;
    if (w) w[kTcOffWh + tc_off(a, j, kH2)] = hv;
    if (bwd) bwd[kTcBwdOffWhT + tc_off(j, a, kTcNH)] = hv;
  } else if ((r = pi - L.off_bh) >= 0 && r < L.NH) {
    if (bias) bias[kH1 + kH2 + r] = val;
  }
}
__device__ __forceinline__ void tc_pack_updated(const NetLayout& L, const StepScalars& S, int pi, float2 pt, const TcPackOut& P) {
  if (S.phases & 16) tc_pack_param(L, pi, pt.x, P.fwd_online, P.bwd_online);
  if ((S.phases & (32 | 64)) && P.fwd_target != nullptr) tc_pack_param(L, pi, pt.y, P.fwd_target, nullptr);
}

// ------------------------------------------------------------------------------------------ TD / head deltas
// raw heads [16] -> Q values q[0..A): fully unrolled with predicates so the arrays stay in registers (A <= 15)
__device__ __forceinline__ void heads_to_q(const float (&h)[16], int A, int dueling, float (&q)[16]) {
  if (dueling) {
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < 15; ++a) sum += (a < A) ? h[1 + a] : 0.f;
    const float mean = sum / static_cast<float>(A);
#pragma unroll
    for (int a = 0; a < 15; ++a) q[a] = h[0] + (h[1 + a] - mean);
    q[15] = 0.f;
  } else {
#pragma unroll
    for (int a = 0; a < 16; ++a) q[a] = h[a];
  }
}
__device__ __forceinline__ int argmax_first16(const float (&q)[16], int A) {
  int best = 0;
  float bv = q[0];
#pragma unroll
  for (int a = 1; a < 15; ++a)
    if (a < A && q[a] > bv) { bv = q[a]; best = a; }
  return best;
}

constexpr int kTdThreads = 128;
__global__ void __launch_bounds__(kTdThreads) k_tc_td(AgentCtx C, StepScalars S, TcTrainBufs T) {
  pdl_enter();
  const SpanScope span_(SPAN_TD);
  __shared__ float s_part[kTdThreads / 32];
  const long long i = blockIdx.x * static_cast<long long>(kTdThreads) + threadIdx.x;
  const NetLayout& L = C.L;
  const bool per = S.prioritized != 0;
  float lterm = 0.f;
  if (i < S.B) {
    const int rf = C.rp.row_floats;
    float hn[16], ht[16], hs[16], qn[16], qt[16], qs[16];
    // every input of the row is requested before the first one is used (the transition fields and the IS weight used to be a
    // second, dependent memory round trip behind the head rows)
    const int act = __float_as_int(__ldcg(C.X + i * rf + 2 * L.D));
    const float rew = __ldcg(C.X + i * rf + 2 * L.D + 1), done = __ldcg(C.X + i * rf + 2 * L.D + 2);
    const float w = per ? __ldcg(C.is_w + i) : 1.f;
#pragma unroll
    for (int k = 0; k < 16; k += 4) {
      *reinterpret_cast<float4*>(hn + k) = *reinterpret_cast<const float4*>(T.heads_n + i * 16 + k);
      *reinterpret_cast<float4*>(ht + k) = *reinterpret_cast<const float4*>(T.heads_t + i * 16 + k);
      *reinterpret_cast<float4*>(hs + k) = *reinterpret_cast<const float4*>(T.heads_s + i * 16 + k);
    }
    heads_to_q(hn, L.A, L.dueling, qn);
    heads_to_q(ht, L.A, L.dueling, qt);
    heads_to_q(hs, L.A, L.dueling, qs);
    float qsel = qt[0];
    if (S.double_dqn) {
      const int astar = argmax_first16(qn, L.A);
#pragma unroll
      for (int a = 1; a < 15; ++a) qsel = (a == astar) ? qt[a] : qsel;
    } else {
#pragma unroll
      for (int a = 1; a < 15; ++a) qsel = (a < L.A) ? fmaxf(qsel, qt[a]) : qsel;
    }
    const float y = rew + ((1.f - done) * S.gamma) * qsel;
    float q_sa = qs[0];
#pragma unroll
    for (int a = 1; a < 15; ++a) q_sa = (a == act) ? qs[a] : q_sa;
    const float delta = q_sa - y;
    const float atd = fabsf(y - q_sa);
    const float z = fabsf(delta);
    const float hub = (z < 1.f) ? (0.5f * z) * z : z - 0.5f;
    const float go = per ? (1.f / static_cast<float>(S.Bglobal)) * w : 1.f / static_cast<float>(S.Bglobal);
    const float g = fminf(fmaxf(delta, -1.f), 1.f) * go;
    lterm = per ? w * hub : hub;
    C.y[i] = y; C.q_sa[i] = q_sa; C.abs_td[i] = atd; C.hub[i] = hub; C.gcoef[i] = g;
    float dh[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) dh[a] = 0.f;
    if (L.dueling) {
      const float mean = g / static_cast<float>(L.A);
      dh[0] = g;
#pragma unroll
      for (int a = 0; a < 15; ++a) dh[1 + a] = (a < L.A) ? ((a == act) ? g : 0.f) - mean : 0.f;
    } else {
#pragma unroll
      for (int a = 0; a < 15; ++a) dh[a] = (a < L.A && a == act) ? g : 0.f;
    }
    // (this mode keeps the per-sample scalars only; the fp32 Q / delta rows of the exact path are not materialised)
    uint4 lo, hi;
    lo.x = pack_bf16x2(dh[0], dh[1]); lo.y = pack_bf16x2(dh[2], dh[3]); lo.z = pack_bf16x2(dh[4], dh[5]); lo.w = pack_bf16x2(dh[6], dh[7]);
    hi.x = pack_bf16x2(dh[8], dh[9]); hi.y = pack_bf16x2(dh[10], dh[11]); hi.z = pack_bf16x2(dh[12], dh[13]); hi.w = pack_bf16x2(dh[14], dh[15]);
    *reinterpret_cast<uint4*>(T.DHb + i * 16) = lo;
    *reinterpret_cast<uint4*>(T.DHb + i * 16 + 8) = hi;
  }
  // block loss partial, fixed order
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) lterm += __shfl_down_sync(0xffffffffu, lterm, sh);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = lterm;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kTdThreads / 32; ++w) s += s_part[w];
    C.loss_part[blockIdx.x] = s;
  }
}

__device__ __forceinline__ uint32_t tc_idesc_bf16_mn(int M, int N) { return tc_idesc_bf16(M, N) | (1u << 15) | (1u << 16); }   // A and B MN-major

// ------------------------------------------------------------------------------------------ fused backward
// k_tc_bwd_fused: dgrad chain AND weight gradients of a 128-row tile in one CTA; DZ1 / DZ2 never leave the SM and every
// activation is read from HBM exactly once (X | H1 | H2 | DH: 52 MB per 65,536-row step).
//
// One shared-memory image serves BOTH operand roles.  A [128 b][C] tile stored as 8x8 cores (element (b, c) at
// core * 64 + (b%8)*8 + c%8) is a K-major operand (rows = b, K = c) AND an MN-major operand (MN = c, K = b): only the two
// descriptor strides swap roles.  X, H1, H2 arrive in the forward kernel's own layout tc_off(b, c, C) -- the forward
// kernel's shared-memory images, bulk-stored per tile and bulk-loaded here with three TMA copies -- with cores of
// neighbouring columns 128 B apart and 8-row groups (C/8)*128 B apart: as MN-major operands LBO = (C/8)*128, SBO = 128.
// The tiles this kernel produces (DH, DZ2, DZ1) use uoff(b, c) (8-row groups 128 B apart, column cores 2048 B apart):
// K-major (LBO, SBO) = (2048, 128), MN-major (128, 2048).  So the DZ2 tile the first epilogue writes is the A operand of
// dz1 = DZ2.W2 (K-major) and the B operand of dW2 = H1^T.DZ2 (MN-major) without a second copy.
//
// Bias gradients ride on the tensor core: the staged X tile gets 1.0 in its spare column 15, so column 15 of
// dW0 = DZ1^T.X is db0 and an extra N = 16 product DZ2^T.X yields db2 (dbh: 16 threads sum the 128x16 DH tile).
//
// TMEM columns: [0,128) scratch (dh2, dz1[:, :128], dz1[:, 128:]) | [128,384) dW2^T (two M halves) | [384,416) dW0 (two
// halves) | [416,432) dWh | [432,448) DZ2^T.X.   Per tile: stage -> dh2 -> mask(H2) -> DZ2 -> dz1a -> mask(H1) -> DZ1a
// (into the H2 buffer, free once dWh has completed) -> dW0a, dz1b -> DZ1b (into the DZ2 buffer, free once dW2 / dz1b
// have completed) -> dW0b.  The next tile's H1 / H2 loads are issued as soon as their buffers are free (before dW0b).
__host__ __device__ __forceinline__ int uoff(int b, int c) { return (((c >> 3) << 4) + (b >> 3)) * 64 + (b & 7) * 8 + (c & 7); }
constexpr int kBfOffDH = kTcBwdBytes;                               // byte offsets inside the dynamic shared memory
constexpr int kBfOffX = kBfOffDH + kTcRows * kTcNH * 2;
constexpr int kBfOffH2 = kBfOffX + kTcRows * kTcK1 * 2;             // later: DZ1 half a
constexpr int kBfOffDZ2 = kBfOffH2 + kTcRows * kH2 * 2;             // later: DZ1 half b
constexpr int kBfOffH1 = kBfOffDZ2 + kTcRows * kH2 * 2;
constexpr int kBfOffBars = kBfOffH1 + kTcRows * kH1 * 2;
constexpr int kTcBwdFusedSmemBytes = kBfOffBars + 8 * 8 + 16;       // 208,976 B

// epilogue: scratch columns [col0, col0+64) of this warp's 32 rows -> activation derivative from the saved OUTPUT `act`
// (forward-layout tile with act_K columns, column offset act_c0; ReLU: [h > 0], ELU: h > 0 ? 1 : h + 1 = exp(z)) -> bf16 ->
// uoff tile `dst`
template <int kind>
__device__ __forceinline__ void bf_mask_epilogue_t(uint32_t tS, int q, int row, int col0, const __nv_bfloat16* __restrict__ act, int act_c0,
                                                   int act_K, __nv_bfloat16* __restrict__ dst) {
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    const int col = col0 + 32 * blk;
    uint32_t v[32];
    tc_ld32(tS + (static_cast<uint32_t>(32 * q) << 16) + col, v);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 hq = *reinterpret_cast<const uint4*>(act + tc_off(row, act_c0 + col + 8 * c, act_K));
      const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {     // bf16 > 0  <=>  sign clear and magnitude bits non-zero
        const uint32_t lo = hw[e] & 0xffffu, hi = hw[e] >> 16;
        const bool pos_lo = (lo != 0u && lo < 0x8000u), pos_hi = (hi != 0u && hi < 0x8000u);
        const float u_lo = __uint_as_float(v[8 * c + 2 * e]), u_hi = __uint_as_float(v[8 * c + 2 * e + 1]);
        if (kind == 0) {
          f[2 * e] = pos_lo ? u_lo : 0.f;
          f[2 * e + 1] = pos_hi ? u_hi : 0.f;
        } else {                        // ELU: the bf16 bits widened to fp32 are h itself
          f[2 * e] = pos_lo ? u_lo : u_lo * (__uint_as_float(lo << 16) + 1.f);
          f[2 * e + 1] = pos_hi ? u_hi : u_hi * (__uint_as_float(hi << 16) + 1.f);
        }
      }
      uint4 pk;
      pk.x = pack_bf16x2(f[0], f[1]); pk.y = pack_bf16x2(f[2], f[3]); pk.z = pack_bf16x2(f[4], f[5]); pk.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(dst + uoff(row, col + 8 * c)) = pk;
    }
  }
}

__device__ __forceinline__ void bf_mask_epilogue(uint32_t tS, int q, int row, int col0, const __nv_bfloat16* __restrict__ act, int act_c0,
                                                 int act_K, __nv_bfloat16* __restrict__ dst, int kind) {
  if (kind != 0) bf_mask_epilogue_t<1>(tS, q, row, col0, act, act_c0, act_K, dst);
  else bf_mask_epilogue_t<0>(tS, q, row, col0, act, act_c0, act_K, dst);
}

__global__ void __launch_bounds__(kThreads, 1) k_tc_bwd_fused(AgentCtx C, const unsigned char* __restrict__ packed_bwd, long long n, TcTrainBufs T) {
  pdl_enter();
  const SpanScope span_(SPAN_BWD);
  extern __shared__ __align__(128) unsigned char tsm[];
  const __nv_bfloat16* sW = reinterpret_cast<const __nv_bfloat16*>(tsm);          // Wh^T | W2 (K-major, packed by k_tc_pack_bwd)
  __nv_bfloat16* sDH = reinterpret_cast<__nv_bfloat16*>(tsm + kBfOffDH);
  __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(tsm + kBfOffX);
  __nv_bfloat16* sH2 = reinterpret_cast<__nv_bfloat16*>(tsm + kBfOffH2);
  __nv_bfloat16* sDZ2 = reinterpret_cast<__nv_bfloat16*>(tsm + kBfOffDZ2);
  __nv_bfloat16* sH1 = reinterpret_cast<__nv_bfloat16*>(tsm + kBfOffH1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + kBfOffBars);                 // [0] weights, [1..5] MMA stages A..E
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const NetLayout& L = C.L;
  const long long n_tiles = (n + kTcRows - 1) / kTcRows;
  float* part = T.partials + static_cast<size_t>(blockIdx.x) * L.total;
  if (tid == 0) {
    for (int b = 0; b < 6; ++b) mbar_init(bars + b, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tW2 = tmem + 128, tW0 = tmem + 384, tWh = tmem + 416, tB2 = tmem + 432;
  if (tid == 0) {
    mbar_expect_tx(bars + 0, kTcBwdBytes);
    bulk_g2s(tsm, packed_bwd, kTcBwdBytes, bars + 0);
  }
  const uint32_t idK_128 = tc_idesc_bf16(kTcRows, 128);                         // K-major A and B, N = 128
  const uint32_t idMN_128 = tc_idesc_bf16_mn(128, 128), idMN_16 = tc_idesc_bf16_mn(128, 16);
  constexpr uint32_t kUL = 2048, kUS = 128;                                     // universal tile: K-major (LBO, SBO) = (2048, 128); MN-major = (128, 2048)
  const int q = warp & 3, half = warp >> 2;
  const int row = 32 * q + lane;
  float dbh = 0.f;                     // threads 0..15: bias gradient of head column tid
  uint32_t phase = 0;
  bool first = true;
  uint64_t* barL = bars + 5;           // the tile's TMA loads
  constexpr uint32_t kLoadBytes = (kTcRows * kH1 + kTcRows * kH2 + kTcRows * kTcK1) * 2;
  // descriptor strides of the forward-layout tiles used as MN-major operands: (LBO = 8-row-group stride, SBO = 128)
  constexpr uint32_t kH1L = (kH1 / 8) * 128, kH2L = (kH2 / 8) * 128, kXL = (kTcK1 / 8) * 128;
  if (tid == 0 && static_cast<long long>(blockIdx.x) < n_tiles) {
    const long long t = blockIdx.x;
    mbar_expect_tx(barL, kLoadBytes);
    bulk_g2s(sH1, T.H1b + t * (kTcRows * kH1), kTcRows * kH1 * 2, barL);
    bulk_g2s(sH2, T.H2b + t * (kTcRows * kH2), kTcRows * kH2 * 2, barL);
    bulk_g2s(sX, T.Xb + t * (kTcRows * kTcK1), kTcRows * kTcK1 * 2, barL);
  }
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kTcRows;
    const int rows = static_cast<int>(min(static_cast<long long>(kTcRows), n - row0));
    const long long next = tile + gridDim.x;
    // ---- S0: DH by the register path (uoff layout); X | H1 | H2 arrive by TMA in the forward kernel's layout
    {
      const int b = tid >> 1, c8 = tid & 1;
      uint4 dh = make_uint4(0, 0, 0, 0);
      if (b < rows) dh = *reinterpret_cast<const uint4*>(T.DHb + (row0 + b) * kTcNH + 8 * c8);
      *reinterpret_cast<uint4*>(sDH + uoff(b, 8 * c8)) = dh;
    }
    mbar_wait(barL, phase);
    if (tid < rows) sX[tc_off(tid, 15, kTcK1)] = __float2bfloat16_rn(1.f);   // spare column: bias gradients as a GEMM column
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (first) mbar_wait(bars + 0, 0);
    const uint32_t acc = first ? 0u : 1u;
    // ---- S1: dh2 = DH . Wh (scratch), dWh += H2^T . DH
    if (tid == 0) {
      tc_fence_after();
      tc_mma_bf16(tS, tc_smem_desc(sDH, kUL, kUS), tc_smem_desc(sW + kTcBwdOffWhT, 128, (kTcNH / 8) * 128), idK_128, 0u);
      tc_commit(bars + 1);
      const uint64_t aH2 = tc_smem_desc(sH2, kH2L, 128), bDH = tc_smem_desc(sDH, kUS, kUL);
#pragma unroll
      for (int ks = 0; ks < kTcRows / 16; ++ks) tc_mma_bf16(tWh, aH2 + (2u * kH2L / 16u) * ks, bDH + 16u * ks, idMN_16, (ks > 0) ? 1u : acc);
    }
    if (tid < kTcNH) {                 // dbh: column sums of the staged DH tile
      float s = 0.f;
      for (int b = 0; b < rows; ++b) s += __bfloat162float(sDH[uoff(b, tid)]);
      dbh += s;
    }
    mbar_wait(bars + 1, phase);
    tc_fence_after();
    // ---- S2: DZ2 = dh2 (.) [H2 > 0]
    bf_mask_epilogue(tS, q, row, 64 * half, sH2, 0, kH2, sDZ2, L.act);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- S3: dz1[:, :128] = DZ2 . W2[:128] (scratch); dW2 += H1^T . DZ2 (two M halves); db2 column via DZ2^T . X
    if (tid == 0) {
      tc_fence_after();
      const uint64_t aDZ2k = tc_smem_desc(sDZ2, kUL, kUS), bW2 = tc_smem_desc(sW + kTcBwdOffW2, 128, (kH2 / 8) * 128);
#pragma unroll
      for (int k = 0; k < kH2 / 16; ++k) tc_mma_bf16(tS, aDZ2k + 256u * k, bW2 + 16u * k, idK_128, k > 0 ? 1u : 0u);
      tc_commit(bars + 2);
      const uint64_t aH1 = tc_smem_desc(sH1, kH1L, 128), aH1b = tc_smem_desc(sH1 + tc_off(0, 128, kH1), kH1L, 128);
      const uint64_t bDZ2 = tc_smem_desc(sDZ2, kUS, kUL), bX = tc_smem_desc(sX, kXL, 128);
#pragma unroll
      for (int ks = 0; ks < kTcRows / 16; ++ks) {
        const uint32_t a2 = (ks > 0) ? 1u : acc;
        tc_mma_bf16(tW2, aH1 + (2u * kH1L / 16u) * ks, bDZ2 + 16u * ks, idMN_128, a2);
        tc_mma_bf16(tW2 + 128, aH1b + (2u * kH1L / 16u) * ks, bDZ2 + 16u * ks, idMN_128, a2);
        tc_mma_bf16(tB2, bDZ2 + 16u * ks, bX + (2u * kXL / 16u) * ks, idMN_16, a2);
      }
    }
    mbar_wait(bars + 2, phase);        // dz1a done; the commit also covers dWh -> the H2 buffer is free
    tc_fence_after();
    // ---- S4: DZ1[:, :128] = dz1a (.) [H1[:, :128] > 0]  -> H2 buffer (uoff layout)
    bf_mask_epilogue(tS, q, row, 64 * half, sH1, 0, kH1, sH2, L.act);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- S5: dW0[:128] += DZ1a^T . X ; dz1[:, 128:] = DZ2 . W2[128:]
    if (tid == 0) {
      tc_fence_after();
      const uint64_t aZ = tc_smem_desc(sH2, kUS, kUL), bX = tc_smem_desc(sX, kXL, 128);
#pragma unroll
      for (int ks = 0; ks < kTcRows / 16; ++ks) tc_mma_bf16(tW0, aZ + 16u * ks, bX + (2u * kXL / 16u) * ks, idMN_16, (ks > 0) ? 1u : acc);
      const uint64_t aDZ2k = tc_smem_desc(sDZ2, kUL, kUS), bW2 = tc_smem_desc(sW + kTcBwdOffW2 + tc_off(128, 0, kH2), 128, (kH2 / 8) * 128);
#pragma unroll
      for (int k = 0; k < kH2 / 16; ++k) tc_mma_bf16(tS, aDZ2k + 256u * k, bW2 + 16u * k, idK_128, k > 0 ? 1u : 0u);
      tc_commit(bars + 3);
    }
    mbar_wait(bars + 3, phase);        // dz1b done; covers dW2 / db2 / dz1a / dW0a -> the DZ2 and H2 buffers are free
    tc_fence_after();
    // ---- S6: DZ1[:, 128:] -> DZ2 buffer
    bf_mask_epilogue(tS, q, row, 64 * half, sH1, 128, kH1, sDZ2, L.act);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- S7: dW0[128:] += DZ1b^T . X ; the next tile's H1 / H2 start loading (nothing reads those buffers any more)
    if (tid == 0) {
      tc_fence_after();
      const uint64_t aZ = tc_smem_desc(sDZ2, kUS, kUL), bX = tc_smem_desc(sX, kXL, 128);
#pragma unroll
      for (int ks = 0; ks < kTcRows / 16; ++ks) tc_mma_bf16(tW0 + 16, aZ + 16u * ks, bX + (2u * kXL / 16u) * ks, idMN_16, (ks > 0) ? 1u : acc);
      tc_commit(bars + 4);
      if (next < n_tiles) {
        mbar_expect_tx(barL, kLoadBytes);
        bulk_g2s(sH1, T.H1b + next * (kTcRows * kH1), kTcRows * kH1 * 2, barL);
        bulk_g2s(sH2, T.H2b + next * (kTcRows * kH2), kTcRows * kH2 * 2, barL);
      }
    }
    mbar_wait(bars + 4, phase);        // every operand buffer may be rewritten
    tc_fence_after();
    if (tid == 0 && next < n_tiles) bulk_g2s(sX, T.Xb + next * (kTcRows * kTcK1), kTcRows * kTcK1 * 2, barL);
    __syncthreads();
    phase ^= 1u;
    first = false;
  }
  // ---- accumulators -> this CTA's partial gradient blob (device parameter layout)
  for (int mh = 0; mh < 2; ++mh) {            // dW2^T rows k = 128*mh + row, this warp's 64 columns j
    const int k = 128 * mh + row;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int j0 = 64 * half + 32 * b;
      uint32_t v[32];
      tc_ld32(tW2 + 128 * mh + (static_cast<uint32_t>(32 * q) << 16) + j0, v);
#pragma unroll
      for (int e = 0; e < 32; e += 4)
        *reinterpret_cast<float4*>(part + L.off_w2t + k * kW2LD + j0 + e) =
            make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
      // the 4 padding columns of the row as well: every 32-byte sector of the partial blob is then written completely
      // (a partially written sector has to be merged with DRAM contents when the reduction reads it)
      if (j0 + 32 == kH2) *reinterpret_cast<float4*>(part + L.off_w2t + k * kW2LD + kH2) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  {   // dW0^T[d][i] and db0[i] (column 15): half h handles i = 128*h + row
    uint32_t v[16];
    tc_ld16(tW0 + 16 * half + (static_cast<uint32_t>(32 * q) << 16), v);
    const int i = 128 * half + row;
#pragma unroll
    for (int d = 0; d < 15; ++d)
      if (d < L.D) part[L.off_w0t + d * kH1 + i] = __uint_as_float(v[d]);
    part[L.off_b0 + i] = __uint_as_float(v[15]);
  }
  if (half == 0) {   // dWh[a][j], j = row
    uint32_t v[16];
    tc_ld16(tWh + (static_cast<uint32_t>(32 * q) << 16), v);
#pragma unroll
    for (int a = 0; a < 16; ++a)
      if (a < L.NH) part[L.off_wh + a * kH2 + row] = __uint_as_float(v[a]);
  } else {           // db2[j] = column 15 of DZ2^T . X
    uint32_t v[16];
    tc_ld16(tB2 + (static_cast<uint32_t>(32 * q) << 16), v);
    part[L.off_b2 + row] = __uint_as_float(v[15]);
  }
  if (tid < ((L.NH + 3) & ~3)) part[L.off_bh + tid] = (tid < L.NH) ? dbh : 0.f;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// fixed-order reduction of the per-CTA partials -> gradient blob -> Adam (+ Polyak) -> refreshed bf16 images; also the loss.
// A block owns 128 consecutive parameters: warp w sums partials w, w+8, ... with float4 loads (512 contiguous bytes per
// warp load), the eight warp sums are combined in warp order through shared memory.
__global__ void __launch_bounds__(256) k_tc_reduce_adam(AgentCtx C, StepScalars S, TcTrainBufs T, int n_loss_parts, TcPackOut P, int dbg_skip) {
  pdl_enter();
  const SpanScope span_(SPAN_REDUCE_ADAM);
  __shared__ float4 s_sum[8][32];
  const NetLayout& L = C.L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p4 = blockIdx.x * 32 + lane;                 // float4 index
  const int n4 = L.total >> 2;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p4 < n4 && !(dbg_skip & 1)) {
    // batches of 8 predicated loads, all in flight before the first add (a plain `#pragma unroll` leaves a remainder loop of
    // dependent load -> add iterations at full L2 latency each: 14 of this kernel's 21 us were spent there)
    const float4* src = reinterpret_cast<const float4*>(T.partials) + p4;
    for (int c0 = warp; c0 < T.n_part; c0 += 64) {
      float4 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = c0 + 8 * q;
        v[q] = (c < T.n_part) ? __ldcg(src + static_cast<size_t>(c) * n4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
    }
  }
  s_sum[warp][lane] = acc;
  __syncthreads();
  if (threadIdx.x < 128 && !(dbg_skip & 2)) {
    const int l4 = threadIdx.x >> 2, e = threadIdx.x & 3;
    const int pi = (blockIdx.x * 32 + l4) * 4 + e;
    if (pi < L.total) {
      float g = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) g += reinterpret_cast<const float*>(&s_sum[w][l4])[e];
      C.grads[pi] = g;
      const float2 pt = adam_polyak_element(C, S, pi, g);
      tc_pack_updated(L, S, pi, pt, P);
    }
  }
  // loss: the per-block partials of k_tc_td (up to 1024) summed by the LAST block in a fixed order: thread t adds partials
  // t, t+256, ... (loads in flight together), then the 256 thread sums are added in thread order.  (One thread walking the
  // list serially took 20 of this kernel's 23 us: 512 dependent L2 round trips.)
  if (blockIdx.x == gridDim.x - 1 && !(dbg_skip & 4)) {
    __shared__ float s_loss[256];
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int c = threadIdx.x + 256 * q; v[q] = (c < n_loss_parts) ? __ldcg(C.loss_part + c) : 0.f; }
    __syncthreads();      // s_sum readers are done
    s_loss[threadIdx.x] = ((v[0] + v[1]) + v[2]) + v[3];
    __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int c = 0; c < 256; ++c) s += s_loss[c];
    const float loss = s / static_cast<float>(S.Bglobal);
    C.loss[0] = loss;
    host_loss_store(C.host_loss, loss, S.epoch);
  }
  }
}

}  // namespace rmc
