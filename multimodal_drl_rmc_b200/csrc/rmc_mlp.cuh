// rmc_mlp.cuh -- the fused learner-step kernel (one cooperative launch per step) and the
// batched act / Q-value kernel for the macro-state MLP  D -> 256 -> ReLU|ELU -> 128 -> ReLU|ELU -> heads.
//
// Phase A (row parallel, CTAs own tiles of kTM batch rows):
//     TMA bulk copy of the whole target / online parameter blob into shared memory,
//     sample (PER prefix search or uniform) + coalesced row gather,
//     Q_target(s'), then [Q_online(s'); Q_online(s)] in one 2*kTM-row pass,
//     double-DQN target, |td|, Huber, priorities + last-writer stamps, and the dgrad chain
//     (dQ -> dheads -> dh2 -> dz2 -> dz1); activations/deltas of the s rows go to L2-resident scratch.
// ---- one agent-wide barrier ----
// Phase B (parameter parallel): output-stationary weight-gradient tiles with a fixed summation
//     order (deterministic), fused Adam (+ Polyak) on the owning CTA; the last CTA of a PER agent
//     meanwhile applies the priority write-back to the sum tree.
//
// Arithmetic follows SURVEY.md Appendix A / dqn/agent.py:245-272, dqn/network.py:83,90-96 and
// torch.optim.Adam (single-tensor path); fp32 FFMA accumulation only (no tensor cores here: the
// 1e-5 parity bar needs fp32 accumulation and at B<=1024 the step is latency bound).
#pragma once
#include "rmc_device.cuh"
#include "rmc_tree.cuh"

namespace rmc {

// ------------------------------------------------------------------ shared memory carve-up (floats)
struct SmemPlan {
  int w;        // parameter blob (L.total floats)
  int xt;       // [kMaxD][kR]      x transposed
  int h1t;      // [kH1][kR]        h1 transposed
  int h2;       // [kR][kH2]
  int part;     // [kWarps][kR][kH2] K-split partials of layer 2
  int q;        // [kR][kQLD]
  int dz2;      // [kTM][kH2]
  int dh;       // [kTM][kQLD]
  int meta;     // [kTM][4]  action(bits), reward, done, is_w
  int red;      // [kWarps] loss partials
  int rows;     // [kTM][kMaxRowFloats] gathered rows of the CTA's tile (single-tile fast path)
  int qt;       // [kTM][kQLD] Q_target(s') of the tile
  int isw;      // [kTM] importance weights of the tile
  int top;      // [kTopNodes] float64: top levels of the sum tree (8-byte aligned)
  int bar;      // mbarrier (8 bytes, 8-byte aligned)
  int total_floats;
};
__host__ __device__ inline SmemPlan make_smem_plan(int param_floats) {
  SmemPlan s;
  int o = 0;
  s.w = o; o += param_floats;             // multiple of 4
  s.xt = o; o += kMaxD * kR;
  s.h1t = o; o += kH1 * kR;
  s.h2 = o; o += kR * kH2;
  s.part = o; o += kWarps * kR * kH2;
  s.q = o; o += kR * kQLD;
  s.dz2 = o; o += kTM * kH2;
  s.dh = o; o += kTM * kQLD;
  s.meta = o; o += kTM * 4;
  s.red = o; o += kWarps;
  s.rows = o; o += kTM * kMaxRowFloats;
  s.qt = o; o += kTM * kQLD;
  s.isw = o; o += kTM;
  o = (o + 3) & ~3;
  s.top = o; o += 2 * kTopNodes + 2;
  o = (o + 3) & ~3;
  s.bar = o; o += 4;
  s.total_floats = o;
  return s;
}
// phase-B staging (reuses the same dynamic shared memory)
constexpr int kGChunk = 256;                         // batch rows staged per chunk
// phase-B unit shapes: 16x16 (W2), 16x32 (W0: D<=16 rows x 32 hidden units), 32x16 (heads: 32 h2 columns x NH)
constexpr int kGemmSmemFloats = 2 * kGChunk * (32 + 32) + kWarps * 32 * 32 + kWarps * 32;

// ------------------------------------------------------------------ forward of R rows (R = 4 or 8)
// sXT[d][kR], rows [0,R) are computed.  Results: sH1T[k][r], sH2[r][j], sQ[r][a] (Q values, or raw
// head outputs when raw_heads: [0]=val, [1..A]=adv for dueling).
template <int R>
__device__ __forceinline__ void mlp_forward(const float* __restrict__ sW, const NetLayout& L, const float* __restrict__ sXT,
                                            float* __restrict__ sH1T, float* __restrict__ sH2, float* __restrict__ sPart,
                                            float* __restrict__ sQ, float* __restrict__ sRaw) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- layer 1: thread i owns hidden unit i for all R rows
  {
    float acc[R];
    const float b = sW[L.off_b0 + tid];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = b;
    const float* w0t = sW + L.off_w0t + tid;
    for (int d = 0; d < L.D; ++d) {
      const float w = w0t[d * kH1];
      const float4* xp = reinterpret_cast<const float4*>(sXT + d * kR);
      const float4 x0 = xp[0];
      acc[0] = fmaf(x0.x, w, acc[0]); acc[1] = fmaf(x0.y, w, acc[1]);
      acc[2] = fmaf(x0.z, w, acc[2]); acc[3] = fmaf(x0.w, w, acc[3]);
      if (R == 8) {
        const float4 x1 = xp[1];
        acc[4 % R] = fmaf(x1.x, w, acc[4 % R]); acc[5 % R] = fmaf(x1.y, w, acc[5 % R]);
        acc[6 % R] = fmaf(x1.z, w, acc[6 % R]); acc[7 % R] = fmaf(x1.w, w, acc[7 % R]);
      }
    }
    float4* hp = reinterpret_cast<float4*>(sH1T + tid * kR);
    hp[0] = make_float4(act_fwd(acc[0], L.act), act_fwd(acc[1], L.act), act_fwd(acc[2], L.act), act_fwd(acc[3], L.act));
    if (R == 8) hp[1] = make_float4(act_fwd(acc[4 % R], L.act), act_fwd(acc[5 % R], L.act), act_fwd(acc[6 % R], L.act), act_fwd(acc[7 % R], L.act));
  }
  __syncthreads();
  // ---- layer 2, K split over the 8 warps (32 k each); lane owns 4 columns for all R rows
  {
    float acc[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
    const float* w2 = sW + L.off_w2t + (warp * 32) * kW2LD + lane * 4;
    const float* h1 = sH1T + (warp * 32) * kR;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(w2 + k * kW2LD);
      const float4 a0 = *reinterpret_cast<const float4*>(h1 + k * kR);
      float a[R];
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      if (R == 8) {
        const float4 a1 = *reinterpret_cast<const float4*>(h1 + k * kR + 4);
        a[4 % R] = a1.x; a[5 % R] = a1.y; a[6 % R] = a1.z; a[7 % R] = a1.w;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][0] = fmaf(a[r], w.x, acc[r][0]);
        acc[r][1] = fmaf(a[r], w.y, acc[r][1]);
        acc[r][2] = fmaf(a[r], w.z, acc[r][2]);
        acc[r][3] = fmaf(a[r], w.w, acc[r][3]);
      }
    }
    float* pp = sPart + (warp * kR) * kH2 + lane * 4;
#pragma unroll
    for (int r = 0; r < R; ++r) *reinterpret_cast<float4*>(pp + r * kH2) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
  // ---- reduce the 8 partials in fixed order, bias, ReLU -> sH2[r][j]
  {
    const int r = tid >> 5;            // 0..7
    if (r < R) {
      const int c = lane * 4;
      float4 s = *reinterpret_cast<const float4*>(sW + L.off_b2 + c);
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const float4 p = *reinterpret_cast<const float4*>(sPart + (w * kR + r) * kH2 + c);
        s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
      }
      *reinterpret_cast<float4*>(sH2 + r * kH2 + c) = make_float4(act_fwd(s.x, L.act), act_fwd(s.y, L.act), act_fwd(s.z, L.act), act_fwd(s.w, L.act));
    }
  }
  __syncthreads();
  // ---- heads: warp r <-> row r, lanes over j, butterfly reduce
  if (warp < R) {
    const float* h = sH2 + warp * kH2;
    const float h0 = h[lane], h1v = h[lane + 32], h2v = h[lane + 64], h3v = h[lane + 96];
    float out = 0.f;   // lane a keeps head a
    for (int a = 0; a < L.NH; ++a) {
      const float* wh = sW + L.off_wh + a * kH2;
      float s = h0 * wh[lane];
      s = fmaf(h1v, wh[lane + 32], s);
      s = fmaf(h2v, wh[lane + 64], s);
      s = fmaf(h3v, wh[lane + 96], s);
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
      if (lane == a) out = s + sW[L.off_bh + a];
    }
    if (sRaw != nullptr && lane < kQLD) sRaw[warp * kQLD + lane] = (lane < L.NH) ? out : 0.f;
    if (L.dueling) {   // Q = val + (adv - mean(adv))   (dqn/network.py:83)
      const float val = __shfl_sync(0xffffffffu, out, 0);
      float sum = 0.f;
      for (int a = 1; a <= L.A; ++a) sum += __shfl_sync(0xffffffffu, out, a);
      const float mean = sum / static_cast<float>(L.A);
      const float adv = __shfl_sync(0xffffffffu, out, (lane + 1) & 31);   // lane a gets adv[a]
      if (lane < kQLD) sQ[warp * kQLD + lane] = (lane < L.A) ? (val + (adv - mean)) : 0.f;
    } else {
      if (lane < kQLD) sQ[warp * kQLD + lane] = (lane < L.A) ? out : 0.f;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int argmax_first(const float* q, int A) {
  int best = 0;
  float bv = q[0];
  for (int a = 1; a < A; ++a) {
    const float v = q[a];
    if (v > bv) { bv = v; best = a; }   // strict '>' keeps the first maximum (torch.argmax)
  }
  return best;
}

// stage one parameter blob into shared memory with a single TMA bulk copy
__device__ __forceinline__ void stage_params(float* sW, const float* gW, int floats, uint64_t* bar, uint32_t& parity) {
  // all generic-proxy accesses to sW by this CTA are complete (caller synchronised); order them
  // before the async-proxy write
  if (threadIdx.x == 0) {
    fence_proxy_async();
    const uint32_t bytes = static_cast<uint32_t>(floats) * 4u;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(sW, gW, bytes, bar);
  }
}
__device__ __forceinline__ void wait_params(uint64_t* bar, uint32_t& parity) {
  mbar_wait(bar, parity);
  parity ^= 1u;
}

// ------------------------------------------------------------------ Adam / Polyak on one element
struct ParamVals { float p, m, v, t; };
__device__ __forceinline__ ParamVals param_load(const AgentCtx& C, const StepScalars& S, int pi) {
  ParamVals x;
  x.p = __ldcg(C.online + pi);
  x.m = (S.phases & 16) ? __ldcg(C.adam_m + pi) : 0.f;
  x.v = (S.phases & 16) ? __ldcg(C.adam_v + pi) : 0.f;
  x.t = (S.phases & 32) ? __ldcg(C.target + pi) : 0.f;
  return x;
}
// returns (online weight after the step, target weight after the step -- meaningful only with POLYAK / HARDSYNC)
__device__ __forceinline__ float2 param_apply(const AgentCtx& C, const StepScalars& S, int pi, float g, ParamVals x) {
  float p = x.p;
  float tnew = x.t;
  if (S.phases & 16 /*ADAM*/) {
    float m = x.m, v = x.v;
    // Rounding sequence of torch 2.11's CPU kernels, found by bit-matching torch.optim.Adam
    // (tests/test_oracle_golden.py::test_numpy_adam_bit_matches_torch): lerp and addcmul are FMAs.
    m = fmaf(S.adam_w1, g - m, m);                     // exp_avg.lerp_(grad, 1-beta1)
    v = fmaf(S.adam_w2 * g, g, v * S.adam_b2);         // mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    const float denom = __fsqrt_rn(v) / S.adam_bc2_sqrt + S.adam_eps;
    p = p + (S.adam_neg_step * m) / denom;             // addcdiv_(exp_avg, denom, value=-step_size)
    C.online[pi] = p;
    C.adam_m[pi] = m;
    C.adam_v[pi] = v;
  }
  if (S.phases & 32 /*POLYAK: dqn/agent.py:105-110, post-Adam weights*/) {
    tnew = S.polyak_k * p + S.polyak_1mk * x.t;
    C.target[pi] = tnew;
  } else if (S.phases & 64 /*HARDSYNC: dqn/agent.py:102-103*/) {
    tnew = p;
    C.target[pi] = p;
  }
  return make_float2(p, tnew);
}
// The same update for N independent elements with the arithmetic of all of them in ONE straight-line block (the
// sqrt / division chains of the elements interleave instead of running back to back) and the stores separate.
template <int N>
__device__ __forceinline__ void param_math_n(const StepScalars& S, const float (&g)[N], const ParamVals (&x)[N], ParamVals (&o)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) o[k] = x[k];
  if (S.phases & 16 /*ADAM*/) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float m = fmaf(S.adam_w1, g[k] - x[k].m, x[k].m);
      const float v = fmaf(S.adam_w2 * g[k], g[k], x[k].v * S.adam_b2);
      const float denom = __fsqrt_rn(v) / S.adam_bc2_sqrt + S.adam_eps;
      o[k].p = x[k].p + (S.adam_neg_step * m) / denom;
      o[k].m = m;
      o[k].v = v;
    }
  }
  if (S.phases & 32 /*POLYAK*/) {
#pragma unroll
    for (int k = 0; k < N; ++k) o[k].t = S.polyak_k * o[k].p + S.polyak_1mk * x[k].t;
  } else if (S.phases & 64 /*HARDSYNC*/) {
#pragma unroll
    for (int k = 0; k < N; ++k) o[k].t = o[k].p;
  }
}
__device__ __forceinline__ void param_store(const AgentCtx& C, const StepScalars& S, int pi, const ParamVals& o) {
  if (S.phases & 16) { C.online[pi] = o.p; C.adam_m[pi] = o.m; C.adam_v[pi] = o.v; }
  if (S.phases & (32 | 64)) C.target[pi] = o.t;
}
__device__ __forceinline__ float2 adam_polyak_element(const AgentCtx& C, const StepScalars& S, int pi, float g) {
  return param_apply(C, S, pi, g, param_load(C, S, pi));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ phase B: one 32x32 gradient tile
// out[m][n] = sum_b A[b][m0+m] * Bm[b][n0+n]   (+ column sums of Bm as the bias gradient)
struct GemmUnit {
  const float* A; int lda; int m0; int m_valid;
  const float* Bm; int ldb; int n0; int n_valid;
  int out_base; int out_sm; int out_sn;    // param index = out_base + m*out_sm + n*out_sn
  int bias_base;                           // param index of bias[n] or -1
};

template <int N>
__device__ __forceinline__ void lds_vec(float* dst, const float* src) {
  static_assert(N == 2 || N == 4 || N == 8, "lane tile width");
  if constexpr (N == 2) {
    const float2 v = *reinterpret_cast<const float2*>(src);
    dst[0] = v.x; dst[1] = v.y;
  } else {
#pragma unroll
    for (int k = 0; k < N; k += 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + k);
      dst[k] = v.x; dst[k + 1] = v.y; dst[k + 2] = v.z; dst[k + 3] = v.w;
    }
  }
}

template <int TMO, int TNO>
__device__ void wgrad_unit(const AgentCtx& C, const StepScalars& S, const GemmUnit& U, float* smem) {
  constexpr int LM = TMO / 8, LN = TNO / 4;       // lane tile (8 x 4 lanes cover the unit)
  constexpr int NOUT = TMO * TNO / kThreads;      // outputs per thread (1 or 2)
  float* stage = smem;                            // 2 x { A [kGChunk][TMO], B [kGChunk][TNO] }
  float* Ps = smem + 2 * kGChunk * (TMO + TNO);   // [kWarps][TMO*TNO]
  float* Pb = Ps + kWarps * TMO * TNO;            // [kWarps][TNO]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mg = lane >> 2, ng = lane & 3;
  // this thread's output elements and their (prefetched) parameter state: independent of the gradient
  int pi[NOUT], ps_off[NOUT];
  ParamVals pv[NOUT] = {};
#pragma unroll
  for (int q = 0; q < NOUT; ++q) {
    const int o = tid + q * kThreads, om = o / TNO, on = o % TNO;
    pi[q] = (om < U.m_valid && on < U.n_valid) ? U.out_base + om * U.out_sm + on * U.out_sn : -1;
    ps_off[q] = om * TNO + ((on + om / LM) & (TNO - 1));      // partial-sum columns are rotated by the lane row (bank spread)
    if (pi[q] >= 0) pv[q] = param_load(C, S, pi[q]);
  }
  // bias column sums: lanes of the LAST warp (for the W0 units its second output row is a masked padding row, so the
  // bias update replaces work instead of adding a third dependent Adam chain to warp 0)
  const int bt = tid - (kThreads - 32);
  const int pb = (U.bias_base >= 0 && bt >= 0 && bt < U.n_valid && bt < TNO) ? U.bias_base + bt : -1;
  ParamVals pvb{};
  if (pb >= 0) pvb = param_load(C, S, pb);
  float acc[LM][LN];
  float bsum[LN];
#pragma unroll
  for (int i = 0; i < LM; ++i)
#pragma unroll
    for (int j = 0; j < LN; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int j = 0; j < LN; ++j) bsum[j] = 0.f;

  // 16-byte chunks per staged row.  When the source rows are at least as wide as the tile (X, DH and the activation
  // arrays all are: their padding columns hold finite values that only feed outputs masked by pi < 0) whole tiles are
  // copied; otherwise columns past the valid ones are zeroed once and never written.
  // Two staging buffers: the cp.async copies of chunk c+1 are in flight while chunk c is multiplied.
  const bool full = (U.m0 + TMO <= U.lda) && (U.n0 + TNO <= U.ldb);
  const int ca = full ? TMO / 4 : min(TMO / 4, (U.m_valid + 3) >> 2), cb = full ? TNO / 4 : min(TNO / 4, (U.n_valid + 3) >> 2);
  constexpr int kBuf = kGChunk * (TMO + TNO);
  __syncthreads();
  if (ca + cb < (TMO + TNO) / 4) {
    for (int t = tid; t < 2 * kBuf; t += kThreads) stage[t] = 0.f;
    __syncthreads();
  }
  // division-free copy schedule: a thread keeps one 16-byte column slot of the staged row and walks down the rows
  constexpr int CH = (TMO + TNO) / 4, SL = (CH <= 8) ? 8 : 16, RPP = kThreads / SL;
  const int slot = tid & (SL - 1), row0 = tid / SL;
  const bool a_slot = slot < TMO / 4;
  const int cc = a_slot ? slot : slot - TMO / 4;
  const bool slot_on = a_slot ? (cc < ca) : (slot < CH && cc < cb);
  const float* src0 = a_slot ? U.A + U.m0 + 4 * cc : U.Bm + U.n0 + 4 * cc;
  const int src_ld = a_slot ? U.lda : U.ldb;
  const int dst_off = a_slot ? 4 * cc : kGChunk * TMO + 4 * cc, dst_ld = a_slot ? TMO : TNO;
  auto issue = [&](long long b0, float* buf) {
    const int rows = static_cast<int>(min(static_cast<long long>(kGChunk), S.B - b0));
    if (slot_on) {
      const float* src = src0 + (b0 + row0) * src_ld;
      float* dst = buf + dst_off + row0 * dst_ld;
#pragma unroll 4
      for (int r = row0; r < rows; r += RPP, src += RPP * src_ld, dst += RPP * dst_ld) cp_async16(dst, src);   // LDGSTS, all in flight
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  RMC_STAMP(C, 16);
  issue(0, stage);
  int cur = 0;
  for (long long b0 = 0; b0 < S.B; b0 += kGChunk, cur ^= 1) {
    const int rows = static_cast<int>(min(static_cast<long long>(kGChunk), S.B - b0));
    const bool more = b0 + kGChunk < S.B;
    if (more) issue(b0 + kGChunk, stage + (cur ^ 1) * kBuf);
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (b0 == 0) RMC_STAMP(C, 17);
    const float* As = stage + cur * kBuf;
    const float* Bs = As + kGChunk * TMO;
    // warp w takes rows w, w+8, ... (fixed order -> deterministic sums)
#pragma unroll 4
    for (int r = warp; r < rows; r += kWarps) {
      float a[LM], b[LN];
      lds_vec<LM>(a, As + r * TMO + mg * LM);       // 8- / 16-byte shared-memory loads (offsets are multiples of LM, LN)
      lds_vec<LN>(b, Bs + r * TNO + ng * LN);
#pragma unroll
      for (int i = 0; i < LM; ++i)
#pragma unroll
        for (int j = 0; j < LN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
      for (int j = 0; j < LN; ++j) bsum[j] += b[j];
    }
    __syncthreads();     // buffer `cur` may be refilled by the next iteration's prefetch
  }
  // cross-warp reduction in fixed order
#pragma unroll
  for (int i = 0; i < LM; ++i)
#pragma unroll
    for (int j = 0; j < LN; ++j) Ps[warp * (TMO * TNO) + (mg * LM + i) * TNO + ((ng * LN + j + mg) & (TNO - 1))] = acc[i][j];
  if (mg == 0) {
#pragma unroll
    for (int j = 0; j < LN; ++j) Pb[warp * TNO + ng * LN + j] = bsum[j];
  }
  __syncthreads();
  RMC_STAMP(C, 18);
  // gradients of this thread's outputs, then Adam / Polyak for all of them together (interleaved sqrt / division chains)
  float gq[NOUT];
  ParamVals po[NOUT];
#pragma unroll
  for (int q = 0; q < NOUT; ++q) {
    float g = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) g += Ps[w * (TMO * TNO) + ps_off[q]];
    gq[q] = g;
  }
  param_math_n<NOUT>(S, gq, pv, po);
#pragma unroll
  for (int q = 0; q < NOUT; ++q) {
    if (pi[q] >= 0) {
      C.grads[pi[q]] = gq[q];
      param_store(C, S, pi[q], po[q]);
    }
  }
  if (pb >= 0) {
    float g = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) g += Pb[w * TNO + bt];
    C.grads[pb] = g;
    param_apply(C, S, pb, g, pvb);
  }
  RMC_STAMP(C, 19);
  __syncthreads();
}

// Unit decomposition.  fine (many workers): [W0: mt0 x 8 of 16x32] [W2: 16 x 8 of 16x16] [heads: 4 of 32x16] = 140
// for D <= 16; coarse (few workers per agent, e.g. 8-agent ensembles): [W0: mt0 x 8 of 16x32] [W2: 8 x 4 of 32x32]
// [heads: 4 of 32x16] = 44.
__device__ __forceinline__ int wgrad_unit_count(const NetLayout& L, bool coarse) {
  const int w2 = coarse ? (kH1 / 32) * (kH2 / 32) : (kH1 / 16) * (kH2 / 16);
  return ((L.D + 15) / 16) * (kH1 / 32) + w2 + kH2 / 32;
}

__device__ void wgrad_run_unit(const AgentCtx& C, const StepScalars& S, int u, bool coarse, float* smem) {
  const NetLayout& L = C.L;
  GemmUnit U;
  const int mt0 = (L.D + 15) / 16;
  const int n_w0 = mt0 * (kH1 / 32), n_w2 = coarse ? (kH1 / 32) * (kH2 / 32) : (kH1 / 16) * (kH2 / 16);
  if (u < n_w0) {                       // dW0^T[d][i] = sum_b X[b][d] * DZ1[b][i] ; db0 = colsum(DZ1)
    const int mt = u / (kH1 / 32), nt = u % (kH1 / 32);
    U.A = C.X; U.lda = C.rp.row_floats; U.m0 = mt * 16; U.m_valid = min(16, L.D - mt * 16);
    U.Bm = C.DZ1; U.ldb = kH1; U.n0 = nt * 32; U.n_valid = 32;
    U.out_base = L.off_w0t + U.m0 * kH1 + U.n0; U.out_sm = kH1; U.out_sn = 1;
    U.bias_base = (mt == 0) ? L.off_b0 + U.n0 : -1;
    wgrad_unit<16, 32>(C, S, U, smem);
  } else if (u < n_w0 + n_w2) {         // dW2^T[k][j] = sum_b H1[b][k] * DZ2[b][j] ; db2 = colsum(DZ2)
    const int T = coarse ? 32 : 16;
    const int v = u - n_w0, kt = v / (kH2 / T), jt = v % (kH2 / T);
    U.A = C.H1; U.lda = kH1; U.m0 = kt * T; U.m_valid = T;
    U.Bm = C.DZ2; U.ldb = kH2; U.n0 = jt * T; U.n_valid = T;
    U.out_base = L.off_w2t + U.m0 * kW2LD + U.n0; U.out_sm = kW2LD; U.out_sn = 1;
    U.bias_base = (kt == 0) ? L.off_b2 + U.n0 : -1;
    if (coarse) wgrad_unit<32, 32>(C, S, U, smem);
    else wgrad_unit<16, 16>(C, S, U, smem);
  } else {                              // dWh[a][j] = sum_b H2[b][j] * DH[b][a] ; dbh = colsum(DH)
    const int jt = u - n_w0 - n_w2;
    U.A = C.H2; U.lda = kH2; U.m0 = jt * 32; U.m_valid = 32;
    U.Bm = C.DH; U.ldb = kQLD; U.n0 = 0; U.n_valid = L.NH;
    U.out_base = L.off_wh + U.m0; U.out_sm = 1; U.out_sn = kH2;
    U.bias_base = (jt == 0) ? L.off_bh : -1;
    wgrad_unit<32, 16>(C, S, U, smem);
  }
}


#include "rmc_rows_ws.cuh"

// ------------------------------------------------------------------ streamed phase B units (see StreamPlan below)
__device__ __forceinline__ uint4 ld_relaxed_quad(const unsigned* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_pair(unsigned* p, unsigned a, unsigned b);
__device__ __forceinline__ uint2 ld_relaxed_pair(const unsigned* p);

// One operand of a streamed unit: `cols` (16 or 32) columns starting at c0 of a [rows][ld_cols] array of {epoch, value}
// words, staged as floats into dst[row][cols].
struct WordTile { const unsigned* words; int ld_cols; int c0; };

// Stage NQ x (kThreads / (COLS/2)) rows per pass.  A thread keeps one 16-byte slot (two words) of the staged row and walks
// down the rows with all loads of the pass in flight (one L2 round trip); before asking for the pass it waits on ONE
// probe word (the last row it stages) so that a CTA that is free early does not stream its whole operand through L2 on
// every polling attempt; a pass that still holds a stale word is reloaded as a whole.
template <int COLS>
struct WordStager {
  static constexpr int kSlots = COLS / 2, kRowsPerPass = kThreads / kSlots;
  const unsigned* src; size_t ldw; float* dst; int row0;
  __device__ __forceinline__ WordStager(const WordTile& t, float* dst_base) {
    const int slot = threadIdx.x % kSlots;
    row0 = threadIdx.x / kSlots;
    src = t.words + 2 * (t.c0 + 2 * slot);
    ldw = 2 * static_cast<size_t>(t.ld_cols);
    dst = dst_base + 2 * slot;
  }
  template <int NQ>
  __device__ __forceinline__ void issue(int base, int rows, uint4 (&w)[NQ]) const {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int r = base + row0 + kRowsPerPass * q;
      if (r < rows) w[q] = ld_relaxed_quad(src + static_cast<size_t>(r) * ldw);
    }
  }
  template <int NQ>
  __device__ __forceinline__ bool fresh(int base, int rows, const uint4 (&w)[NQ], unsigned epoch) const {
    bool ok = true;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int r = base + row0 + kRowsPerPass * q;
      if (r < rows) ok = ok && (w[q].x == epoch) && (w[q].z == epoch);
    }
    return ok;
  }
  template <int NQ>
  __device__ __forceinline__ void store(int base, int rows, const uint4 (&w)[NQ]) const {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int r = base + row0 + kRowsPerPass * q;
      if (r < rows) *reinterpret_cast<float2*>(dst + r * COLS) = make_float2(__uint_as_float(w[q].y), __uint_as_float(w[q].w));
    }
  }
  // the last row of the pass starting at `base` that this thread stages (rows > base + row0 required)
  template <int NQ>
  __device__ __forceinline__ const unsigned* probe(int base, int rows) const {
    int last = base + row0 + kRowsPerPass * (NQ - 1);
    if (last >= rows) last = base + row0 + ((rows - 1 - base - row0) / kRowsPerPass) * kRowsPerPass;
    return src + static_cast<size_t>(last) * ldw;
  }
};
__device__ __forceinline__ void wait_word(const unsigned* p, unsigned epoch, SpinGuard& guard, volatile float* err) {
  while (ld_relaxed_pair(p).x != epoch) {
    __nanosleep(200);
    if (guard.expired()) { spin_report_timeout(err, epoch); break; }
  }
}

// A streamed gradient unit: out[m][n] = sum_b A[b][a0 + m] * Bm[b][b0 + n] over the batch rows (TMO x TNO outputs, one or
// two per thread), the bias gradient as the column sums of Bm, fused Adam (+ Polyak) on the owning thread -- the same
// arithmetic and summation order as wgrad_unit (warp w takes rows w, w+8, ...; warps combined in order), only the operands
// arrive as polled words instead of cp.async copies behind a barrier.
struct StreamUnit {
  WordTile A, Bm;
  int m_valid, n_valid;
  int out_base, out_sm, out_sn;    // param index = out_base + m*out_sm + n*out_sn
  int bias_base;                   // param index of bias[n] or -1
};
template <int TMO, int TNO>
__device__ void stream_unit(const AgentCtx& C, const StepScalars& S, const StreamUnit& U, float* smem) {
  constexpr int kRowsMax = kStreamTilesMax * kTM;      // <= 296 rows
  constexpr int LM = TMO / 8, LN = TNO / 4, NOUT = TMO * TNO / kThreads;
  float* As = smem;                                 // [rows][TMO]
  float* Bs = smem + kRowsMax * TMO;                // [rows][TNO]
  float* Ps = Bs + kRowsMax * TNO;                  // [kWarps][TMO*TNO]
  float* Pb = Ps + kWarps * TMO * TNO;              // [kWarps][TNO]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = static_cast<int>(S.B);
  int pi[NOUT], ps_off[NOUT];
  ParamVals pv[NOUT] = {};
#pragma unroll
  for (int q = 0; q < NOUT; ++q) {
    const int o = tid + q * kThreads, om = o / TNO, on = o % TNO;
    pi[q] = (om < U.m_valid && on < U.n_valid) ? U.out_base + om * U.out_sm + on * U.out_sn : -1;
    ps_off[q] = om * TNO + ((on + om / LM) & (TNO - 1));
    if (pi[q] >= 0) pv[q] = param_load(C, S, pi[q]);
  }
  const int bt = tid - (kThreads - 32);
  const int pbi = (U.bias_base >= 0 && bt >= 0 && bt < U.n_valid && bt < TNO) ? U.bias_base + bt : -1;
  ParamVals pvb{};
  if (pbi >= 0) pvb = param_load(C, S, pbi);
  __syncthreads();                                  // the staging area may still be in use by the CTA's previous phase
  RMC_STAMP(C, 16);
  {
    const WordStager<TMO> sa(U.A, As);
    const WordStager<TNO> sb(U.Bm, Bs);
    constexpr int kPass = (TNO > 16) ? 128 : 256;     // rows per pass (the wide unit halves it: its staging registers would double)
    constexpr int NQA = kPass / WordStager<TMO>::kRowsPerPass, NQB = kPass / WordStager<TNO>::kRowsPerPass;
    SpinGuard guard;
    for (int base = 0; base < rows; base += kPass) {
      const bool has_a = base + sa.row0 < rows, has_b = base + sb.row0 < rows;
      uint4 wa[NQA], wb[NQB];
      bool ok, first = true;
      do {
        // First attempt without asking: a CTA that comes here late (a row CTA after its own dz1 -- the chain that ends the
        // launch) finds its operands complete and saves the probe's L2 round trip.  A CTA that is early pays one wasted pass,
        // then waits on ONE probe word (B is the operand produced last) instead of streaming the operands on every attempt.
        if (has_a) sa.template issue<NQA>(base, rows, wa);
        if (has_b) sb.template issue<NQB>(base, rows, wb);
        ok = (!has_a || sa.template fresh<NQA>(base, rows, wa, S.epoch)) && (!has_b || sb.template fresh<NQB>(base, rows, wb, S.epoch));
        if (!ok) {
          if (first) {
            if (has_b) wait_word(sb.template probe<NQB>(base, rows), S.epoch, guard, C.host_loss);
          } else {
            __nanosleep(100);
          }
          first = false;
          if (guard.expired()) { spin_report_timeout(C.host_loss, S.epoch); break; }
        }
      } while (!ok);
      if (has_a) sa.template store<NQA>(base, rows, wa);
      if (has_b) sb.template store<NQB>(base, rows, wb);
    }
  }
  __syncthreads();
  RMC_STAMP(C, 17);
  const int mg = lane >> 2, ng = lane & 3;
  float acc[LM][LN];
  float bsum[LN];
#pragma unroll
  for (int i = 0; i < LM; ++i)
#pragma unroll
    for (int j = 0; j < LN; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int j = 0; j < LN; ++j) bsum[j] = 0.f;
#pragma unroll 8
  for (int r = warp; r < rows; r += kWarps) {       // warp w takes rows w, w+8, ... (fixed order -> deterministic sums)
    float a[LM], b[LN];
    lds_vec<LM>(a, As + r * TMO + mg * LM);
    lds_vec<LN>(b, Bs + r * TNO + ng * LN);
#pragma unroll
    for (int i = 0; i < LM; ++i)
#pragma unroll
      for (int j = 0; j < LN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
    for (int j = 0; j < LN; ++j) bsum[j] += b[j];
  }
#pragma unroll
  for (int i = 0; i < LM; ++i)
#pragma unroll
    for (int j = 0; j < LN; ++j) Ps[warp * (TMO * TNO) + (mg * LM + i) * TNO + ((ng * LN + j + mg) & (TNO - 1))] = acc[i][j];
  if (mg == 0) {
#pragma unroll
    for (int j = 0; j < LN; ++j) Pb[warp * TNO + ng * LN + j] = bsum[j];
  }
  __syncthreads();
  RMC_STAMP(C, 18);
  // The bias element of the last warp's lanes goes through the SAME straight-line Adam block as the thread's own outputs (its
  // sqrt / division chain interleaves with theirs; as a second, dependent chain it made the last warp -- and with it the
  // CTA -- finish 0.4 us late, on exactly the W0 units that end the launch).  Same arithmetic: param_math_n == param_apply.
  float gq[NOUT + 1];
  ParamVals pin[NOUT + 1], po[NOUT + 1];
#pragma unroll
  for (int q = 0; q < NOUT; ++q) {
    float g = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) g += Ps[w * (TMO * TNO) + ps_off[q]];
    gq[q] = g;
    pin[q] = pv[q];
  }
  {
    float g = 0.f;
    if (pbi >= 0) {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) g += Pb[w * TNO + bt];
    }
    gq[NOUT] = g;
    pin[NOUT] = pvb;
  }
  param_math_n<NOUT + 1>(S, gq, pin, po);
#pragma unroll
  for (int q = 0; q < NOUT; ++q) {
    if (pi[q] >= 0) {
      C.grads[pi[q]] = gq[q];
      param_store(C, S, pi[q], po[q]);
    }
  }
  if (pbi >= 0) {
    C.grads[pbi] = gq[NOUT];
    param_store(C, S, pbi, po[NOUT]);
  }
  RMC_STAMP(C, 19);
  __syncthreads();
}

// unit tables of the streamed phase B
__device__ __forceinline__ StreamUnit stream_unit_w2(const AgentCtx& C, int u) {      // 16 x 16, u in [0, 128)
  const NetLayout& L = C.L;
  const int kt = u / (kH2 / 16), jt = u % (kH2 / 16);
  StreamUnit U;
  U.A = WordTile{C.hp_words, kH1, kt * 16};
  U.Bm = WordTile{C.zp_words, kH2, jt * 16};
  U.m_valid = 16; U.n_valid = 16;
  U.out_base = L.off_w2t + kt * 16 * kW2LD + jt * 16; U.out_sm = kW2LD; U.out_sn = 1;
  U.bias_base = (kt == 0) ? L.off_b2 + jt * 16 : -1;
  return U;
}
__device__ __forceinline__ void stream_run_w2(const AgentCtx& C, const StepScalars& S, int u, float* smem) {
  stream_unit<16, 16>(C, S, stream_unit_w2(C, u), smem);
}
// two neighbouring W2 units as ONE 16 x 32 unit (two outputs per thread; same per-output summation order, hence the same bits)
__device__ __forceinline__ void stream_run_w2_wide(const AgentCtx& C, const StepScalars& S, int kt, int jt0, float* smem) {
  const NetLayout& L = C.L;
  StreamUnit U;
  U.A = WordTile{C.hp_words, kH1, kt * 16};
  U.Bm = WordTile{C.zp_words, kH2, jt0 * 16};
  U.m_valid = 16; U.n_valid = 32;
  U.out_base = L.off_w2t + kt * 16 * kW2LD + jt0 * 16; U.out_sm = kW2LD; U.out_sn = 1;
  U.bias_base = (kt == 0) ? L.off_b2 + jt0 * 16 : -1;
  stream_unit<16, 32>(C, S, U, smem);
}
__device__ __forceinline__ int stream_w0_count(const NetLayout& L) { return ((L.D + 15) / 16) * (kH1 / 16); }
__device__ __forceinline__ StreamUnit stream_unit_w0(const AgentCtx& C, int u) {      // 16 (d) x 16 (k)
  const NetLayout& L = C.L;
  const int mt = u / (kH1 / 16), nt = u % (kH1 / 16);
  StreamUnit U;
  U.A = WordTile{C.x_words, kMaxD, mt * 16};
  U.Bm = WordTile{C.z1_words, kH1, nt * 16};
  U.m_valid = min(16, L.D - mt * 16); U.n_valid = 16;
  U.out_base = L.off_w0t + mt * 16 * kH1 + nt * 16; U.out_sm = kH1; U.out_sn = 1;
  U.bias_base = (mt == 0) ? L.off_b0 + nt * 16 : -1;
  return U;
}
__device__ __forceinline__ StreamUnit stream_unit_heads(const AgentCtx& C, int u) {   // 16 (j) x 16 (a), u in [0, 8)
  const NetLayout& L = C.L;
  StreamUnit U;
  U.A = WordTile{C.h2_words, kH2, u * 16};
  U.Bm = WordTile{C.dh_words, kQLD, 0};
  U.m_valid = 16; U.n_valid = L.NH;
  U.out_base = L.off_wh + u * 16; U.out_sm = 1; U.out_sn = kH2;
  U.bias_base = (u == 0) ? L.off_bh : -1;
  return U;
}
// One schedule for all three kinds of unit.  Units in the order their operands come into existence -- heads (after the TD
// block), W2 (after dz2), W0 (after dz1) -- are dealt to the CTAs in the order those become free (see the call site).  All
// three kinds are the same 16 x 16 code with different operand tables: ONE copy of it per call site (the kernel is a long
// stretch of code that every CTA runs once, so its size is paid for in instruction fetch).
__device__ __forceinline__ void stream_run_any(const AgentCtx& C, const StepScalars& S, int id, float* smem) {
  constexpr int nH = kH2 / 16, nW2 = (kH1 / 16) * (kH2 / 16);
  const StreamUnit U = (id < nH) ? stream_unit_heads(C, id) : (id < nH + nW2) ? stream_unit_w2(C, id - nH) : stream_unit_w0(C, id - nH - nW2);
  stream_unit<16, 16>(C, S, U, smem);
}

// loss of a streamed launch without a priority write-back team: the tiles' {epoch, loss partial} words, summed in tile order
__device__ void stream_publish_loss(const AgentCtx& C, const StepScalars& S, int n_tiles, float* smem) {
  float* s_lp = smem;
  const int tid = threadIdx.x;
  __syncthreads();
  float lp = 0.f;
  if (tid < n_tiles) {
    SpinGuard guard;
    uint2 w = ld_relaxed_pair(C.qt_flag + 512 + 2 * tid);      // kLossWordBase
    while (w.x != S.epoch) {
      __nanosleep(100);
      w = ld_relaxed_pair(C.qt_flag + 512 + 2 * tid);
      if (guard.expired()) { spin_report_timeout(C.host_loss, S.epoch); break; }
    }
    lp = __uint_as_float(w.y);
  }
  s_lp[tid] = lp;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int c = 0; c < n_tiles; ++c) s += s_lp[c];
    const float loss = s / static_cast<float>(S.Bglobal);
    C.loss[0] = loss;
    host_loss_store(C.host_loss, loss, S.epoch);
  }
  __syncthreads();
}

// ------------------------------------------------------------------ the fused learner step
// Layout of AgentCtx::qt_flag (kFlagWords u32 per agent):
//   [0, 512)      (unused)
//   [512, 1024)   per tile {epoch, loss partial}            (early write-back, 8-byte words)
//   [1024, 4096)  per batch row {epoch, |td|}               (early write-back, 8-byte words)
//   [4096, ...)   per (tile, row, action) {epoch, Q_target(s')}   (role split: target CTA -> row CTA)
// The 8-byte words carry their payload WITH the flag (one relaxed vector store), so publishing them needs no fence on
// the row CTA's critical path; the stores they must be ordered after (leaf indices and old leaf values from the
// sampling phase) are covered by a fence that the publishing warp executes while warp 0 computes the TD block.
constexpr int kLossWordBase = 512, kTdWordBase = 1024, kQtWordBase = 4096;   // [4096, 4096 + 2*74*64): {epoch, Q_target} per (tile, row, action)
constexpr int kEarlyMaxRows = (kQtWordBase - kTdWordBase) / 2;
static_assert(kQtWordBase + 2 * 74 * kTM * kQLD <= kFlagWords, "flag words");
__device__ __forceinline__ void st_relaxed_pair(unsigned* p, unsigned a, unsigned b) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_relaxed_pair(const unsigned* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

// Who does what after the A->B barrier (identical on every CTA of the agent).
struct PhaseBPlan { bool tree_here, team, coarse, early; int n_workers; };
__device__ __forceinline__ PhaseBPlan phase_b_plan(const AgentCtx& C, const StepScalars& S, int G, bool per) {
  PhaseBPlan p;
  p.tree_here = per && (S.phases & 4) && C.rp.prioritized && S.B <= kTreeCtaMax;
  p.coarse = G < 100;   // few CTAs per agent (ensembles): 32x32 gradient tiles instead of 16x16
  // big tree + enough CTAs: a team of kTreeTeam CTAs shares the write-back; otherwise one CTA does it
  p.team = p.tree_here && (2 * C.rp.cap - 1 >= 2 * kTopRebuild + 1) && (G >= wgrad_unit_count(C.L, false) + kTreeTeam);
  const int n_tree = !p.tree_here ? 0 : (p.team ? kTreeTeam : 1);
  p.n_workers = (p.tree_here && G > 1) ? G - n_tree : G;
  p.early = p.tree_here && G > 1 && (S.phases & 2) && S.B <= kEarlyMaxRows;   // |td| is produced by this launch's row CTAs
  return p;
}

// Streamed phase B (one tile per row CTA + role split + FORWARD and BACKWARD in one launch + at least kStreamIdleMin CTAs that
// own neither a row tile nor a target tile nor tree work):
//   * the row CTAs publish what the weight gradients need as {epoch, value} words the moment it exists -- x with the
//     sampled rows, H1 / H2 right after the forward pass, the head deltas after the TD block, dz2 and dz1 as they are formed;
//   * every gradient unit polls its operand words and starts as soon as its CTA is free: W2 units (16 x 16 outputs, the
//     bulk of the parameters, one per row / target CTA) on target CTAs right after their Q values left -- they used to idle
//     until the agent barrier -- and on row CTAs right after dz1; the head units and the W0 units on the CTAs that are idle
//     in phase A (heads first: their operands exist 3 us before dz1);
//   * nobody waits at a barrier and nobody executes a fence between the two phases.
// Same arithmetic and summation order as the barrier path (wgrad_unit).  The agent barrier is only arrived at (the host
// counts on G arrivals per launch).
constexpr int kStreamIdleMin = 8;
struct StreamPlan { bool on; int idle0, n_idle; };
__device__ __forceinline__ StreamPlan stream_plan(const AgentCtx& C, const StepScalars& S, const PhaseBPlan& pb, int G, long long n_tiles, bool one_tile,
                                                  bool split) {
  StreamPlan sp;
  sp.idle0 = 2 * static_cast<int>(n_tiles);
  sp.n_idle = pb.n_workers - sp.idle0;
  sp.on = one_tile && split && !pb.coarse && (S.phases & 8) && sp.n_idle >= kStreamIdleMin && n_tiles <= kStreamTilesMax && C.x_words != nullptr &&
          (!pb.tree_here || pb.early) &&
          (kH2 / 16 + (kH1 / 16) * (kH2 / 16) + ((C.L.D + 15) / 16) * (kH1 / 16)) - (sp.n_idle + 2 * static_cast<int>(n_tiles)) <= sp.n_idle;
  return sp;
}

// loss = (1/B) sum of the per-tile partials in tile order; also into mapped host memory (value, system fence, epoch)
__device__ __forceinline__ void publish_loss(const AgentCtx& C, const StepScalars& S, const float* parts, int np) {
  float s = 0.f;
  for (int c0 = 0; c0 < np; c0 += 16) {      // 16 loads in flight, then the adds in tile order (a plain loop is one L2 round trip per partial)
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = (c0 + q < np) ? parts[c0 + q] : 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += (c0 + q < np) ? v[q] : 0.f;
  }
  const float loss = s / static_cast<float>(S.Bglobal);
  C.loss[0] = loss;
  host_loss_store(C.host_loss, loss, S.epoch);
}
// Three instantiations, chosen by the host (step_path in rmc_b200.cu): kPath = 1 when every row CTA owns at most one 4-row
// tile (the default single-agent batches: rows and Q_target stay in shared memory, role split, streamed phase B), 0 for
// several 4-row tiles per CTA (ensembles, mid-size batches), 2 for the batch-stationary phases of rmc_rows_ws.cuh (a row
// CTA owns at least two 16-row tiles).  Splitting them keeps each kernel's straight-line code small: the step is sensitive to
// instruction-fetch stalls.
template <int kPath>
__global__ void __launch_bounds__(kThreads, 1) k_learner_step(const __grid_constant__ AgentCtx single, const AgentCtx* __restrict__ many,
                                                              const __grid_constant__ StepScalars S) {
  constexpr bool kOneTile = kPath == 1;
  extern __shared__ __align__(16) float smem[];
  // private copy of the context: field reads become register / local-memory accesses that the compiler can hoist
  // (through a reference into parameter-or-global memory every C.x is a generic load that no store may cross)
  const AgentCtx Cv = (many != nullptr) ? many[blockIdx.y] : single;
  const AgentCtx& C = Cv;
  const unsigned agent = blockIdx.y;
  const NetLayout L = C.L;
  const SmemPlan P = make_smem_plan(L.total);
  float* sW = smem + P.w;
  float* sXT = smem + P.xt;
  float* sH1T = smem + P.h1t;
  float* sH2 = smem + P.h2;
  float* sPart = smem + P.part;
  float* sQ = smem + P.q;
  float* sDZ2 = smem + P.dz2;
  float* sDH = smem + P.dh;
  float* sMeta = smem + P.meta;
  float* sRed = smem + P.red;
  float* sRows = smem + P.rows;
  float* sQT = smem + P.qt;
  float* sIsw = smem + P.isw;
  double* sTop = reinterpret_cast<double*>(smem + P.top);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P.bar);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const long long B = S.B;
  const long long n_tiles = (B + kTM - 1) / kTM;
  const int rf = C.rp.row_floats;
  const int D = L.D;
  const bool do_rows = (S.phases & (1 | 2)) != 0;
  const bool per = S.prioritized != 0;
  const long long first_leaf = C.rp.cap - 1;
  uint32_t parity = 0;
  // programmatic dependent launch (no-ops for ordinary / cooperative launches): let the next launch be
  // scheduled early, and order everything below after the previous grid's memory operations
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  RMC_STAMP(C, 0);
  // launch-gap diagnostic: per-launch [min start, max end] in dbg[kDbgCtas*kDbgSlots + 2*(epoch % 64) ..] (enabled with the stamps)
  unsigned long long* gap = (C.dbg != nullptr) ? C.dbg + kDbgCtas * kDbgSlots + 2 * (S.epoch & 63u) : nullptr;
  if (gap != nullptr && threadIdx.x == 0) atomicMin(gap, global_timer_ns());

  // Role split: when every row CTA owns one tile and enough CTAs are idle in phase A, CTA n_tiles+t computes
  // Q_target(s') of tile t concurrently (it repeats tile t's sampling -- same uniforms, same leaves -- and stages
  // only the target blob), while row CTA t stages only the online blob and picks Q_target up through a flag.
  constexpr bool one_tile = kOneTile;               // == (n_tiles <= S.n_row_ctas), chosen by the host: rows / Q_target stay in smem
  const bool split = one_tile && ((S.phases & 3) == 3) && (static_cast<long long>(G) >= 2 * n_tiles);
  const bool is_row = do_rows && cta < S.n_row_ctas && cta < n_tiles;
  const bool is_tgt = do_rows && split && cta >= n_tiles && cta < 2 * n_tiles;
  const int tile0 = is_tgt ? cta - static_cast<int>(n_tiles) : cta;
  // Early write-back: the priority write-back needs only (leaf, |td|), which exist right after the TD block -- long before
  // the row CTAs finish dgrad and reach the agent barrier.  With one tile per row CTA every row CTA release-stores the
  // launch's epoch into its tile's flag after TD; the tree CTAs arrive at the barrier without waiting, poll the flags and
  // start the write-back (and publish the loss) while dgrad is still running.  All flags set => every sampler of this
  // step (row CTAs and, through the Q_target hand-off, their target partners) has finished its descent.
  const PhaseBPlan pb = phase_b_plan(C, S, G, per);
  const bool early_td = one_tile && pb.early;
  const StreamPlan sp = stream_plan(C, S, pb, G, n_tiles, one_tile, split);
  const bool stream_b = kOneTile && sp.on;
  // Several row tiles per CTA: the batch-stationary phases of rmc_rows_ws.cuh (kPath 2) once a row CTA owns at least two 16-row
  // tiles (large batches; chosen by the host, step_path in rmc_b200.cu); below that (the 8-agent ensemble launch: 16 rows
  // per CTA) 4-row tiles through the latency-shaped passes are faster -- every phase would run once, on cold code, and the
  // partial-gradient exchange would cost more than the gradient units (measured: 83-87 us against 73 us for 8 x 256).
  constexpr bool use_ws = kPath == 2;
  if (is_row || is_tgt) {
    const long long n_nodes = 2 * C.rp.cap - 1;
    const int n_top = static_cast<int>(min(static_cast<long long>(kTopNodes), n_nodes));
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    // replay state and tree total: issued now so that their latency overlaps the top-of-tree load below
    const long long size = __ldcg(&C.rp.st->size), dp = __ldcg(&C.rp.st->dp);
    const double total = per ? __ldcg(C.rp.tree) : 0.0;
    const float min_p_f = per ? __ldcg(&C.rp.st->min_p) : 0.f;
    // top levels of the tree: one coalesced read, then smem descents.  The loads are issued first and the sample's uniform
    // (Philox, independent of memory) is drawn while they are in flight.
    double pre_u = 0.0;
    {
      double topv[(kTopNodes + kThreads - 1) / kThreads];
      const bool want_top = (S.phases & 1) && C.rp.prioritized;
#pragma unroll
      for (int q = 0; q < (kTopNodes + kThreads - 1) / kThreads; ++q) {
        const int t = tid + q * kThreads;
        topv[q] = (want_top && t < n_top) ? __ldcg(C.rp.tree + t) : 0.0;
      }
      if (kOneTile && want_top && warp < kTM) {
        const long long i = static_cast<long long>(tile0) * kTM + warp;
        if (i < B) pre_u = (S.u != nullptr) ? S.u[agent * B + i] : philox_uniform(S.seed, S.counter, agent, static_cast<uint32_t>(S.shard_off + i));
      }
#pragma unroll
      for (int q = 0; q < (kTopNodes + kThreads - 1) / kThreads; ++q) {
        const int t = tid + q * kThreads;
        if (want_top && t < n_top) sTop[t] = topv[q];
      }
    }
    __syncthreads();
    RMC_STAMP(C, 8);
    const bool do_fwd = (S.phases & 2) != 0;
    if (do_fwd) stage_params(sW, (split && !is_tgt) ? C.online : C.target, L.total, bar, parity);

    // -------- pass 1 over this CTA's tiles: sample + gather, then Q_target(s')
    // warps 0..kTM-1: one sample each; warp kTM meanwhile computes the max IS weight (replay_memory.py:76-77),
    // which the samplers pick up at a 5-warp named barrier after their own descent.
    const bool tree_sampling = C.rp.prioritized != 0;
    // Importance weights (float64 pow, ~0.6 us): with the role split the TARGET partner evaluates them -- it has about a
    // microsecond of slack before its Q values are needed -- and ships them in the spare 16th word of each Q_target row;
    // the row CTA goes from the descent straight to the forward pass.
    const bool isw_by_partner = split && (S.phases & 1) && tree_sampling;
    const bool does_isw = isw_by_partner ? is_tgt : !is_tgt;
    if ((S.phases & 1) && !one_tile) {
      // several tiles per CTA (ensembles, large batches without the grid-wide sampler): all 8 warps draw this CTA's
      // samples, 8 descents in flight instead of 4; rows go straight to the L2-resident scratch
      // (the max importance weight -- replay_memory.py:76-77 -- is evaluated per warp by a third lane of the SAME float64 pow
      //  call that serves the warp's two samples: no pow and no CTA barrier in front of the descents)
      // this CTA samples the rows it will process: 16- / 8-row tiles (rmc_rows_ws.cuh) or 4-row tiles
      const int wrs = use_ws ? ws_rows(B, S.n_row_ctas) : kTM;
      const long long n_wide = (B + wrs - 1) / wrs;
      const long long my_wide = (cta < n_wide) ? (n_wide - cta + S.n_row_ctas - 1) / S.n_row_ctas : 0;
      // a warp takes its samples two at a time: both descents share every memory round trip (per_descend_cached2), both rows
      // are requested together and the two float64 pows of the importance weights overlap
      for (long long s = warp; s < my_wide * wrs; s += 2 * kWarps) {
        const long long sb = s + kWarps;
        const long long ia = (cta + (s / wrs) * S.n_row_ctas) * wrs + (s % wrs);
        const long long ib = (cta + (sb / wrs) * S.n_row_ctas) * wrs + (sb % wrs);
        const bool has_a = ia < B, has_b = sb < my_wide * wrs && ib < B;
        if (!has_a && !has_b) continue;
        long long slot_a = 0, node_a = 0, slot_b = 0, node_b = 0;
        double p_a = 0.0, p_b = 0.0, num_a = 1.0, num_b = 1.0;
        if (tree_sampling) {
          const long long ga = S.shard_off + ia, gb = S.shard_off + ib;
          const double ua = !has_a ? 0.0 : (S.u != nullptr) ? S.u[agent * B + ia] : philox_uniform(S.seed, S.counter, agent, static_cast<uint32_t>(ga));
          const double ub = !has_b ? 0.0 : (S.u != nullptr) ? S.u[agent * B + ib] : philox_uniform(S.seed, S.counter, agent, static_cast<uint32_t>(gb));
          // (a missing first sample can only be the ragged end of the batch, where the second is missing as well)
          per_descend_cached2(sTop, n_top, C.rp.tree, n_nodes, stratum_value(total, S.Bglobal, ga, ua), stratum_value(total, S.Bglobal, gb, ub),
                              has_b, &node_a, &p_a, &node_b, &p_b);
          slot_a = node_a - first_leaf;
          slot_b = node_b - first_leaf;
        } else {
          const long long pos_a = (S.idx != nullptr) ? S.idx[agent * B + ia]
                                                     : static_cast<long long>(feistel_perm(S.shard_off + ia, size, S.seed, S.counter, agent));
          slot_a = deque_pos_to_slot(pos_a, size, dp, C.rp.cap);
          node_a = slot_a;
          if (has_b) {
            const long long pos_b = (S.idx != nullptr) ? S.idx[agent * B + ib]
                                                       : static_cast<long long>(feistel_perm(S.shard_off + ib, size, S.seed, S.counter, agent));
            slot_b = deque_pos_to_slot(pos_b, size, dp, C.rp.cap);
            node_b = slot_b;
          }
        }
        const RowRegs ra = gather_row_load(C.rp, slot_a);          // row loads in flight during the pows
        const RowRegs rb = gather_row_load(C.rp, has_b ? slot_b : slot_a);
        double w_max = 1.0;
        if (tree_sampling) {      // ONE pow call per warp: lane 0 -> sample a, lane 1 -> sample b, lane 2 -> the max weight (min priority)
          const double pr = (lane == 0) ? p_a : (lane == 1) ? (has_b ? p_b : p_a) : static_cast<double>(min_p_f);
          const double y = pow(static_cast<double>(size) * (pr / total), -S.beta);
          num_a = __shfl_sync(0xffffffffu, y, 0);
          num_b = __shfl_sync(0xffffffffu, y, 1);
          w_max = __shfl_sync(0xffffffffu, y, 2);
        }
        {
          float* dst = C.X + ia * rf;
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (lane + 32 * q < rf) dst[lane + 32 * q] = ra.v[q];
          if (lane == 0) {
            C.nodes[ia] = node_a;
            C.is_w[ia] = tree_sampling ? static_cast<float>(num_a / w_max) : 1.f;
            C.leaf_p[ia] = p_a;
          }
        }
        if (has_b) {
          float* dst = C.X + ib * rf;
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (lane + 32 * q < rf) dst[lane + 32 * q] = rb.v[q];
          if (lane == 0) {
            C.nodes[ib] = node_b;
            C.is_w[ib] = tree_sampling ? static_cast<float>(num_b / w_max) : 1.f;
            C.leaf_p[ib] = p_b;
          }
        }
      }
    } else if (S.phases & 1) {
      bool first_iter = true;
      for (long long tile = tile0; tile < n_tiles; tile += S.n_row_ctas) {
        if (warp < kTM) {
          const long long i = tile * kTM + warp;
          const bool ok = i < B;
          long long slot = 0, node = 0;
          double p = 0.0, numer = 1.0;
          RowRegs rr{};
          if (ok) {
            if (tree_sampling) {
              const long long gi = S.shard_off + i;
              const double ui = kOneTile ? pre_u
                                         : ((S.u != nullptr) ? S.u[agent * B + i] : philox_uniform(S.seed, S.counter, agent, static_cast<uint32_t>(gi)));
              const double v = stratum_value(total, S.Bglobal, gi, ui);
              RMC_STAMP(C, 9);
              node = per_descend_cached(sTop, n_top, C.rp.tree, n_nodes, v, &p);
              slot = node - first_leaf;
              RMC_STAMP(C, 10);
            } else {
              const long long pos = (S.idx != nullptr) ? S.idx[agent * B + i]
                                                       : static_cast<long long>(feistel_perm(S.shard_off + i, size, S.seed, S.counter, agent));
              slot = deque_pos_to_slot(pos, size, dp, C.rp.cap);
              node = slot;
            }
            rr = gather_row_load(C.rp, slot);                      // row loads in flight during the pow
            if (tree_sampling && does_isw) numer = pow(static_cast<double>(size) * (p / total), -S.beta);
            RMC_STAMP(C, 11);
          }
          if (tree_sampling && does_isw) asm volatile("bar.sync 1, %0;" ::"n"((kTM + 1) * 32) : "memory");
          if (ok) {
            const float w = (tree_sampling && does_isw) ? static_cast<float>(numer / sTop[kTopNodes]) : 1.f;
            if (is_tgt) {                                          // target role: rows stay in shared memory only
              gather_row_store_smem(C.rp, rr, sRows + warp * kMaxRowFloats);
              if (lane == 0) sIsw[warp] = w;
            } else {
              gather_row_store(C.rp, rr, C.X + i * rf, sRows + warp * kMaxRowFloats);
              if (lane == 0) { C.nodes[i] = node; sIsw[warp] = w; C.leaf_p[i] = p; if (does_isw) C.is_w[i] = w; }
            }
            RMC_STAMP(C, 12);
          }
        } else if (warp == kTM && tree_sampling && does_isw) {
          if (first_iter && lane == 0)
            sTop[kTopNodes] = is_weight_max(static_cast<double>(size), total, static_cast<double>(min_p_f), S.beta);
          __syncwarp();
          asm volatile("bar.sync 1, %0;" ::"n"((kTM + 1) * 32) : "memory");
        }
        first_iter = false;
      }
    }
    __syncthreads();   // X rows written by this CTA are visible to it
    RMC_STAMP(C, 1);
    if (do_fwd) {
      wait_params(bar, parity);
      RMC_STAMP(C, 2);
      if (!one_tile && !use_ws) {
        // several tiles per CTA: Q_target(s') of TWO tiles per pass (8 rows share one sweep over the target weights)
        for (long long ta = tile0; ta < n_tiles; ta += 2 * S.n_row_ctas) {
          const long long tb = ta + S.n_row_ctas;
          for (int t = tid; t < kR * D; t += kThreads) {
            const int r = t / D, d = t % D;
            const long long tl = (r < kTM) ? ta : tb;
            const long long i = tl * kTM + (r % kTM);
            sXT[d * kR + r] = (tl < n_tiles && i < B) ? __ldcg(C.X + i * rf + D + d) : 0.f;
          }
          __syncthreads();
          mlp_forward<kR>(sW, L, sXT, sH1T, sH2, sPart, sQ, nullptr);
          if (tid < kR * kQLD) {
            const int r = tid / kQLD;
            const long long tl = (r < kTM) ? ta : tb;
            const long long i = tl * kTM + (r % kTM);
            if (tl < n_tiles && i < B) C.QT[i * kQLD + (tid % kQLD)] = sQ[tid];
          }
          __syncthreads();
        }
      } else if (!one_tile) {
        // batch-stationary passes (rmc_rows_ws.cuh): both forward phases run below, from one copy of the code
      } else if (!split || is_tgt)
      for (long long tile = tile0; tile < n_tiles; tile += S.n_row_ctas) {
        // x^T of the s' rows -> sXT[d][0..3]
        for (int t = tid; t < kTM * D; t += kThreads) {
          const int r = t / D, d = t % D;
          const long long i = tile * kTM + r;
          sXT[d * kR + r] = (i < B) ? (one_tile ? sRows[r * kMaxRowFloats + D + d] : __ldcg(C.X + i * rf + D + d)) : 0.f;
        }
        __syncthreads();
        mlp_forward<kTM>(sW, L, sXT, sH1T, sH2, sPart, sQ, nullptr);
        if (tid < kTM * kQLD) {
          const int r = tid / kQLD;
          const long long i = tile * kTM + r;
          sQT[tid] = sQ[tid];
          if (i < B) C.QT[i * kQLD + (tid % kQLD)] = sQ[tid];
        }
        __syncthreads();
      }
      // publish Q_target(s') of the tile as {epoch, value} words (payload rides with the flag: no fence, one round trip
      // on the consumer's side) -- then on to phase B
      if (is_tgt && tid < kTM * kQLD) {
        const float payload = (isw_by_partner && (tid % kQLD) == kQLD - 1) ? sIsw[tid / kQLD] : sQT[tid];      // A <= 15: column 15 is spare
        st_relaxed_pair(C.qt_flag + kQtWordBase + 2 * (tile0 * kTM * kQLD + tid), S.epoch, __float_as_uint(payload));
      }
      // -------- pass 2: online weights; [s'; s] rows
      RMC_STAMP(C, 3);
      if (!use_ws && !split) {
        stage_params(sW, C.online, L.total, bar, parity);
        wait_params(bar, parity);
      }
      RMC_STAMP(C, 4);
      float loss_local = 0.f;   // thread 0 accumulates this CTA's tiles in order
      if (use_ws) {
        // online pass, dgrad and the weight gradients of this CTA's rows; one partial gradient blob per CTA (summed after
        // the agent barrier by ws_reduce_apply)
        const WsTiles Tw = ws_tiles(S, cta);
        const bool has_rows = Tw.first < Tw.n_wide;
        const bool grads = (S.phases & 8) && C.gpart != nullptr;
        float* gp = (grads && has_rows) ? C.gpart + static_cast<size_t>(cta) * L.total : nullptr;
        const WsSmem Wd = ws_smem(smem + P.xt);
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {     // target pass, then online pass: ONE copy of the forward code
          if (ph == 1) {
            stage_params(sW, C.online, L.total, bar, parity);
            wait_params(bar, parity);
            RMC_STAMP(C, 4);
          }
          if (has_rows) {
            const float lp = ws_rows_phase(C, S, sW, Wd, Tw, per, gp, ph == 0);
            if (ph == 1) loss_local = lp;
          }
          if (ph == 0) RMC_STAMP(C, 3);
        }
        RMC_STAMP(C, 13);
        if (has_rows) {
          if (L.D <= 16) ws_dgrad_phase<16>(C, S, sW, sW, Wd, Tw, gp);
          else ws_dgrad_phase<kMaxD>(C, S, sW, sW, Wd, Tw, gp);
          RMC_STAMP(C, 14);
          if (gp != nullptr) ws_wgrad_phase(C, S, sW, Wd, Tw, gp);
        }
      } else if (!is_tgt)
      for (long long tile = cta; tile < n_tiles; tile += S.n_row_ctas) {
        for (int t = tid; t < kR * D; t += kThreads) {
          const int r = t / D, d = t % D;
          const long long i = tile * kTM + (r % kTM);
          const int col = (r < kTM) ? (D + d) : d;   // rows 0..3: s', rows 4..7: s
          sXT[d * kR + r] = (i < B) ? (one_tile ? sRows[(r % kTM) * kMaxRowFloats + col] : __ldcg(C.X + i * rf + col)) : 0.f;
        }
        if (tid < kTM) {
          const long long i = tile * kTM + tid;
          const bool ok = i < B;
          const float* row = sRows + tid * kMaxRowFloats;
          sMeta[tid * 4 + 0] = ok ? (one_tile ? row[2 * D] : __ldcg(C.X + i * rf + 2 * D)) : 0.f;
          sMeta[tid * 4 + 1] = ok ? (one_tile ? row[2 * D + 1] : __ldcg(C.X + i * rf + 2 * D + 1)) : 0.f;
          sMeta[tid * 4 + 2] = ok ? (one_tile ? row[2 * D + 2] : __ldcg(C.X + i * rf + 2 * D + 2)) : 0.f;
          sMeta[tid * 4 + 3] = ok ? (one_tile ? sIsw[tid] : __ldcg(C.is_w + i)) : 0.f;
        }
        __syncthreads();
        mlp_forward<kR>(sW, L, sXT, sH1T, sH2, sPart, sQ, nullptr);
        RMC_STAMP(C, 13);
        if (stream_b) {  // x, H1 and H2 of the s rows leave now, long before the gradient units ask for them
          const float4 hv = *reinterpret_cast<const float4*>(sH1T + tid * kR + kTM);
          const float h[kTM] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int r = 0; r < kTM; ++r) {
            const long long i = tile * kTM + r;
            if (i < B) st_relaxed_pair(C.hp_words + 2 * (i * kH1 + tid), S.epoch, __float_as_uint(h[r]));
          }
          {
            const int j = tid & (kH2 - 1), r0 = tid >> 7;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int r = r0 + 2 * rr;
              const long long i = tile * kTM + r;
              if (i < B) st_relaxed_pair(C.h2_words + 2 * (i * kH2 + j), S.epoch, __float_as_uint(sH2[(kTM + r) * kH2 + j]));
            }
          }
          if (tid < kTM * kMaxD) {
            const int r = tid / kMaxD, d = tid % kMaxD;
            const long long i = tile * kTM + r;
            if (i < B) st_relaxed_pair(C.x_words + 2 * (i * kMaxD + d), S.epoch, __float_as_uint(d < D ? sXT[d * kR + kTM + r] : 0.f));
          }
        }
        if (split) {     // Q_target(s') of this tile comes from its partner CTA
          if (tid < kTM * kQLD) {
            const unsigned* wp = C.qt_flag + kQtWordBase + 2 * (tile * kTM * kQLD + tid);
            uint2 w = ld_relaxed_pair(wp);
            SpinGuard guard;
            while (w.x != S.epoch) {
              __nanosleep(20);
              w = ld_relaxed_pair(wp);
              if (guard.expired()) { spin_report_timeout(C.host_loss, S.epoch); break; }
            }
            sQT[tid] = __uint_as_float(w.y);
            if (isw_by_partner && (tid % kQLD) == kQLD - 1) {      // the partner's importance weight of row tid / kQLD
              const int r = tid / kQLD;
              const long long i = tile * kTM + r;
              sMeta[r * 4 + 3] = (i < B) ? __uint_as_float(w.y) : 0.f;
              if (i < B) C.is_w[i] = __uint_as_float(w.y);
            }
          }
          __syncthreads();
        }
        // early write-back: the publishing lanes order the sampling-phase stores of this CTA (nodes, leaf_p; made visible to
        // them by the barriers since) before their flag words -- executed in the shadow of warp 0's TD block
        if (early_td && warp == 1 && lane <= kTM) __threadfence();
        // ---- TD target, |td|, Huber, dQ coefficient (threads 0..kTM-1)
        if (tid < kTM) {
          const int r = tid;
          const long long i = tile * kTM + r;
          float g = 0.f, lterm = 0.f, atd_pub = 0.f;
          int act = 0;
          if (i < B) {
            float qtv[kQLD];
#pragma unroll
            for (int a = 0; a < kQLD; ++a) qtv[a] = one_tile ? sQT[r * kQLD + a] : __ldcg(C.QT + i * kQLD + a);
            float qsel;
            if (S.double_dqn) {                                   // dqn/agent.py:252-256
              const int astar = argmax_first(sQ + r * kQLD, L.A);
              qsel = qtv[0];
#pragma unroll
              for (int a = 1; a < kQLD; ++a) qsel = (a == astar) ? qtv[a] : qsel;
            } else {                                              // dqn/agent.py:172-173
              qsel = qtv[0];
#pragma unroll
              for (int a = 1; a < kQLD; ++a) qsel = (a < L.A) ? fmaxf(qsel, qtv[a]) : qsel;
            }
            act = __float_as_int(sMeta[r * 4 + 0]);
            const float rew = sMeta[r * 4 + 1], done = sMeta[r * 4 + 2], w = sMeta[r * 4 + 3];
            const float y = rew + ((1.f - done) * S.gamma) * qsel;   // dqn/agent.py:258
            const float q_sa = sQ[(kTM + r) * kQLD + act];
            const float delta = q_sa - y;
            const float atd = fabsf(y - q_sa);                     // dqn/agent.py:264
            const float z = fabsf(delta);
            const float hub = (z < 1.f) ? (0.5f * z) * z : z - 0.5f;   // SmoothL1, beta = 1
            const float go = per ? (1.f / static_cast<float>(S.Bglobal)) * w : 1.f / static_cast<float>(S.Bglobal);
            g = fminf(fmaxf(delta, -1.f), 1.f) * go;
            lterm = per ? w * hub : hub;
            atd_pub = atd;
            C.y[i] = y; C.q_sa[i] = q_sa; C.abs_td[i] = atd; C.hub[i] = hub; C.gcoef[i] = g;
            // (|td| -> priority and the last-writer stamps are produced by the write-back itself, phase B)
          }
          // dheads
          float* dh = sDH + r * kQLD;
          for (int a = 0; a < kQLD; ++a) dh[a] = 0.f;
          if (L.dueling) {
            const float mean = g / static_cast<float>(L.A);
            dh[0] = g;                                             // dval = sum_a dQ
            for (int a = 0; a < L.A; ++a) dh[1 + a] = ((a == act) ? g : 0.f) - mean;
          } else {
            dh[act] = g;
          }
          sRed[r] = lterm;
          sRed[kTM + r] = atd_pub;
        }
        // debug / parity outputs of the Q rows
        if (tid < kR * kQLD) {
          const int r = tid / kQLD;
          const long long i = tile * kTM + (r % kTM);
          if (i < B) ((r < kTM) ? C.QN : C.Q)[i * kQLD + (tid % kQLD)] = sQ[tid];
        }
        __syncthreads();
        RMC_STAMP(C, 14);
        if (tid == 0) {
          for (int r = 0; r < kTM; ++r) loss_local += sRed[r];
        }
        if ((early_td || stream_b) && warp == 1 && lane <= kTM) {     // release the write-back team: {epoch, |td|} per row, {epoch, loss partial}
          if (lane < kTM) {
            const long long i = tile * kTM + lane;
            if (i < B) st_relaxed_pair(C.qt_flag + kTdWordBase + 2 * i, S.epoch, __float_as_uint(sRed[kTM + lane]));
          } else {
            float lp = 0.f;
            for (int r = 0; r < kTM; ++r) lp += sRed[r];
            st_relaxed_pair(C.qt_flag + kLossWordBase + 2 * tile, S.epoch, __float_as_uint(lp));
          }
        }
        // ---- dh2 -> dz2 (thread: column j, rows r and r+2)
        {
          const int j = tid & (kH2 - 1), r0 = tid >> 7;   // r0 in {0,1}
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int r = r0 + 2 * rr;
            const float* dh = sDH + r * kQLD;
            float s = 0.f;
            for (int a = 0; a < L.NH; ++a) s = fmaf(dh[a], sW[L.off_wh + a * kH2 + j], s);
            const float h2v = sH2[(kTM + r) * kH2 + j];
            const float dz = act_bwd(s, h2v, L.act);
            sDZ2[r * kH2 + j] = dz;
            const long long i = tile * kTM + r;
            if (stream_b) {
              if (i < B) st_relaxed_pair(C.zp_words + 2 * (i * kH2 + j), S.epoch, __float_as_uint(dz));
            } else if (i < B) { C.DZ2[i * kH2 + j] = dz; C.H2[i * kH2 + j] = h2v; }
          }
          if (stream_b) {
            if (tid < kTM * kQLD) {
              const long long i = tile * kTM + tid / kQLD;
              if (i < B) st_relaxed_pair(C.dh_words + 2 * (i * kQLD + (tid % kQLD)), S.epoch, __float_as_uint(sDH[tid]));
            }
          } else if (tid < kTM * kQLD) {
            const long long i = tile * kTM + tid / kQLD;
            if (i < B) C.DH[i * kQLD + (tid % kQLD)] = sDH[tid];
          }
        }
        __syncthreads();
        RMC_STAMP(C, 15);
        // ---- dz1[r][k] = (sum_j dz2[r][j] W2^T[k][j]) * [h1 > 0]   (thread k)
        {
          float acc[kTM] = {0.f, 0.f, 0.f, 0.f};
          const float* wrow = sW + L.off_w2t + tid * kW2LD;
#pragma unroll 4
          for (int j = 0; j < kH2; j += 4) {
            const float4 w = *reinterpret_cast<const float4*>(wrow + j);
#pragma unroll
            for (int r = 0; r < kTM; ++r) {
              const float4 dz = *reinterpret_cast<const float4*>(sDZ2 + r * kH2 + j);
              acc[r] = fmaf(dz.x, w.x, acc[r]);
              acc[r] = fmaf(dz.y, w.y, acc[r]);
              acc[r] = fmaf(dz.z, w.z, acc[r]);
              acc[r] = fmaf(dz.w, w.w, acc[r]);
            }
          }
          if (stream_b) {
#pragma unroll
            for (int r = 0; r < kTM; ++r) {
              const long long i = tile * kTM + r;
              if (i < B) st_relaxed_pair(C.z1_words + 2 * (i * kH1 + tid), S.epoch, __float_as_uint(act_bwd(acc[r], sH1T[tid * kR + kTM + r], L.act)));
            }
          } else {
#pragma unroll
          for (int r = 0; r < kTM; ++r) {
            const long long i = tile * kTM + r;
            const float h1v = sH1T[tid * kR + kTM + r];
            if (i < B) {
              C.DZ1[i * kH1 + tid] = act_bwd(acc[r], h1v, L.act);
              C.H1[i * kH1 + tid] = h1v;
            }
          }
          }
        }
        __syncthreads();
      }
      if (tid == 0 && !is_tgt) C.loss_part[cta] = loss_local;
      RMC_STAMP(C, 5);
    }
  } else if (do_rows && (S.phases & 2) && tid == 0 && cta < S.n_row_ctas) {
    C.loss_part[cta] = 0.f;
  }

  const int phaseB = S.phases & (4 | 8 | 16 | 32 | 64);
  if (!phaseB) return;
  const bool tree_here = pb.tree_here, team = pb.team, coarse = pb.coarse;
  const int n_workers = pb.n_workers, wid = cta;
  const bool tree_cta = tree_here && G > 1 && cta >= n_workers;
  if (stream_b && !tree_cta) {
    // ---------------------------------------------------------------- streamed phase B (no barrier: see StreamPlan)
    if (tid == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(C.barrier), "r"(1u) : "memory");   // the host counts G arrivals per launch
    RMC_STAMP(C, 6);
    {
      const int nt = static_cast<int>(n_tiles), s_idx = cta - sp.idle0;
      const int k = (s_idx >= 0) ? s_idx : (cta >= nt ? sp.n_idle + (cta - nt) : sp.n_idle + nt + cta);      // order in which CTAs become free
      const int total = kH2 / 16 + (kH1 / 16) * (kH2 / 16) + stream_w0_count(L), slots = sp.n_idle + 2 * nt;
      const int extra = max(0, total - slots);
      if (cta == nt && (S.phases & 2)) stream_publish_loss(C, S, nt, smem);      // first target CTA: its unit's operands are still 2 us away
      // More units than CTAs (152 units on 140 CTAs at every default batch): `nw` = extra rounded up to a multiple of four of
      // the CTAs that are idle in phase A (free from the start, 2 us before anybody else) take TWO neighbouring W2 units as one
      // 16 x 32 unit (kt = index / 4, columns 32 * (index % 4)); that removes 2 * nw units from the list, and what is left fits
      // the other CTAs one to one in the order they become free -- remaining idle CTAs, target CTAs, row CTAs -- and the order
      // the operands come into existence -- heads, W2, W0.  Nobody runs two units back to back (before, `extra` idle CTAs ran
      // two 16 x 16 units in sequence and ended the launch 2.3 us after the row CTAs).
      constexpr int nH = kH2 / 16, nJT = kH2 / 16;
      const int nw = (extra + nJT / 2 - 1) / (nJT / 2) * (nJT / 2);
      const bool wide_ok = extra > 0 && nw <= sp.n_idle && 2 * nw <= (kH1 / 16) * nJT;
      if (wide_ok) {
        if (s_idx >= 0 && s_idx < nw) {
          stream_run_w2_wide(C, S, s_idx / (nJT / 2), 2 * (s_idx % (nJT / 2)), smem);
        } else {
          const int rest_idle = sp.n_idle - nw;
          const int j = (s_idx >= 0) ? s_idx - nw : (cta >= nt) ? rest_idle + (cta - nt) : rest_idle + nt + cta;
          const int id = (j < nH) ? j : j + 2 * nw;                  // the list without the W2 units of the wide CTAs
          if (id < total) stream_run_any(C, S, id, smem);
        }
      } else {
        if (k < extra) stream_run_any(C, S, k, smem);
        if (extra + k < total) stream_run_any(C, S, extra + k, smem);
      }
    }
    RMC_STAMP(C, 7); if (gap != nullptr && threadIdx.x == 0) atomicMax(gap + 1, global_timer_ns());
    return;
  }
  if (do_rows) {
    if (early_td && tree_cta) {
      __shared__ float s_lp[kThreads];
      __syncthreads();
      if (tid == 0) red_release_add_u32(C.barrier, 1u);                  // arrive, do not wait
      SpinGuard guard;
      bool gave_up = false;
      for (long long i = tid; i < B && !gave_up; i += kThreads)
        while (ld_relaxed_pair(C.qt_flag + kTdWordBase + 2 * i).x != S.epoch) {
          __nanosleep(20);
          if (guard.expired()) { spin_report_timeout(C.host_loss, S.epoch); gave_up = true; break; }
        }
      float lp = 0.f;
      // (streamed phase B: the first TARGET CTA publishes the loss -- it is free long before its gradient unit's operands
      // exist, whereas here the system-scope fence of the publication would delay this member's share of the write-back)
      const bool loss_here = !stream_b && cta == n_workers;
      if (loss_here && tid < n_tiles) {
        uint2 w = ld_relaxed_pair(C.qt_flag + kLossWordBase + 2 * tid);
        while (w.x != S.epoch && !gave_up) {
          __nanosleep(20);
          w = ld_relaxed_pair(C.qt_flag + kLossWordBase + 2 * tid);
          if (guard.expired()) { spin_report_timeout(C.host_loss, S.epoch); gave_up = true; }
        }
        lp = __uint_as_float(w.y);
      }
      __threadfence();                                                   // acquire side of the flag words
      s_lp[tid] = lp;
      __syncthreads();
      if (loss_here && tid == 0) publish_loss(C, S, s_lp, static_cast<int>(n_tiles));
    } else {
      agent_barrier(C.barrier, S.barrier_target, C.host_loss, S.epoch);
    }
  }
  RMC_STAMP(C, 6);

  // ---------------------------------------------------------------- phase B
  if (tree_cta) {
    const long long tsize = C.rp.st->size;
    // |td| per row: the published {epoch, |td|} words when the write-back started early, else the row CTAs' array
    const float* td_src = early_td ? reinterpret_cast<const float*>(C.qt_flag + kTdWordBase) + 1 : C.abs_td;
    const int td_stride = early_td ? 2 : 1;
    unsigned long long* tdbg = C.dbg ? C.dbg + (blockIdx.y * gridDim.x + blockIdx.x) * kDbgSlots : nullptr;
    if (team) {
      tree_update_team(C.rp, C.nodes, td_src, td_stride, C.pri, B, tsize, (S.phases & 1) ? C.leaf_p : nullptr, S.per_eps, S.per_alpha,
                       S.per_pmax, cta - n_workers, (S.phases & 1) != 0, reinterpret_cast<double*>(smem), tdbg);
    } else {
      // |td| -> priority for the whole batch here (off the row CTAs' critical path), then the write-back
      for (long long i = tid; i < B; i += kThreads) C.pri[i] = td_to_priority(__ldcg(td_src + i * td_stride), S.per_eps, S.per_alpha, S.per_pmax);
      __syncthreads();
      RMC_STAMP(C, 8);
      tree_update_cta(C.rp, C.nodes, C.pri, B, tsize, tsize, false, (S.phases & 1) ? C.leaf_p : nullptr,
                      reinterpret_cast<double*>(smem), tdbg);
    }
    RMC_STAMP(C, 7); if (gap != nullptr && threadIdx.x == 0) atomicMax(gap + 1, global_timer_ns());
    return;
  }
  if (cta == 0 && (S.phases & 2) && tid == 0 && !early_td) {   // loss = (1/B) sum of the per-CTA partials, fixed order
    const int np = static_cast<int>(min(static_cast<long long>(S.n_row_ctas), n_tiles));
    publish_loss(C, S, C.loss_part, np);
  }
  if ((S.phases & 8) && use_ws) {
    // BACKWARD of a launch whose row CTAs left per-CTA partial gradient blobs (rmc_rows_ws.cuh): sum them, Adam / Polyak
    const int wrs = ws_rows(B, S.n_row_ctas);
    const int n_parts = static_cast<int>(min(static_cast<long long>(S.n_row_ctas), (B + wrs - 1) / wrs));
    ws_reduce_apply(C, S, C.gpart, n_parts, wid, n_workers);
  } else if (S.phases & 8) {                      // BACKWARD (+ fused Adam/Polyak)
    const int n_units = wgrad_unit_count(L, coarse);
    for (int u = wid; u < n_units; u += n_workers) wgrad_run_unit(C, S, u, coarse, smem);
  } else if (S.phases & (16 | 32 | 64)) {         // element-wise Adam from given grads / target sync only
    const float* gsrc = (S.grads_in != nullptr) ? S.grads_in + static_cast<size_t>(agent) * L.total : C.grads;
    for (int pi = wid * kThreads + tid; pi < L.total; pi += n_workers * kThreads)
      adam_polyak_element(C, S, pi, (S.phases & 16) ? __ldcg(gsrc + pi) : 0.f);
  }
  if (tree_here && G == 1) {
    __syncthreads();
    const long long tsize = C.rp.st->size;
    for (long long i = tid; i < B; i += kThreads) C.pri[i] = td_to_priority(__ldcg(C.abs_td + i), S.per_eps, S.per_alpha, S.per_pmax);
    __syncthreads();
    tree_update_cta(C.rp, C.nodes, C.pri, B, tsize, tsize, false, (S.phases & 1) ? C.leaf_p : nullptr,
                    reinterpret_cast<double*>(smem), nullptr);
  }
  RMC_STAMP(C, 7); if (gap != nullptr && threadIdx.x == 0) atomicMax(gap + 1, global_timer_ns());
}

// ------------------------------------------------------------------ per-env-step push of an ensemble
// Every member of an ensemble launch stores the n rows its own environments produced (Agent.store_transitions,
// dqn/agent.py:70-73 -> replay_memory.py:49-57): ONE launch, block b = member b, the rows in the kernel-argument buffer.
// (One k_push_tiny per member was 8 launches and 8 serialised one-warp kernels in front of every ensemble step.)
constexpr int kGroupRowFloats = 896;      // 3.5 KB of packed rows per launch, member-major
struct GroupRows { float v[kGroupRowFloats]; };
__global__ void __launch_bounds__(32) k_push_tiny_group(const AgentCtx* __restrict__ many, const __grid_constant__ GroupRows rows, int n, int rf, float pmax) {
  __shared__ float s_f[64];
  __shared__ int s_i[64];
  const ReplayDev R = many[blockIdx.x].rp;
  push_tiny_cta(R, rows.v + static_cast<size_t>(blockIdx.x) * n * rf, n, pmax, s_f, s_i);
}

// ------------------------------------------------------------------ batched act / Q values
// mode 0: greedy actions (dueling -> argmax raw adv, plain -> argmax Q; dqn/network.py:67-74,110-117)
// mode 1: Q values [n][A]
// mode 2: raw head outputs [n][NH]
__global__ void __launch_bounds__(kThreads, 1) k_mlp_infer(NetLayout L, const float* __restrict__ params, const float* __restrict__ obs,
                                                           long long n, long long* __restrict__ actions, float* __restrict__ q_out, int mode) {
  extern __shared__ __align__(16) float smem[];
  const SmemPlan P = make_smem_plan(L.total);
  float* sW = smem + P.w;
  float* sXT = smem + P.xt;
  float* sH1T = smem + P.h1t;
  float* sH2 = smem + P.h2;
  float* sPart = smem + P.part;
  float* sQ = smem + P.q;
  float* sRaw = smem + P.dz2;   // [kR][kQLD] fits in the dz2 area
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P.bar);
  const int tid = threadIdx.x;
  const long long n_tiles = (n + kR - 1) / kR;
  if (blockIdx.x >= n_tiles) return;
  uint32_t parity = 0;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  stage_params(sW, params, L.total, bar, parity);
  wait_params(bar, parity);
  const int D = L.D;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int t = tid; t < kR * D; t += kThreads) {
      const int r = t / D, d = t % D;
      const long long i = tile * kR + r;
      sXT[d * kR + r] = (i < n) ? __ldg(obs + i * D + d) : 0.f;
    }
    __syncthreads();
    mlp_forward<kR>(sW, L, sXT, sH1T, sH2, sPart, sQ, sRaw);
    if (mode == 0) {
      if (tid < kR) {
        const long long i = tile * kR + tid;
        if (i < n) actions[i] = L.dueling ? argmax_first(sRaw + tid * kQLD + 1, L.A) : argmax_first(sQ + tid * kQLD, L.A);
      }
    } else if (mode == 1) {
      for (int t = tid; t < kR * L.A; t += kThreads) {
        const int r = t / L.A, a = t % L.A;
        const long long i = tile * kR + r;
        if (i < n) q_out[i * L.A + a] = sQ[r * kQLD + a];
      }
    } else {   // mode 2: raw head outputs [n][NH] (dueling: value then advantages; network.py:98-108)
      for (int t = tid; t < kR * L.NH; t += kThreads) {
        const int r = t / L.NH, a = t % L.NH;
        const long long i = tile * kR + r;
        if (i < n) q_out[i * L.NH + a] = sRaw[r * kQLD + a];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ per-env-step act (n <= kActTinyMax host states)
// The call the trainer makes once per environment step (train.py:91-93 -> Agent.choose_actions) with n_env state vectors.
// Both copies and the stream synchronisation of the general path are gone: the states ride in the kernel-argument buffer,
// the actions are stored straight into mapped pinned host memory, and the launch's epoch follows them (system fence in
// between), so the host waits on one word.  Same arithmetic as k_mlp_infer mode 0 and the same Philox epsilon-greedy draw
// as k_eps_greedy (eps < 0: greedy only).  One CTA per 8 rows; the last CTA to finish publishes the epoch.
constexpr int kActTinyMax = 32;
struct ActTinyObs { float v[kActTinyMax * 16]; };     // [n][D] packed, D <= 16
__device__ __forceinline__ long long eps_greedy_pick(long long greedy, long long i, float eps, int n_actions, unsigned long long seed,
                                                     unsigned long long counter) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32)),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) ^ 0x9E3779B9u));
  const float u1 = static_cast<float>(r.x >> 8) * (1.0f / 16777216.0f);
  return (u1 <= eps) ? static_cast<long long>((static_cast<unsigned long long>(r.y) * static_cast<unsigned long long>(n_actions)) >> 32) : greedy;
}
__global__ void __launch_bounds__(kThreads, 1) k_act_tiny(NetLayout L, const float* __restrict__ params, const __grid_constant__ ActTinyObs X, int n,
                                                          float eps, unsigned long long seed, unsigned long long counter,
                                                          volatile long long* host_actions, volatile unsigned* host_epoch, unsigned* ctr, unsigned epoch) {
  extern __shared__ __align__(16) float smem[];
  const SmemPlan P = make_smem_plan(L.total);
  float* sW = smem + P.w;
  float* sXT = smem + P.xt;
  float* sRaw = smem + P.dz2;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P.bar);
  const int tid = threadIdx.x, D = L.D;
  uint32_t parity = 0;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  stage_params(sW, params, L.total, bar, parity);
  const int row0 = blockIdx.x * kR;
  for (int t = tid; t < kR * D; t += kThreads) {
    const int r = t / D, d = t - r * D;
    sXT[d * kR + r] = (row0 + r < n) ? X.v[(row0 + r) * D + d] : 0.f;
  }
  wait_params(bar, parity);
  __syncthreads();
  mlp_forward<kR>(sW, L, sXT, smem + P.h1t, smem + P.h2, smem + P.part, smem + P.q, sRaw);
  if (tid < kR && row0 + tid < n) {
    const long long i = row0 + tid;
    long long a = L.dueling ? argmax_first(sRaw + tid * kQLD + 1, L.A) : argmax_first(smem + P.q + tid * kQLD, L.A);
    if (eps >= 0.f) a = eps_greedy_pick(a, i, eps, L.A, seed, counter);
    host_actions[i] = a;
    __threadfence_system();
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(ctr, 1u);
    if (prev == gridDim.x - 1) {
      *ctr = 0u;
      __threadfence_system();
      *host_epoch = epoch;
    }
  }
}

// epsilon-greedy on the device (Agent.choose_actions, dqn/agent.py:92-99): row i explores when u1 <= eps and then takes
// floor(u2 * A); (u1, u2) = Philox4x32-10(seed, counter, i).  The reference draws from Python's Mersenne Twister, so this
// mode reproduces its distribution, not its stream (the host-RNG mode of Agent.choose_actions reproduces the stream).
__global__ void k_eps_greedy(long long* __restrict__ actions, long long n, float eps, int n_actions, unsigned long long seed, unsigned long long counter) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32)),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32) ^ 0x9E3779B9u));
  const float u1 = static_cast<float>(r.x >> 8) * (1.0f / 16777216.0f);
  if (u1 <= eps) actions[i] = static_cast<long long>((static_cast<unsigned long long>(r.y) * static_cast<unsigned long long>(n_actions)) >> 32);
}

// parameter (de)interleave between torch state_dict order and the device layout
__global__ void k_params_scatter(float* dev_blob, const float* src_torch, const int* map, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) dev_blob[map[i]] = src_torch[i];
}
__global__ void k_params_gather(float* dst_torch, const float* dev_blob, const int* map, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) dst_torch[i] = dev_blob[map[i]];
}

}  // namespace rmc
