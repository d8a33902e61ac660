// rmc_rows_ws.cuh -- row phase AND weight gradients of the fused learner step for launches in which a row CTA owns at
// least two 16-row tiles: batches of thousands of rows (BASELINE configs[4], on one GPU and per rank of the sharded step).
// Included by rmc_mlp.cuh; used by k_learner_step<2> (the host picks the instantiation: step_path in rmc_b200.cu).  The
// 8-agent ensemble launch (16 rows per CTA) was measured on this path and stays on the 4-row tiles -- see DESIGN.md 3.1b.
//
// Why a second form of the row phase.  The single-tile path (k_learner_step<1>) is shaped by latency: 4 rows per CTA,
// both operands of every product read from shared memory.  Per FMA that is 1.5-3 bytes of shared-memory traffic, and the
// SM delivers 128 B/clk against 128 FMA/clk -- so once a CTA owns tens or hundreds of rows the passes run at the
// shared-memory roofline, a quarter to a third of the FMA rate (measured: a 16-row layer-2 pass 8.6 us against a 2.2 us FMA
// floor; the 16x16 gradient units 13-18 % of the FMA peak at B = 65,536).
//
// Here one operand of every large product lives in REGISTERS for a whole phase and the batch rows stream past it:
//   phase T  target pass      W2_target slice (16 k x 8 columns = 128 registers per thread); rows -> Q_target(s')
//   phase F  online pass      W2_online slice; rows -> Q(s'), Q(s), TD target, Huber, head deltas, dz2; head / b2 / bh
//                             gradients accumulate in registers
//   phase D  dgrad            W2^T slice (8 k x 16 columns); dz2 rows -> dz1; thread t ends up owning hidden unit t, so the
//                             W0 / b0 gradients accumulate in registers right there
//   phase G  wgrad            128 accumulators of dW2 (8 k x 16 columns) per thread; H1 / dz2 rows stream past
// Shared-memory traffic per FMA drops to 0.5-0.75 B and the K reductions of T / F / D happen inside a warp with a
// transposing butterfly (8 or 7 shuffles per row and lane instead of 32).  Every CTA leaves ONE partial gradient blob for
// its rows; after the agent barrier the workers sum the partials in CTA order and apply Adam / Polyak (same element
// arithmetic as everywhere else: param_apply).  Everything is fixed-order, hence deterministic.
//
// Arithmetic is that of dqn/agent.py:245-272 / Appendix A of SURVEY.md as in rmc_mlp.cuh; only summation orders differ from
// the single-tile path (K sums in 16 interleaved partial sums; batch sums per CTA, then over CTAs) -- covered by the 1e-5
// parity bar (tests/test_gpu_headline_sizes.py, tests/test_gpu_parallel.py) and by the float64 ground-truth check.
#pragma once
// (included inside namespace rmc of rmc_mlp.cuh, after its Adam / cp.async helpers)

constexpr int kWR = 16;                                              // rows per tile (a launch may use 8 of them, see ws_rows)
constexpr int kWsTileFloats = kWR * kH1 + kWR * kH2 + kWR * kMaxD;   // staged tile of phases D / G: H1 | dz2 | x

struct WsSmem { float *xt, *h1, *h2, *dz2, *q, *qn, *dh, *raw, *qt, *meta, *red; };
__device__ __forceinline__ WsSmem ws_smem(float* base) {            // aliases the [xt .. part] region of SmemPlan (11,520 floats)
  WsSmem w;
  float* p = base;
  w.xt = p; p += kMaxD * kWR;          // [kMaxD][16]   x transposed (layer 1)
  w.h1 = p; p += kWR * kH1;            // [16][256]
  w.h2 = p; p += kWR * kH2;            // [16][128]
  w.dz2 = p; p += kWR * kH2;           // [16][128]
  w.q = p; p += kWR * kQLD;            // [16][16]  Q_online(s)
  w.qn = p; p += kWR * kQLD;           // [16][16]  Q_online(s')
  w.dh = p; p += kWR * kQLD;           // [16][16]
  w.raw = p; p += kWR * kQLD;          // [16][16]  raw head outputs of the pass
  w.qt = p; p += kWR * kQLD;           // [16][16]  Q_target(s') of the tile
  w.meta = p; p += kWR * 4;            // [16][4] action bits, reward, done, is_w
  w.red = p; p += 2 * kWR;
  return w;
}
static_assert(kMaxD * kWR + kWR * kH1 + 2 * kWR * kH2 + 5 * kWR * kQLD + kWR * 4 + 2 * kWR <= kMaxD * kR + kH1 * kR + kR * kH2 + kWarps * kR * kH2,
              "batch-stationary row phase must fit the xt..part region");

// rows per tile of this launch: 16, or 8 when 16-row tiles would leave row CTAs without work (mid-size batches)
__device__ __forceinline__ int ws_rows(long long B, int n_row_ctas) { return ((B + kWR - 1) / kWR >= n_row_ctas) ? kWR : kWR / 2; }

// ---- layer-2 weight slice of the forward passes: k in {64q + 4ks + e}, columns 16*warp + 8*cg + c   (ks = lane & 15, cg = lane >> 4)
__device__ __forceinline__ void ws_load_w2_fwd(const float* __restrict__ sW, const NetLayout& L, float (&w)[16][8]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ks = lane & 15, cg = lane >> 4;
  const float* base = sW + L.off_w2t + 16 * warp + 8 * cg;
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float* row = base + (64 * q + 4 * ks + e) * kW2LD;
      const float4 a = *reinterpret_cast<const float4*>(row), b = *reinterpret_cast<const float4*>(row + 4);
      w[4 * q + e][0] = a.x; w[4 * q + e][1] = a.y; w[4 * q + e][2] = a.z; w[4 * q + e][3] = a.w;
      w[4 * q + e][4] = b.x; w[4 * q + e][5] = b.y; w[4 * q + e][6] = b.z; w[4 * q + e][7] = b.w;
    }
}
// 8 partial sums per lane, 16 lanes (the K slices of one column group) -> every lane keeps the complete sum of ONE column:
// exchange halves (xor 8, 4, 2), then a plain xor-1 add.  Returns the sum of column (ks >> 1) & 7 of the lane's group.
__device__ __forceinline__ float ws_reduce16(const float (&a)[8], int ks) {
  float t[4], u[2];
  const bool up3 = (ks & 8) != 0, up2 = (ks & 4) != 0, up1 = (ks & 2) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = up3 ? a[i] : a[i + 4], keep = up3 ? a[i + 4] : a[i];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = up2 ? t[i] : t[i + 2], keep = up2 ? t[i + 2] : t[i];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = up1 ? u[0] : u[1], keep = up1 ? u[1] : u[0];
  float v = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
// 8 partial sums per lane, 8 lanes -> every lane keeps the complete sum of element (lane & 7)
__device__ __forceinline__ float ws_reduce8(const float (&a)[8], int js) {
  float t[4], u[2];
  const bool up2 = (js & 4) != 0, up1 = (js & 2) != 0, up0 = (js & 1) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = up2 ? a[i] : a[i + 4], keep = up2 ? a[i + 4] : a[i];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = up1 ? t[i] : t[i + 2], keep = up1 ? t[i + 2] : t[i];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float send = up0 ? u[0] : u[1], keep = up0 ? u[1] : u[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// Input prefetch: the x columns of the NEXT pass are fetched into registers while the current pass computes (a pass that
// started with a dependent L2 round trip spent a tenth of its time waiting for 1 KB of inputs).
struct WsXPre { float v[2]; };
__device__ __forceinline__ WsXPre ws_x_fetch(const AgentCtx& C, long long row0, long long B, int col0, int wr) {
  const int D = C.L.D, rf = C.rp.row_floats;
  WsXPre x;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int t = threadIdx.x + q * kThreads, r = t / D, d = t % D;
    const long long i = row0 + r;
    x.v[q] = (t < kWR * D && r < wr && i < B) ? __ldcg(C.X + i * rf + col0 + d) : 0.f;
  }
  return x;
}
__device__ __forceinline__ void ws_x_store(const AgentCtx& C, float* __restrict__ sXT, const WsXPre& x) {     // -> sXT[d][16]
  const int D = C.L.D;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int t = threadIdx.x + q * kThreads, r = t / D, d = t % D;
    if (t < kWR * D) sXT[d * kWR + r] = x.v[q];
  }
}

// forward of the tile's rows: sXT -> sH1 [r][256], sH2 [r][128], sQ [r][16] (Q values).  w2 = this thread's layer-2 slice.
__device__ __forceinline__ void ws_forward(const float* __restrict__ sW, const NetLayout& L, const float (&w2)[16][8], const float* __restrict__ sXT,
                                           float* __restrict__ sH1, float* __restrict__ sH2, float* __restrict__ sRaw, float* __restrict__ sQ, int wr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {   // layer 1: thread = 4 hidden units x 4 rows (bias first, then d ascending: the order of mlp_forward)
    const int ug = tid & 63, rg = tid >> 6;
    if (4 * rg < wr) {
      const float4 b = *reinterpret_cast<const float4*>(sW + L.off_b0 + 4 * ug);
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { acc[r][0] = b.x; acc[r][1] = b.y; acc[r][2] = b.z; acc[r][3] = b.w; }
      const float* w0 = sW + L.off_w0t + 4 * ug;
      const float* xp = sXT + 4 * rg;
#pragma unroll 2
      for (int d = 0; d < L.D; ++d) {
        const float4 w = *reinterpret_cast<const float4*>(w0 + d * kH1);
        const float4 x = *reinterpret_cast<const float4*>(xp + d * kWR);
        const float xr[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][0] = fmaf(xr[r], w.x, acc[r][0]); acc[r][1] = fmaf(xr[r], w.y, acc[r][1]);
          acc[r][2] = fmaf(xr[r], w.z, acc[r][2]); acc[r][3] = fmaf(xr[r], w.w, acc[r][3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float4*>(sH1 + (4 * rg + r) * kH1 + 4 * ug) =
            make_float4(act_fwd(acc[r][0], L.act), act_fwd(acc[r][1], L.act), act_fwd(acc[r][2], L.act), act_fwd(acc[r][3], L.act));
    }
  }
  __syncthreads();
  {   // layer 2: the weights stay in registers, two rows per step
    const int ks = lane & 15, cg = lane >> 4;
    const int col = 16 * warp + 8 * cg + ((ks >> 1) & 7);
    const float bias = sW[L.off_b2 + col];
#pragma unroll 2
    for (int r0 = 0; r0 < wr; r0 += 2) {
      float a0[8], a1[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
      const float* h = sH1 + r0 * kH1 + 4 * ks;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 x0 = *reinterpret_cast<const float4*>(h + 64 * q), x1 = *reinterpret_cast<const float4*>(h + kH1 + 64 * q);
        const float h0[4] = {x0.x, x0.y, x0.z, x0.w}, h1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            a0[c] = fmaf(h0[e], w2[4 * q + e][c], a0[c]);
            a1[c] = fmaf(h1[e], w2[4 * q + e][c], a1[c]);
          }
      }
      const float s0 = ws_reduce16(a0, ks) + bias, s1 = ws_reduce16(a1, ks) + bias;
      if ((ks & 1) == 0) {
        sH2[r0 * kH2 + col] = act_fwd(s0, L.act);
        sH2[(r0 + 1) * kH2 + col] = act_fwd(s1, L.act);
      }
    }
  }
  __syncthreads();
  {   // heads: 16 lanes per row (lane slice s: columns 4s + 64q), 8 heads per round; the 16 partial sums of a head meet in the
      // same transposing butterfly as layer 2.  Rows 2*warp, 2*warp + 1 belong to this warp.
    const int r = tid >> 4, sl = tid & 15;
    const float4 ha = *reinterpret_cast<const float4*>(sH2 + r * kH2 + 4 * sl), hb = *reinterpret_cast<const float4*>(sH2 + r * kH2 + 4 * sl + 64);
#pragma unroll 1
    for (int g8 = 0; g8 < L.NH; g8 += 8) {
      float part[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const float* wh = sW + L.off_wh + min(g8 + a, L.NH - 1) * kH2 + 4 * sl;
        const float4 wa = *reinterpret_cast<const float4*>(wh), wb = *reinterpret_cast<const float4*>(wh + 64);
        float p = ha.x * wa.x;
        p = fmaf(ha.y, wa.y, p); p = fmaf(ha.z, wa.z, p); p = fmaf(ha.w, wa.w, p);
        p = fmaf(hb.x, wb.x, p); p = fmaf(hb.y, wb.y, p); p = fmaf(hb.z, wb.z, p); p = fmaf(hb.w, wb.w, p);
        part[a] = p;
      }
      const float tot = ws_reduce16(part, sl);
      const int a = g8 + ((sl >> 1) & 7);
      if ((sl & 1) == 0 && a < L.NH && r < wr) sRaw[r * kQLD + a] = tot + sW[L.off_bh + a];
    }
    __syncwarp();
    if (r < wr) {
      float qv = 0.f;
      if (sl < L.A) {
        if (L.dueling) {   // Q = val + (adv - mean(adv))   (dqn/network.py:83), adv summed in action order
          float sum = 0.f;
          for (int a = 1; a <= L.A; ++a) sum += sRaw[r * kQLD + a];
          qv = sRaw[r * kQLD] + (sRaw[r * kQLD + 1 + sl] - sum / static_cast<float>(L.A));
        } else {
          qv = sRaw[r * kQLD + sl];
        }
      }
      sQ[r * kQLD + sl] = qv;
    }
  }
  __syncthreads();
}

// this CTA's tiles: cta, cta + n_row_ctas, ... < n_wide
struct WsTiles { long long n_wide; int wr, stride, first; };
__device__ __forceinline__ WsTiles ws_tiles(const StepScalars& S, int cta) {
  WsTiles t;
  t.wr = ws_rows(S.B, S.n_row_ctas);
  t.n_wide = (S.B + t.wr - 1) / t.wr;
  t.stride = S.n_row_ctas;
  t.first = cta;
  return t;
}

// ---- phases T and F: the forward passes over this CTA's rows.  One function (and ONE copy of the forward code: a CTA of an
// ensemble launch runs every phase once, so what it pays for is cold instruction fetch) serves both:
//   target = true   sW holds the target blob: Q_target(s') of every row -> C.QT
//   target = false  sW holds the online blob: Q(s'), Q(s), TD target, Huber, head deltas, dz2, activations -> scratch; the
//                   head / b2 / bh gradient partials of the rows -> gpart (when set).  Returns the CTA's loss partial (thread 0).
__device__ __forceinline__ float ws_rows_phase(const AgentCtx& C, const StepScalars& S, const float* __restrict__ sW, const WsSmem& W, const WsTiles& T,
                                            bool per, float* __restrict__ gpart, bool target) {
  const NetLayout& L = C.L;
  const int tid = threadIdx.x, D = L.D, rf = C.rp.row_floats, wr = T.wr;
  const long long B = S.B;
  WsXPre xpre = ws_x_fetch(C, T.first * wr, B, D, wr);                 // s' rows of the first tile: in flight during the weight load
  float w2[16][8];
  ws_load_w2_fwd(sW, L, w2);
  // gradient partials that need nothing but this phase's shared-memory tiles: thread (j = tid & 127, half = tid >> 7) owns
  // dWh[a][j] for the rows of its half, db2[j] likewise; threads 0..15 own dbh[a]
  const int gj = tid & (kH2 - 1), ghalf = tid >> 7;
  float gwh[kQLD], gb2 = 0.f, gbh = 0.f;
#pragma unroll
  for (int a = 0; a < kQLD; ++a) gwh[a] = 0.f;
  float loss_local = 0.f;
#pragma unroll 1
  for (long long wt = T.first; wt < T.n_wide; wt += T.stride) {
    const long long row0 = wt * wr;
    float mpre = 0.f, qtpre = 0.f;
    if (!target) {     // this tile's transition fields and Q_target rows: in flight during the two forward passes
      const int r = tid >> 2, f = tid & 3;
      const long long i = row0 + r;
      if (tid < 4 * kWR && r < wr && i < B) mpre = (f < 3) ? __ldcg(C.X + i * rf + 2 * D + f) : __ldcg(C.is_w + i);
      const long long iq = row0 + tid / kQLD;
      if (tid / kQLD < wr && iq < B) qtpre = __ldcg(C.QT + iq * kQLD + (tid % kQLD));
    }
#pragma unroll 1
    for (int pass = 0; pass < (target ? 1 : 2); ++pass) {
      ws_x_store(C, W.xt, xpre);
      __syncthreads();
      // next pass's inputs: the s rows of this tile, or the s' rows of the next tile
      if (!target && pass == 0) xpre = ws_x_fetch(C, row0, B, 0, wr);
      else if (wt + T.stride < T.n_wide) xpre = ws_x_fetch(C, (wt + T.stride) * wr, B, D, wr);
      ws_forward(sW, L, w2, W.xt, W.h1, W.h2, W.raw, (target || pass == 1) ? W.q : W.qn, wr);
    }
    if (target) {
      const int r = tid / kQLD;
      const long long i = row0 + r;
      if (r < wr && i < B) C.QT[i * kQLD + (tid % kQLD)] = W.q[tid];
      __syncthreads();
      continue;
    }
    if (tid < 4 * kWR) W.meta[tid] = mpre;
    W.qt[tid] = qtpre;
    __syncthreads();
    // ---- TD target, |td|, Huber, dQ coefficient (threads 0..15), same arithmetic as the single-tile path
    if (tid < kWR) {
      const int r = tid;
      const long long i = row0 + r;
      float g = 0.f, lterm = 0.f;
      int act = 0;
      if (r < wr && i < B) {
        const float* qtv = W.qt + r * kQLD;
        float qsel;
        if (S.double_dqn) {
          qsel = qtv[argmax_first(W.qn + r * kQLD, L.A)];
        } else {
          qsel = qtv[0];
          for (int a = 1; a < L.A; ++a) qsel = fmaxf(qsel, qtv[a]);
        }
        act = __float_as_int(W.meta[r * 4 + 0]);
        const float rew = W.meta[r * 4 + 1], done = W.meta[r * 4 + 2], w = W.meta[r * 4 + 3];
        const float y = rew + ((1.f - done) * S.gamma) * qsel;
        const float q_sa = W.q[r * kQLD + act];
        const float delta = q_sa - y;
        const float atd = fabsf(y - q_sa);
        const float z = fabsf(delta);
        const float hub = (z < 1.f) ? (0.5f * z) * z : z - 0.5f;
        const float go = per ? (1.f / static_cast<float>(S.Bglobal)) * w : 1.f / static_cast<float>(S.Bglobal);
        g = fminf(fmaxf(delta, -1.f), 1.f) * go;
        lterm = per ? w * hub : hub;
        C.y[i] = y; C.q_sa[i] = q_sa; C.abs_td[i] = atd; C.hub[i] = hub; C.gcoef[i] = g;
      }
      float* dh = W.dh + r * kQLD;
      for (int a = 0; a < kQLD; ++a) dh[a] = 0.f;
      if (L.dueling) {
        const float mean = g / static_cast<float>(L.A);
        dh[0] = g;
        for (int a = 0; a < L.A; ++a) dh[1 + a] = ((a == act) ? g : 0.f) - mean;
      } else {
        dh[act] = g;
      }
      W.red[r] = lterm;
    }
    {   // parity outputs of the Q rows
      const int r = tid / kQLD;
      const long long i = row0 + r;
      if (r < wr && i < B) { C.QN[i * kQLD + (tid % kQLD)] = W.qn[tid]; C.Q[i * kQLD + (tid % kQLD)] = W.q[tid]; }
    }
    __syncthreads();
    if (tid == 0)
      for (int r = 0; r < wr; ++r) loss_local += W.red[r];
    // ---- dh2 -> dz2 (thread: column j, rows half, half+2, ...), with the head / b2 gradients of those rows
    {
      float whj[kQLD];
#pragma unroll
      for (int a = 0; a < kQLD; ++a) whj[a] = (a < L.NH) ? sW[L.off_wh + a * kH2 + gj] : 0.f;
#pragma unroll 2
      for (int r = ghalf; r < wr; r += 2) {
        const float* dh = W.dh + r * kQLD;
        const float4 d0 = *reinterpret_cast<const float4*>(dh), d1 = *reinterpret_cast<const float4*>(dh + 4),
                     d2 = *reinterpret_cast<const float4*>(dh + 8), d3 = *reinterpret_cast<const float4*>(dh + 12);
        const float dv[kQLD] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w, d3.x, d3.y, d3.z, d3.w};
        float sum = 0.f;
#pragma unroll
        for (int a = 0; a < kQLD; ++a) sum = fmaf(dv[a], whj[a], sum);      // heads past NH: dv = 0, whj = 0 (adds +0)
        const float h2v = W.h2[r * kH2 + gj];
        const float dz = act_bwd(sum, h2v, L.act);
        W.dz2[r * kH2 + gj] = dz;
        gb2 += dz;
#pragma unroll
        for (int a = 0; a < kQLD; ++a) gwh[a] = fmaf(h2v, dv[a], gwh[a]);
      }
    }
    if (tid < kQLD)
      for (int r = 0; r < wr; ++r) gbh += W.dh[r * kQLD + tid];
    __syncthreads();
    // ---- activations / deltas of the s rows -> L2-resident scratch (phases D / G of this CTA, parity outputs)
    for (int t = tid; t < wr * (kH1 / 4); t += kThreads) {
      const int r = t / (kH1 / 4), c = t % (kH1 / 4);
      const long long i = row0 + r;
      if (i < B) *reinterpret_cast<float4*>(C.H1 + i * kH1 + 4 * c) = *reinterpret_cast<const float4*>(W.h1 + r * kH1 + 4 * c);
    }
    for (int t = tid; t < wr * (kH2 / 4); t += kThreads) {
      const int r = t / (kH2 / 4), c = t % (kH2 / 4);
      const long long i = row0 + r;
      if (i < B) {
        *reinterpret_cast<float4*>(C.H2 + i * kH2 + 4 * c) = *reinterpret_cast<const float4*>(W.h2 + r * kH2 + 4 * c);
        *reinterpret_cast<float4*>(C.DZ2 + i * kH2 + 4 * c) = *reinterpret_cast<const float4*>(W.dz2 + r * kH2 + 4 * c);
      }
    }
    {
      const int r = tid / kQLD;
      const long long i = row0 + r;
      if (r < wr && i < B) C.DH[i * kQLD + (tid % kQLD)] = W.dh[tid];
    }
    __syncthreads();
  }
  if (!target && gpart != nullptr) {
    // combine the two row halves in shared memory, then one store per element
    float* sc = W.h2;                         // [kQLD][128] of half 1 (the h1 / dz2 tiles stay: a CTA with a single tile reuses them in phases D / G)
    float* sb = W.q;                          // [128]
    if (ghalf == 1) {
#pragma unroll
      for (int a = 0; a < kQLD; ++a) sc[a * kH2 + gj] = gwh[a];
      sb[gj] = gb2;
    }
    __syncthreads();
    if (ghalf == 0) {
#pragma unroll
      for (int a = 0; a < kQLD; ++a)
        if (a < L.NH) gpart[L.off_wh + a * kH2 + gj] = gwh[a] + sc[a * kH2 + gj];
      gpart[L.off_b2 + gj] = gb2 + sb[gj];
    }
    if (tid < L.NH) gpart[L.off_bh + tid] = gbh;
    __syncthreads();
  }
  return loss_local;
}

// ---- staging of one tile of the s rows from the scratch for phases D / G: H1 | dz2 | x (rows past the batch: zeros)
__device__ __forceinline__ void ws_stage_tile(const AgentCtx& C, float* __restrict__ buf, long long row0, long long B, int wr, bool with_x, bool with_act = true) {
  float* bh1 = buf;
  float* bdz = buf + kWR * kH1;
  float* bx = bdz + kWR * kH2;
  const int tid = threadIdx.x;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = tid; with_act && t < wr * (kH1 / 4); t += kThreads) {
    const int r = t / (kH1 / 4), c = t % (kH1 / 4);
    const long long i = row0 + r;
    if (i < B) cp_async16(bh1 + r * kH1 + 4 * c, C.H1 + i * kH1 + 4 * c);
    else *reinterpret_cast<float4*>(bh1 + r * kH1 + 4 * c) = z4;
  }
  for (int t = tid; with_act && t < wr * (kH2 / 4); t += kThreads) {
    const int r = t / (kH2 / 4), c = t % (kH2 / 4);
    const long long i = row0 + r;
    if (i < B) cp_async16(bdz + r * kH2 + 4 * c, C.DZ2 + i * kH2 + 4 * c);
    else *reinterpret_cast<float4*>(bdz + r * kH2 + 4 * c) = z4;
  }
  if (with_x) {      // the s columns [0, D) rounded up to whole 16-byte chunks (row stride is a multiple of 4 floats)
    const int xc = (C.L.D + 3) >> 2, rf = C.rp.row_floats;
    for (int t = tid; t < wr * xc; t += kThreads) {
      const int r = t / xc, c = t % xc;
      const long long i = row0 + r;
      if (i < B) cp_async16(bx + r * kMaxD + 4 * c, C.X + i * rf + 4 * c);
      else *reinterpret_cast<float4*>(bx + r * kMaxD + 4 * c) = z4;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// ---- phase D: dz1 = (dz2 . W2) (.) act'(h1) of this CTA's rows, with the W0 / b0 gradient partials (thread t <-> hidden unit t).
// `stage` = two tile buffers inside the parameter-blob area, which is dead once the W2^T slice sits in registers.
template <int ND>
__device__ __forceinline__ void ws_dgrad_phase(const AgentCtx& C, const StepScalars& S, const float* __restrict__ sW, float* __restrict__ stage, const WsSmem& W,
                                               const WsTiles& T, float* __restrict__ gpart) {
  const NetLayout& L = C.L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, js = lane & 7, kg = lane >> 3, wr = T.wr;
  const long long B = S.B;
  float w[8][16];     // W2^T[k][j]: k = 32*warp + 8*kg + e, j = 32q + 4js + f
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float* row = sW + L.off_w2t + (32 * warp + 8 * kg + e) * kW2LD + 4 * js;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(row + 32 * q);
      w[e][4 * q] = v.x; w[e][4 * q + 1] = v.y; w[e][4 * q + 2] = v.z; w[e][4 * q + 3] = v.w;
    }
  }
  __syncthreads();                 // every thread has its slice: the blob area may be overwritten
  float gw0[ND], gb0 = 0.f;
#pragma unroll
  for (int d = 0; d < ND; ++d) gw0[d] = 0.f;
  int it = 0;
  // a CTA with a single tile (ensemble launches) still holds that tile's h1 / dz2 in shared memory: only x is staged
  const bool single = T.first + T.stride >= T.n_wide;
  if (T.first < T.n_wide) ws_stage_tile(C, stage, T.first * wr, B, wr, true, !single);
  for (long long wt = T.first; wt < T.n_wide; wt += T.stride, ++it) {
    const float* cur = stage + (it & 1) * kWsTileFloats;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (wt + T.stride < T.n_wide) ws_stage_tile(C, stage + ((it + 1) & 1) * kWsTileFloats, (wt + T.stride) * wr, B, wr, true);
    const float* sH1 = single ? W.h1 : cur;
    const float* sDZ = single ? W.dz2 : cur + kWR * kH1;
    const float* sX = cur + kWR * kH1 + kWR * kH2;
    const long long row0 = wt * wr;
#pragma unroll 2
    for (int r0 = 0; r0 < wr; r0 += 2) {
      float a0[8], a1[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { a0[e] = 0.f; a1[e] = 0.f; }
      const float* dzp = sDZ + r0 * kH2 + 4 * js;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v0 = *reinterpret_cast<const float4*>(dzp + 32 * q), v1 = *reinterpret_cast<const float4*>(dzp + kH2 + 32 * q);
        const float z0[4] = {v0.x, v0.y, v0.z, v0.w}, z1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a0[e] = fmaf(z0[f], w[e][4 * q + f], a0[e]);
            a1[e] = fmaf(z1[f], w[e][4 * q + f], a1[e]);
          }
      }
      const float s0 = ws_reduce8(a0, js), s1 = ws_reduce8(a1, js);      // hidden unit 32*warp + 8*kg + js = tid
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = r0 + rr;
        const long long i = row0 + r;
        const float dz1 = act_bwd(rr ? s1 : s0, sH1[r * kH1 + tid], L.act);
        if (i < B) C.DZ1[i * kH1 + tid] = dz1;
        gb0 += dz1;
        const float* xp = sX + r * kMaxD;
#pragma unroll
        for (int d4 = 0; d4 < ND / 4; ++d4) {
          const float4 x = *reinterpret_cast<const float4*>(xp + 4 * d4);
          gw0[4 * d4] = fmaf(x.x, dz1, gw0[4 * d4]); gw0[4 * d4 + 1] = fmaf(x.y, dz1, gw0[4 * d4 + 1]);
          gw0[4 * d4 + 2] = fmaf(x.z, dz1, gw0[4 * d4 + 2]); gw0[4 * d4 + 3] = fmaf(x.w, dz1, gw0[4 * d4 + 3]);
        }
      }
    }
  }
  if (gpart != nullptr) {
#pragma unroll
    for (int d = 0; d < ND; ++d)
      if (d < L.D) gpart[L.off_w0t + d * kH1 + tid] = gw0[d];
    gpart[L.off_b0 + tid] = gb0;
  }
  __syncthreads();                 // the last tile's buffer may be restaged by phase G
}

// ---- phase G: dW2^T[k][j] = sum_b H1[b][k] dz2[b][j] over this CTA's rows; thread owns k = 4kg + 128q + e, j = 4jg + 32q' + f
__device__ __forceinline__ void ws_wgrad_phase(const AgentCtx& C, const StepScalars& S, float* __restrict__ stage, const WsSmem& W, const WsTiles& T,
                                               float* __restrict__ gpart) {
  const NetLayout& L = C.L;
  const int tid = threadIdx.x, jg = tid & 7, kg = tid >> 3, wr = T.wr;
  const long long B = S.B;
  float acc[8][16];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 16; ++b) acc[a][b] = 0.f;
  int it = 0;
  const bool single = T.first + T.stride >= T.n_wide;     // the tile is still in shared memory (see ws_dgrad_phase)
  if (T.first < T.n_wide && !single) ws_stage_tile(C, stage, T.first * wr, B, wr, false);
  for (long long wt = T.first; wt < T.n_wide; wt += T.stride, ++it) {
    const float* cur = stage + (it & 1) * kWsTileFloats;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (wt + T.stride < T.n_wide) ws_stage_tile(C, stage + ((it + 1) & 1) * kWsTileFloats, (wt + T.stride) * wr, B, wr, false);
    const float* sH1 = (single ? W.h1 : cur) + 4 * kg;
    const float* sDZ = (single ? W.dz2 : cur + kWR * kH1) + 4 * jg;
#pragma unroll 2
    for (int r = 0; r < wr; ++r) {
      const float4 h0 = *reinterpret_cast<const float4*>(sH1 + r * kH1), h1 = *reinterpret_cast<const float4*>(sH1 + r * kH1 + 128);
      const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      float dz[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(sDZ + r * kH2 + 32 * q);
        dz[4 * q] = v.x; dz[4 * q + 1] = v.y; dz[4 * q + 2] = v.z; dz[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 16; ++b) acc[a][b] = fmaf(hv[a], dz[b], acc[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    float* row = gpart + L.off_w2t + (4 * kg + 128 * (a >> 2) + (a & 3)) * kW2LD + 4 * jg;
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(row + 32 * q) = make_float4(acc[a][4 * q], acc[a][4 * q + 1], acc[a][4 * q + 2], acc[a][4 * q + 3]);
  }
  __syncthreads();
}

// ---- after the agent barrier: gradient = sum of the per-CTA partial blobs in CTA order, then Adam / Polyak on the owner.
// Blob positions that are padding are zero in every partial (the partial buffers are zero-initialised and only parameter
// positions are ever written), so their "gradient" is 0 and the update leaves them as they are.
__device__ void ws_reduce_apply(const AgentCtx& C, const StepScalars& S, const float* __restrict__ gparts, int n_parts, int wid, int n_workers) {
  const int total = C.L.total, stride = n_workers * kThreads;
  constexpr int E = 4, PF = 16;       // elements per thread and partials per round: 64 loads in flight
  for (int base = wid * kThreads + static_cast<int>(threadIdx.x); base < total; base += E * stride) {
    ParamVals pv[E];
    float g[E];
#pragma unroll
    for (int k = 0; k < E; ++k) {
      g[k] = 0.f;
      if (base + k * stride < total) pv[k] = param_load(C, S, base + k * stride);
    }
    for (int c0 = 0; c0 < n_parts; c0 += PF) {
      float v[E][PF];
#pragma unroll
      for (int k = 0; k < E; ++k)
#pragma unroll
        for (int q = 0; q < PF; ++q)
          v[k][q] = (c0 + q < n_parts && base + k * stride < total) ? __ldcg(gparts + static_cast<size_t>(c0 + q) * total + base + k * stride) : 0.f;
#pragma unroll
      for (int k = 0; k < E; ++k)
#pragma unroll
        for (int q = 0; q < PF; ++q) g[k] += (c0 + q < n_parts) ? v[k][q] : 0.f;      // CTA order
    }
#pragma unroll
    for (int k = 0; k < E; ++k) {
      const int pi = base + k * stride;
      if (pi < total) {
        C.grads[pi] = g[k];
        param_apply(C, S, pi, g[k], pv[k]);
      }
    }
  }
}

