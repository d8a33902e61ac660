// rmc_infer64.cuh -- exact fp32 batched act / Q values for LARGE batches (BASELINE configs[2], 65,536 states).
//
// Same arithmetic contract as k_mlp_infer (fp32 operands, fp32 FMA accumulation, act = ReLU | ELU) but tiled for
// throughput instead of latency: a CTA owns 64 rows per pass, so every weight fetched from shared memory is used
// for 64 rows (k_mlp_infer: 8) and layer 2 needs no K-split / partial reduction.
//
//   layer 1   thread = 2 rows x 32 columns (lanes over rows, warp = 32 columns): x by 8-byte loads, weights broadcast
//   layer 2   thread = 4 rows x 8 columns, warp = 16 rows x 64 columns: per k two 128-byte weight wavefronts + one 64-byte
//             activation wavefront feed 32 FMAs per thread -> the FMA pipe, not shared memory, is the limiter
//   heads     thread = (row, 4 heads)
// Activations live transposed in shared memory ([k][64 rows]); H2 overwrites H1 after layer 2.  Shared memory:
// parameter blob (155.7 KB for D = 14) + 64 KB + 4 KB -- which is why this form serves obs_dim <= 16 only.
#pragma once
#include "rmc_mlp.cuh"

namespace rmc {

constexpr int kBigRows = 64;
constexpr int kBigXFloats = 16 * kBigRows;       // x^T [16][64]; later the head outputs [64][16]
__host__ __device__ inline int big_smem_floats(int param_floats) { return param_floats + kH1 * kBigRows + kBigXFloats + 4; }

__global__ void __launch_bounds__(kThreads, 1) k_mlp_infer64(NetLayout L, const float* __restrict__ params, const float* __restrict__ obs,
                                                             long long n, long long* __restrict__ actions, float* __restrict__ q_out, int mode) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sH = smem + L.total;                    // H1^T [256][64], then H2^T [128][64]
  float* sX = sH + kH1 * kBigRows;               // x^T [16][64], then heads [64][16]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sX + kBigXFloats);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_tiles = (n + kBigRows - 1) / kBigRows;
  uint32_t parity = 0;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  stage_params(sW, params, L.total, bar, parity);
  wait_params(bar, parity);
  const int D = L.D, act = L.act;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kBigRows;
    // ---- x^T
    for (int t = tid; t < kBigRows * D; t += kThreads) {
      const int r = t / D, d = t - r * D;
      const long long i = row0 + r;
      sX[d * kBigRows + r] = (i < n) ? __ldg(obs + i * D + d) : 0.f;
    }
    __syncthreads();
    // ---- layer 1: rows 2*lane, 2*lane+1; columns 32*warp .. 32*warp+31
    {
      float a0[32], a1[32];
      const float* b0 = sW + L.off_b0 + 32 * warp;
#pragma unroll
      for (int c = 0; c < 32; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
      for (int d = 0; d < D; ++d) {
        const float2 x = *reinterpret_cast<const float2*>(sX + d * kBigRows + 2 * lane);
        const float* w = sW + L.off_w0t + d * kH1 + 32 * warp;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(w + c);
          a0[c] = fmaf(x.x, wv.x, a0[c]); a0[c + 1] = fmaf(x.x, wv.y, a0[c + 1]); a0[c + 2] = fmaf(x.x, wv.z, a0[c + 2]); a0[c + 3] = fmaf(x.x, wv.w, a0[c + 3]);
          a1[c] = fmaf(x.y, wv.x, a1[c]); a1[c + 1] = fmaf(x.y, wv.y, a1[c + 1]); a1[c + 2] = fmaf(x.y, wv.z, a1[c + 2]); a1[c + 3] = fmaf(x.y, wv.w, a1[c + 3]);
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float bb = b0[c];
        *reinterpret_cast<float2*>(sH + (32 * warp + c) * kBigRows + 2 * lane) = make_float2(act_fwd(a0[c] + bb, act), act_fwd(a1[c] + bb, act));
      }
    }
    __syncthreads();
    // ---- layer 2: warp = 16 rows x 64 columns; lane = 4 rows x (4 + 4) columns
    {
      const int ch = warp & 1, rq = warp >> 1, cgi = lane & 7, rgi = lane >> 3;
      const int c0 = 64 * ch + 4 * cgi, c1 = c0 + 32, r0 = 16 * rq + 4 * rgi;
      float acc[8][4];
#pragma unroll
      for (int c = 0; c < 8; ++c) { acc[c][0] = 0.f; acc[c][1] = 0.f; acc[c][2] = 0.f; acc[c][3] = 0.f; }
      const float* w2 = sW + L.off_w2t;
#pragma unroll 8
      for (int k = 0; k < kH1; ++k) {
        const float4 wa = *reinterpret_cast<const float4*>(w2 + k * kW2LD + c0);
        const float4 wb = *reinterpret_cast<const float4*>(w2 + k * kW2LD + c1);
        const float4 h = *reinterpret_cast<const float4*>(sH + k * kBigRows + r0);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          acc[c][0] = fmaf(h.x, wv[c], acc[c][0]); acc[c][1] = fmaf(h.y, wv[c], acc[c][1]);
          acc[c][2] = fmaf(h.z, wv[c], acc[c][2]); acc[c][3] = fmaf(h.w, wv[c], acc[c][3]);
        }
      }
      __syncthreads();                           // every warp has finished reading H1: H2^T may overwrite it
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = (c < 4) ? c0 + c : c1 + (c - 4);
        const float bb = sW[L.off_b2 + col];
        *reinterpret_cast<float4*>(sH + col * kBigRows + r0) =
            make_float4(act_fwd(acc[c][0] + bb, act), act_fwd(acc[c][1] + bb, act), act_fwd(acc[c][2] + bb, act), act_fwd(acc[c][3] + bb, act));
      }
    }
    __syncthreads();
    // ---- heads: thread = (row, heads 4*hg .. 4*hg+3)
    {
      const int row = tid & 63, hg = tid >> 6;
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < kH2; k += 4) {
        const float h0 = sH[k * kBigRows + row], h1v = sH[(k + 1) * kBigRows + row], h2v = sH[(k + 2) * kBigRows + row], h3v = sH[(k + 3) * kBigRows + row];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int a = 4 * hg + j;
          if (a < L.NH) {
            const float4 w = *reinterpret_cast<const float4*>(sW + L.off_wh + a * kH2 + k);
            s[j] = fmaf(h0, w.x, s[j]); s[j] = fmaf(h1v, w.y, s[j]); s[j] = fmaf(h2v, w.z, s[j]); s[j] = fmaf(h3v, w.w, s[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = 4 * hg + j;
        sX[row * kQLD + a] = (a < L.NH) ? s[j] + sW[L.off_bh + a] : 0.f;      // x^T is dead: the head outputs take its place
      }
    }
    __syncthreads();
    // ---- outputs (row = tid < 64)
    if (tid < kBigRows) {
      const long long i = row0 + tid;
      if (i < n) {
        float hd[kQLD];
#pragma unroll
        for (int a = 0; a < kQLD; a += 4) *reinterpret_cast<float4*>(hd + a) = *reinterpret_cast<const float4*>(sX + tid * kQLD + a);
        if (mode == 2) {
#pragma unroll
          for (int a = 0; a < kQLD; ++a)
            if (a < L.NH) q_out[i * L.NH + a] = hd[a];
        } else if (mode == 0) {                  // dueling: argmax of the RAW advantages (network.py:110-117); plain: argmax Q
          int best = 0;
          float bv = L.dueling ? hd[1] : hd[0];
#pragma unroll
          for (int a = 1; a < 15; ++a) {
            const float v = L.dueling ? hd[a + 1] : hd[a];
            if (a < L.A && v > bv) { bv = v; best = a; }       // strict '>' keeps the first maximum (torch.argmax)
          }
          actions[i] = best;
        } else {                                 // Q = val + (adv - mean(adv))   (network.py:83)
          if (L.dueling) {
            float sum = 0.f;
#pragma unroll
            for (int a = 1; a < kQLD; ++a) sum += (a <= L.A) ? hd[a] : 0.f;
            const float mean = sum / static_cast<float>(L.A);
#pragma unroll
            for (int a = 0; a < kQLD - 1; ++a)
              if (a < L.A) q_out[i * L.A + a] = hd[0] + (hd[1 + a] - mean);
          } else {
#pragma unroll
            for (int a = 0; a < kQLD; ++a)
              if (a < L.A) q_out[i * L.A + a] = hd[a];
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace rmc
