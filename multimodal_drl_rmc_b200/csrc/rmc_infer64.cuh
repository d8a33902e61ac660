// rmc_infer64.cuh -- exact fp32 batched act / Q values for LARGE batches (BASELINE configs[2], 65,536 states).
//
// Same arithmetic contract as k_mlp_infer (fp32 operands, fp32 FMA accumulation, act = ReLU | ELU) but tiled for
// throughput instead of latency: a CTA owns 64 rows per pass, so every weight fetched from shared memory is used
// for 64 rows (k_mlp_infer: 8) and layer 2 needs no K-split / partial reduction.
//
//   layer 1   thread = 8 rows x 8 columns, warp = 64 rows x 32 columns
//   layer 2   thread = 8 rows x 8 columns, K split over the two halves of the CTA (partials meet in the dead H1 tile):
//             1 byte of shared-memory traffic per FMA, the point where the FMA pipes stop starving (see the kernel)
//   heads     thread = (row, 4 heads)
// Activations live transposed in shared memory ([k][64 rows]); H2 overwrites H1 after layer 2.  Shared memory:
// parameter blob (155.7 KB for D = 14) + 64 KB + 4 KB -- which is why this form serves obs_dim <= 16 only.
#pragma once
#include "rmc_mlp.cuh"

namespace rmc {

constexpr int kBigRows = 64;
constexpr int kBigXFloats = 16 * kBigRows;       // x^T [16][64]; later the head outputs [64][16]
__host__ __device__ inline int big_smem_floats(int param_floats) { return param_floats + kH1 * kBigRows + kBigXFloats + 4; }

constexpr int kBigThreads = 512;                 // 16 warps: four per scheduler (ncu on the 8-warp form: dispatch / wait /
                                                 // short-scoreboard stalls with two warps per scheduler held the FMA pipe at 47 %)

__global__ void __launch_bounds__(kBigThreads, 1) k_mlp_infer64(NetLayout L, const float* __restrict__ params, const float* __restrict__ obs,
                                                                long long n, long long* __restrict__ actions, float* __restrict__ q_out, int mode) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sH = smem + L.total;                    // H1^T [256][64], then partial sums and H2^T [128][64]
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
  // Thread tiles: shared memory hands 128 B per clock to the register file whatever the broadcast pattern (an LDS.128 of a
  // warp costs 4 clocks) and the FMA pipes retire 128 FMAs per clock, so a tile of R x C outputs, which loads 4 (R + C)
  // bytes for R C FMAs per k, must be at least 8 x 8 not to starve the FMA pipes (4 x 8: 1.5 B per FMA, at most 67 %).
  // lane & 7 -> rows {4 rt .. 4 rt + 3} and {32 + 4 rt ..} (a quarter-warp reads 128 contiguous bytes of H^T);
  // lane >> 3 and the warp -> consecutive columns (a quarter-warp shares its weights: broadcast).
  const int rt = lane & 7, cq = lane >> 3;
  // The tile's 64 x D inputs are one contiguous run of obs (element t of the run = row t / D, feature t % D): each thread
  // carries at most two of them in registers, fetched one tile ahead so that the global-memory latency hides behind the
  // previous tile's layers.
  float xr[2];
  auto fetch = [&](long long tile_) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int t = tid + j * kBigThreads;
      const long long e = tile_ * kBigRows * D + t;
      xr[j] = (tile_ < n_tiles && t < kBigRows * D && e < n * D) ? __ldg(obs + e) : 0.f;
    }
  };
  fetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kBigRows;
    // ---- x^T
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int t = tid + j * kBigThreads;
      if (t < kBigRows * D) {
        const int r = t / D, d = t - r * D;
        sX[d * kBigRows + r] = xr[j];
      }
    }
    __syncthreads();
    fetch(tile + gridDim.x);
    // ---- layer 1: thread = 8 rows x 4 columns, warp = 64 rows x 16 columns (K = D is short: no split)
    {
      const int c0 = 16 * warp + 4 * cq;
      float acc[4][8];                                 // [column][row]
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[c][r] = 0.f;
      for (int d = 0; d < D; ++d) {
        const float4 xa = *reinterpret_cast<const float4*>(sX + d * kBigRows + 4 * rt);
        const float4 xb = *reinterpret_cast<const float4*>(sX + d * kBigRows + 32 + 4 * rt);
        const float4 wa = *reinterpret_cast<const float4*>(sW + L.off_w0t + d * kH1 + c0);
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float wv[4] = {wa.x, wa.y, wa.z, wa.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < 8; ++r) acc[c][r] = fmaf(xv[r], wv[c], acc[c][r]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float bb = sW[L.off_b0 + c0 + c];
        float* dst = sH + (c0 + c) * kBigRows;
        *reinterpret_cast<float4*>(dst + 4 * rt) =
            make_float4(act_fwd(acc[c][0] + bb, act), act_fwd(acc[c][1] + bb, act), act_fwd(acc[c][2] + bb, act), act_fwd(acc[c][3] + bb, act));
        *reinterpret_cast<float4*>(dst + 32 + 4 * rt) =
            make_float4(act_fwd(acc[c][4] + bb, act), act_fwd(acc[c][5] + bb, act), act_fwd(acc[c][6] + bb, act), act_fwd(acc[c][7] + bb, act));
      }
    }
    __syncthreads();
    // ---- layer 2: thread = 8 rows x 8 columns; the 64 x 128 outputs need 128 such threads, so the four quarters of the
    //      CTA split K (quarter q: k in [64 q, 64 q + 64)) and their partial sums meet in the dead H1 tile
    {
      const int kq = warp >> 2;
      const int c0 = 32 * (warp & 3) + 8 * cq;
      float acc[8][8];
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[c][r] = 0.f;
      const float* w2 = sW + L.off_w2t + (kq * (kH1 / 4)) * kW2LD + c0;
      const float* hp = sH + (kq * (kH1 / 4)) * kBigRows + 4 * rt;
#pragma unroll 2
      for (int k = 0; k < kH1 / 4; ++k) {
        const float4 wa = *reinterpret_cast<const float4*>(w2 + k * kW2LD);
        const float4 wb = *reinterpret_cast<const float4*>(w2 + k * kW2LD + 4);
        const float4 ha = *reinterpret_cast<const float4*>(hp + k * kBigRows);
        const float4 hb = *reinterpret_cast<const float4*>(hp + k * kBigRows + 32);
        const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int r = 0; r < 8; ++r) acc[c][r] = fmaf(hv[r], wv[c], acc[c][r]);
      }
      // (q0 + q2) + (q1 + q3) + bias -> activation -> H2^T [128][64] at the start of the tile buffer
      float* p_lo = sH;                            // [128][64]: quarter 2's sums, then quarter 1's, then H2^T
      float* p_hi = sH + kH2 * kBigRows;           // [128][64]: quarter 3's sums
      auto put = [&](float* base) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float* dst = base + (c0 + c) * kBigRows;
          *reinterpret_cast<float4*>(dst + 4 * rt) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
          *reinterpret_cast<float4*>(dst + 32 + 4 * rt) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
        }
      };
      auto add = [&](const float* base) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float* src = base + (c0 + c) * kBigRows;
          const float4 pa = *reinterpret_cast<const float4*>(src + 4 * rt);
          const float4 pb = *reinterpret_cast<const float4*>(src + 32 + 4 * rt);
          acc[c][0] += pa.x; acc[c][1] += pa.y; acc[c][2] += pa.z; acc[c][3] += pa.w;
          acc[c][4] += pb.x; acc[c][5] += pb.y; acc[c][6] += pb.z; acc[c][7] += pb.w;
        }
      };
      __syncthreads();                           // every warp has finished reading H1
      if (kq == 2) put(p_lo);
      if (kq == 3) put(p_hi);
      __syncthreads();
      if (kq == 0) add(p_lo);
      if (kq == 1) add(p_hi);
      __syncthreads();
      if (kq == 1) put(p_lo);
      __syncthreads();
      if (kq == 0) {
        add(p_lo);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float bb = sW[L.off_b2 + c0 + c];
#pragma unroll
          for (int r = 0; r < 8; ++r) acc[c][r] = act_fwd(acc[c][r] + bb, act);
        }
        put(p_lo);
      }
    }
    __syncthreads();
    // ---- heads: thread = (row, heads 4*hg .. 4*hg+3, half of K); the upper half's sums pass through the free half of the tile buffer
    {
      const int row = tid & 63, hg = (tid >> 6) & 3, kh = tid >> 8;
      float* hpart = sH + kH2 * kBigRows;          // [256 threads][4]
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = kh * (kH2 / 2); k < (kh + 1) * (kH2 / 2); k += 4) {
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
      if (kh == 1) *reinterpret_cast<float4*>(hpart + (tid & 255) * 4) = make_float4(s[0], s[1], s[2], s[3]);
      __syncthreads();
      if (kh == 0) {
        const float4 o = *reinterpret_cast<const float4*>(hpart + tid * 4);
        const float so[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int a = 4 * hg + j;
          sX[row * kQLD + a] = (a < L.NH) ? (s[j] + so[j]) + sW[L.off_bh + a] : 0.f;      // x^T is dead: the head outputs take its place
        }
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
