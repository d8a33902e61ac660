// rmc_hybrid.cuh -- the repo-HEAD network of the reference behind the same learner step:
// env/dqn_config.py:66-193 TwoStreamHybridNetwork
//     state[284] = macro[14] | grid[2][27][5]
//     grid -> Conv(2->32, 3x3, s(1,1), p1) -> act -> Conv(32->64, s(2,1)) -> act -> Conv(64->64, s(2,2)) -> act -> flatten[1344]
//     cat(flatten, macro)[1358] -> Linear(512) -> act -> Linear(256) -> act -> {fc_val[1], fc_adv[A]} | fc_out[A]
// Exact fp32 path (fp32 operands, fmaf accumulation), one kernel per layer and direction; the GEMM-shaped pieces
// (dense layers and the convolutions as implicit GEMMs) share one tiled kernel template.
// Parameters live in ONE flat blob in torch state_dict() order (conv weights [oc][ic][3][3], linear weights
// [out][in]), so set/get_params are plain copies and the Adam kernel walks the blob linearly.
// Activations of a pass live in a per-row scratch record (HybNet::rec floats): conv outputs | features (last conv
// output ++ macro) | dense outputs | heads[16]; the backward pass writes deltas into a record of the same layout.
#pragma once
#include "rmc_mlp.cuh"

namespace rmc {

constexpr int kHybMaxConv = 4, kHybMaxDense = 3;

struct HybConv { int ic, ih, iw, oc, oh, ow, sh, sw; int w_off, b_off; int in_off, out_off; };   // offsets: parameters / record
struct HybDense { int in, out; int w_off, b_off; int in_off, out_off; };
struct HybNet {
  int n_conv, n_dense;
  HybConv conv[kHybMaxConv];
  HybDense dense[kHybMaxDense];
  int macro_len, grid_len, D;       // D = macro_len + grid_len (state length)
  int feat_off, feat_len, conv_flat;// features = [last conv output (conv_flat) | macro]
  int last_off, last_len;           // output of the last dense layer (input of the heads)
  int head_off;                     // heads[16] inside the record
  int hw_off[2], hb_off[2];         // dueling: {fc_val, fc_adv}; plain: {fc_out, -}
  int A, NH, dueling, act;
  int rec;                          // floats per record (multiple of 4)
  int total;                        // parameter floats (torch order), padded to a multiple of 4
};

// which gathered row / column offset feeds pass-row r: online pass = [s' rows | s rows], target pass = s' rows,
// inference = plain [n][D] matrix
struct HybSrc { const float* base; long long stride; long long B; int off_first, off_second; };
__device__ __forceinline__ const float* hyb_src_row(const HybSrc& s, long long r) {
  return (r < s.B) ? s.base + r * s.stride + s.off_first : s.base + (r - s.B) * s.stride + s.off_second;
}

// ---------------------------------------------------------------------------------------------- convolution forward
// grid = (rows, slices): the layer's whole input (<= 4480 floats) of the row is staged in shared memory; a work item is 4
// output channels of one output pixel (every input value fetched feeds 4 FMAs; weights are warp-broadcast L1 reads) and
// a CTA owns a slice of 256 items, so even a 32-row batch spreads over > 148 CTAs.
__global__ void __launch_bounds__(256) k_hyb_conv_fwd(HybNet N, int li, const float* __restrict__ P, HybSrc src, float* __restrict__ rec_base) {
  pdl_enter();
  extern __shared__ float s_in[];
  const HybConv c = N.conv[li];
  const long long r = blockIdx.x;
  float* rec = rec_base + r * N.rec;
  const float* in = (li == 0) ? hyb_src_row(src, r) + N.macro_len : rec + c.in_off;
  const int n_in = c.ic * c.ih * c.iw;
  for (int t = threadIdx.x; t < n_in; t += blockDim.x) s_in[t] = in[t];
  if (li == N.n_conv - 1 && blockIdx.y == 0)          // features = [flattened last conv output | macro]
    for (int t = threadIdx.x; t < N.macro_len; t += blockDim.x) rec[N.feat_off + N.conv_flat + t] = hyb_src_row(src, r)[t];
  __syncthreads();
  const int npix = c.oh * c.ow, ng = c.oc >> 2;
  const float* W = P + c.w_off;
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < ng * npix; t += gridDim.y * blockDim.x) {
    const int g = t / npix, pix = t - g * npix;
    const int oy = pix / c.ow, ox = pix - oy * c.ow;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int wstride = c.ic * 9;
    const float* w0 = W + (4 * g) * wstride;
    for (int ic = 0; ic < c.ic; ++ic) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * c.sh + ky - 1;
        if (iy < 0 || iy >= c.ih) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * c.sw + kx - 1;
          if (ix < 0 || ix >= c.iw) continue;
          const float x = s_in[(ic * c.ih + iy) * c.iw + ix];
          const float* w = w0 + ic * 9 + ky * 3 + kx;
          a0 = fmaf(x, __ldg(w), a0);
          a1 = fmaf(x, __ldg(w + wstride), a1);
          a2 = fmaf(x, __ldg(w + 2 * wstride), a2);
          a3 = fmaf(x, __ldg(w + 3 * wstride), a3);
        }
      }
    }
    const float* b = P + c.b_off + 4 * g;
    float* o = rec + c.out_off + (4 * g) * npix + pix;
    o[0] = act_fwd(a0 + __ldg(b), N.act);
    o[npix] = act_fwd(a1 + __ldg(b + 1), N.act);
    o[2 * npix] = act_fwd(a2 + __ldg(b + 2), N.act);
    o[3 * npix] = act_fwd(a3 + __ldg(b + 3), N.act);
  }
}

// features = [flattened last conv output | macro]: the macro part of every pass row
__global__ void k_hyb_copy_macro(HybNet N, HybSrc src, float* __restrict__ rec_base, long long R) {
  pdl_enter();
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t >= R * N.macro_len) return;
  const long long r = t / N.macro_len;
  const int d = static_cast<int>(t - r * N.macro_len);
  rec_base[r * N.rec + N.feat_off + N.conv_flat + d] = hyb_src_row(src, r)[d];
}

// ---------------------------------------------------------------------------------------------- generic SGEMM
// C[m][n] = sum_k A(m,k) * B(k,n), 64x64x16 tiles, 256 threads, 4x4 register tiles, fixed summation order.
// MODE 0  strided operands (covers x.W^T, dY.W and dY^T.X of the dense layers)
// MODE 1  convolution forward as an implicit GEMM: m = (row, output pixel), k = (ic, ky, kx), n = oc;
//         A is the im2col view of the layer input (zero outside the image), B = W[oc][k]
// MODE 2  convolution data gradient: m = (row, input pixel), k = (oc, ky, kx), n = ic; A gathers the output deltas that the
//         input pixel feeds (stride-aware), B = W[oc][ic][ky][kx]
// MODE 3  convolution weight gradient: m = oc, k = (row, output pixel), n = (ic, ky, kx); A = output deltas, B = im2col
//   epi 0: C = act(C + bias[n])            (forward)
//   epi 1: C = C * act'(H at C's address)  (data gradient; H = the forward activation below)
//   epi 2: C as is                         (weight gradient, written into the gradient blob)
struct HybGemm {
  const float* A; long long a_sm, a_sk;
  const float* B; long long b_sk, b_sn;
  float* C; long long c_sm;
  const float* bias;
  const float* H; long long h_sm;
  int M, N, K, epi, act;
  int splits, k_chunk;              // split-K: blockIdx.z owns k in [z*k_chunk, (z+1)*k_chunk); raw partial sums go to ws[z][M][N]
  float* ws;
  // convolution geometry (MODE 1-3): image operand = img(row) + (c*ih + y)*iw + x, delta operand = dz + row*dz_stride + oc*npix + pix
  HybConv cv;
  const float* img; long long img_stride;       // layer input of row r (nullptr: take it from `src`, row src_row0 + r, after the macro part)
  HybSrc src; long long src_row0; int macro_len;
  const float* dz; long long dz_stride;
};
__device__ __forceinline__ const float* hyb_img_row(const HybGemm& G, long long r) {
  return (G.img != nullptr) ? G.img + r * G.img_stride : hyb_src_row(G.src, G.src_row0 + r) + G.macro_len;
}
// im2col element: row r, output pixel pix, k9 = (ic, ky, kx)
__device__ __forceinline__ float hyb_im2col(const HybGemm& G, int r, int pix, int k9) {
  const HybConv& c = G.cv;
  const int ic = k9 / 9, q = k9 - 9 * ic, ky = q / 3, kx = q - 3 * ky;
  const int oy = pix / c.ow, ox = pix - oy * c.ow;
  const int iy = oy * c.sh + ky - 1, ix = ox * c.sw + kx - 1;
  if (iy < 0 || iy >= c.ih || ix < 0 || ix >= c.iw) return 0.f;
  return __ldg(hyb_img_row(G, r) + (ic * c.ih + iy) * c.iw + ix);
}
template <int MODE>
__device__ __forceinline__ float hyb_load_a(const HybGemm& G, int m, int k) {
  if constexpr (MODE == 0) return __ldg(G.A + m * G.a_sm + k * G.a_sk);
  const HybConv& c = G.cv;
  const int npix = c.oh * c.ow;
  if constexpr (MODE == 1) { const int r = m / npix; return hyb_im2col(G, r, m - r * npix, k); }
  if constexpr (MODE == 2) {
    const int ipix = c.ih * c.iw, r = m / ipix, p = m - r * ipix, iy = p / c.iw, ix = p - iy * c.iw;
    const int oc = k / 9, q = k - 9 * oc, ky = q / 3, kx = q - 3 * ky;
    const int ny = iy + 1 - ky, nx = ix + 1 - kx;
    if (ny < 0 || nx < 0 || ny % c.sh != 0 || nx % c.sw != 0) return 0.f;
    const int oy = ny / c.sh, ox = nx / c.sw;
    if (oy >= c.oh || ox >= c.ow) return 0.f;
    return __ldg(G.dz + r * G.dz_stride + oc * npix + oy * c.ow + ox);
  }
  const int r = k / npix;                                  // MODE 3
  return __ldg(G.dz + r * G.dz_stride + m * npix + (k - r * npix));
}
template <int MODE>
__device__ __forceinline__ float hyb_load_b(const HybGemm& G, int k, int n) {
  if constexpr (MODE == 0 || MODE == 1) return __ldg(G.B + k * G.b_sk + n * G.b_sn);
  const HybConv& c = G.cv;
  if constexpr (MODE == 2) { const int oc = k / 9; return __ldg(G.B + (oc * c.ic + n) * 9 + (k - 9 * oc)); }
  const int npix = c.oh * c.ow, r = k / npix;              // MODE 3
  return hyb_im2col(G, r, k - r * npix, n);
}
// address of C[m][n] (and of the matching activation for epi 1) relative to G.C / G.H
template <int MODE>
__device__ __forceinline__ long long hyb_c_index(const HybGemm& G, int m, int n) {
  if constexpr (MODE == 1) { const int npix = G.cv.oh * G.cv.ow, r = m / npix; return r * G.c_sm + n * npix + (m - r * npix); }
  if constexpr (MODE == 2) { const int ipix = G.cv.ih * G.cv.iw, r = m / ipix; return r * G.c_sm + n * ipix + (m - r * ipix); }
  return m * G.c_sm + n;
}
__device__ __forceinline__ float hyb_epilogue(const HybGemm& G, float v, int n, long long ci) {
  if (G.epi == 0) return act_fwd(v + __ldg(G.bias + n), G.act);
  if (G.epi == 1) return act_bwd(v, __ldg(G.H + ci), G.act);
  return v;
}
template <int MODE>
__global__ void __launch_bounds__(256) k_hyb_gemm(HybGemm G) {
  pdl_enter();
  __shared__ __align__(16) float sA[16][68], sB[16][68];      // rows padded to 68 floats: 16-byte aligned 4-wide fragments
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // thread tile: rows 4*ty.., columns 4*tx..
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping follows the unit-stride dimension of each operand; the NEXT k-tile is fetched into registers while the
  // current one is multiplied (the global-load latency of a tile would otherwise be exposed once per 16 k)
  const bool a_kfast = (MODE == 0) ? (G.a_sk == 1) : (MODE == 3), b_kfast = (MODE == 0 || MODE == 1) ? (G.b_sk == 1) : (MODE == 3);
  float ra[4], rb[4];
  auto fetch = [&](int k0, int k_end) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int t = tid + 256 * e;                     // 1024 elements per operand tile
      {
        const int kk = a_kfast ? (t & 15) : (t >> 6), mm = a_kfast ? (t >> 4) : (t & 63);
        const int m = m0 + mm, k = k0 + kk;
        ra[e] = (m < G.M && k < k_end) ? hyb_load_a<MODE>(G, m, k) : 0.f;
      }
      {
        const int kk = b_kfast ? (t & 15) : (t >> 6), nn = b_kfast ? (t >> 4) : (t & 63);
        const int n = n0 + nn, k = k0 + kk;
        rb[e] = (n < G.N && k < k_end) ? hyb_load_b<MODE>(G, k, n) : 0.f;
      }
    }
  };
  const int k_lo = (G.splits > 1) ? blockIdx.z * G.k_chunk : 0;
  const int k_hi = (G.splits > 1) ? min(G.K, k_lo + G.k_chunk) : G.K;
  fetch(k_lo, k_hi);
  for (int k0 = k_lo; k0 < k_hi; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int t = tid + 256 * e;
      sA[a_kfast ? (t & 15) : (t >> 6)][a_kfast ? (t >> 4) : (t & 63)] = ra[e];
      sB[b_kfast ? (t & 15) : (t >> 6)][b_kfast ? (t >> 4) : (t & 63)] = rb[e];
    }
    __syncthreads();
    if (k0 + 16 < k_hi) fetch(k0 + 16, k_hi);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[kk][4 * ty]), bv = *reinterpret_cast<const float4*>(&sB[kk][4 * tx]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + 4 * ty + i;
    if (m >= G.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + 4 * tx + j;
      if (n >= G.N) continue;
      if (G.splits > 1) {
        G.ws[(static_cast<long long>(blockIdx.z) * G.M + m) * G.N + n] = acc[i][j];
      } else {
        const long long ci = hyb_c_index<MODE>(G, m, n);
        G.C[ci] = hyb_epilogue(G, acc[i][j], n, ci);
      }
    }
  }
}

// split-K second pass: C = epilogue(sum_z ws[z]) in split order (deterministic)
template <int MODE>
__global__ void __launch_bounds__(256) k_hyb_splitk_reduce(HybGemm G) {
  pdl_enter();
  const long long t = blockIdx.x * 256ll + threadIdx.x;
  if (t >= static_cast<long long>(G.M) * G.N) return;
  const int m = static_cast<int>(t / G.N), n = static_cast<int>(t - static_cast<long long>(m) * G.N);
  float v = 0.f;
  for (int z = 0; z < G.splits; ++z) v += __ldcg(G.ws + (static_cast<long long>(z) * G.M + m) * G.N + n);
  const long long ci = hyb_c_index<MODE>(G, m, n);
  G.C[ci] = hyb_epilogue(G, v, n, ci);
}

// per-channel sums of the output deltas of a conv layer, stage 1: CTA (oc, slice) sums its slice of the rows with a
// fixed-order tree -> part[slice][oc]; stage 2 is k_hyb_colsum over the slices
__global__ void __launch_bounds__(256) k_hyb_conv_bias_grad(const float* __restrict__ dz, long long dz_stride, int npix, long long B, float* __restrict__ part) {
  pdl_enter();
  __shared__ float s_red[256];
  const int oc = blockIdx.x;
  const long long rows_per = (B + gridDim.y - 1) / gridDim.y, r_lo = blockIdx.y * rows_per, r_hi = min(B, r_lo + rows_per);
  float s = 0.f;
  for (long long t = r_lo * npix + threadIdx.x; t < r_hi * npix; t += 256) {
    const long long r = t / npix;
    s += __ldg(dz + r * dz_stride + oc * npix + (t - r * npix));
  }
  s_red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) s_red[threadIdx.x] += s_red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + oc] = s_red[0];
}

// column sums: out[n] = sum_m X[m*ld + n]   (bias gradients of the dense layers), one thread per column, fixed order
__global__ void k_hyb_colsum(const float* __restrict__ X, long long ld, int M, int N, float* __restrict__ out) {
  pdl_enter();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += __ldg(X + m * ld + n);
  out[n] = s;
}

// ---------------------------------------------------------------------------------------------- heads
// warp per row: heads[a] = <h, W_a> + b_a  (fc_val / fc_adv or fc_out; network.py:54-63,81-96)
__global__ void __launch_bounds__(256) k_hyb_heads_fwd(HybNet N, const float* __restrict__ P, float* __restrict__ rec_base, long long R) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = blockIdx.x * 8ll + warp;
  if (r >= R) return;
  float* rec = rec_base + r * N.rec;
  const float* h = rec + N.last_off;
  float out = 0.f;
  if (N.last_len <= 256) {
    // every head's weights are requested before the first reduction (the dot products are latency-, not bandwidth-shaped);
    // per head the same order as below: k = lane, lane + 32, ... then the xor tree
    float hv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hv[j] = (lane + 32 * j < N.last_len) ? h[lane + 32 * j] : 0.f;
#pragma unroll
    for (int a0 = 0; a0 < kQLD; a0 += 4) {
      float wv[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int a = a0 + u;
        const float* w = N.dueling ? (a == 0 ? P + N.hw_off[0] : P + N.hw_off[1] + (a - 1) * N.last_len) : P + N.hw_off[0] + a * N.last_len;
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[u][j] = (a < N.NH && lane + 32 * j < N.last_len) ? __ldg(w + lane + 32 * j) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int a = a0 + u;
        if (a < N.NH) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (lane + 32 * j < N.last_len) s = fmaf(hv[j], wv[u][j], s);
#pragma unroll
          for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
          const float b = N.dueling ? (a == 0 ? __ldg(P + N.hb_off[0]) : __ldg(P + N.hb_off[1] + a - 1)) : __ldg(P + N.hb_off[0] + a);
          if (lane == a) out = s + b;
        }
      }
    }
  } else
  for (int a = 0; a < N.NH; ++a) {
    const float* w = N.dueling ? (a == 0 ? P + N.hw_off[0] : P + N.hw_off[1] + (a - 1) * N.last_len) : P + N.hw_off[0] + a * N.last_len;
    float s = 0.f;
    for (int k = lane; k < N.last_len; k += 32) s = fmaf(h[k], __ldg(w + k), s);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    const float b = N.dueling ? (a == 0 ? __ldg(P + N.hb_off[0]) : __ldg(P + N.hb_off[1] + a - 1)) : __ldg(P + N.hb_off[0] + a);
    if (lane == a) out = s + b;
  }
  if (lane < kQLD) rec[N.head_off + lane] = (lane < N.NH) ? out : 0.f;
}

__device__ __forceinline__ void hyb_heads_to_q(const float* h, int A, int dueling, float (&q)[kQLD]) {
  if (dueling) {        // Q = val + (adv - mean(adv))   (dqn/network.py:83)
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < kQLD - 1; ++a) sum += (a < A) ? h[1 + a] : 0.f;
    const float mean = sum / static_cast<float>(A);
#pragma unroll
    for (int a = 0; a < kQLD - 1; ++a) q[a] = h[0] + (h[1 + a] - mean);
    q[kQLD - 1] = 0.f;
  } else {
#pragma unroll
    for (int a = 0; a < kQLD; ++a) q[a] = h[a];
  }
}

// TD target, |td|, Huber, loss partials, head deltas (dqn/agent.py:166-185 / 204-226 / 245-272)
// online record r < B: s' row; r >= B: s row.  Head deltas go to the delta record of sample i.
__global__ void __launch_bounds__(128) k_hyb_td(AgentCtx C, StepScalars S, HybNet N, const float* __restrict__ rec_on, const float* __restrict__ rec_tg,
                                                float* __restrict__ drec) {
  pdl_enter();
  __shared__ float s_part[4];
  const long long i = blockIdx.x * 128ll + threadIdx.x;
  const bool per = S.prioritized != 0;
  float lterm = 0.f;
  if (i < S.B) {
    float hn[kQLD], ht[kQLD], hs[kQLD], qn[kQLD], qt[kQLD], qs[kQLD];
#pragma unroll
    for (int a = 0; a < kQLD; ++a) {
      hn[a] = rec_on[i * N.rec + N.head_off + a];
      ht[a] = rec_tg[i * N.rec + N.head_off + a];
      hs[a] = rec_on[(S.B + i) * N.rec + N.head_off + a];
    }
    hyb_heads_to_q(hn, N.A, N.dueling, qn);
    hyb_heads_to_q(ht, N.A, N.dueling, qt);
    hyb_heads_to_q(hs, N.A, N.dueling, qs);
    float qsel = qt[0];
    if (S.double_dqn) {
      int astar = 0;
      float bv = qn[0];
#pragma unroll
      for (int a = 1; a < kQLD - 1; ++a)
        if (a < N.A && qn[a] > bv) { bv = qn[a]; astar = a; }
#pragma unroll
      for (int a = 1; a < kQLD - 1; ++a) qsel = (a == astar) ? qt[a] : qsel;
    } else {
#pragma unroll
      for (int a = 1; a < kQLD - 1; ++a) qsel = (a < N.A) ? fmaxf(qsel, qt[a]) : qsel;
    }
    const int rf = C.rp.row_floats;
    const int act = __float_as_int(__ldcg(C.X + i * rf + 2 * N.D));
    const float rew = __ldcg(C.X + i * rf + 2 * N.D + 1), done = __ldcg(C.X + i * rf + 2 * N.D + 2);
    const float w = per ? C.is_w[i] : 1.f;
    const float y = rew + ((1.f - done) * S.gamma) * qsel;
    float q_sa = qs[0];
#pragma unroll
    for (int a = 1; a < kQLD - 1; ++a) q_sa = (a == act) ? qs[a] : q_sa;
    const float delta = q_sa - y;
    const float atd = fabsf(y - q_sa);
    const float z = fabsf(delta);
    const float hub = (z < 1.f) ? (0.5f * z) * z : z - 0.5f;
    const float go = per ? (1.f / static_cast<float>(S.Bglobal)) * w : 1.f / static_cast<float>(S.Bglobal);
    const float g = fminf(fmaxf(delta, -1.f), 1.f) * go;
    lterm = per ? w * hub : hub;
    C.y[i] = y; C.q_sa[i] = q_sa; C.abs_td[i] = atd; C.hub[i] = hub; C.gcoef[i] = g;
    float dh[kQLD];
    if (N.dueling) {
      const float mean = g / static_cast<float>(N.A);
      dh[0] = g;
#pragma unroll
      for (int a = 0; a < kQLD - 1; ++a) dh[1 + a] = (a < N.A) ? ((a == act) ? g : 0.f) - mean : 0.f;
    } else {
#pragma unroll
      for (int a = 0; a < kQLD; ++a) dh[a] = (a < N.A && a == act) ? g : 0.f;
    }
#pragma unroll
    for (int a = 0; a < kQLD; ++a) {
      drec[i * N.rec + N.head_off + a] = dh[a];
      C.QT[i * kQLD + a] = qt[a]; C.QN[i * kQLD + a] = qn[a]; C.Q[i * kQLD + a] = qs[a]; C.DH[i * kQLD + a] = dh[a];
    }
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) lterm += __shfl_down_sync(0xffffffffu, lterm, sh);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = lterm;
  __syncthreads();
  if (threadIdx.x == 0) C.loss_part[blockIdx.x] = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
}

// heads backward: d(last)[i][k] = (sum_a dh[i][a] W_a[k]) * act'(h[i][k]);  dW_a[k] = sum_i dh[i][a] h[i][k];  db_a = sum_i dh[i][a]
__global__ void __launch_bounds__(256) k_hyb_heads_dgrad(HybNet N, const float* __restrict__ P, const float* __restrict__ rec_s, float* __restrict__ drec, long long B) {
  pdl_enter();
  const long long t = blockIdx.x * 256ll + threadIdx.x;
  if (t >= B * N.last_len) return;
  const long long i = t / N.last_len;
  const int k = static_cast<int>(t - i * N.last_len);
  const float* dh = drec + i * N.rec + N.head_off;
  float s = 0.f;
  for (int a = 0; a < N.NH; ++a) {
    const float* w = N.dueling ? (a == 0 ? P + N.hw_off[0] : P + N.hw_off[1] + (a - 1) * N.last_len) : P + N.hw_off[0] + a * N.last_len;
    s = fmaf(dh[a], __ldg(w + k), s);
  }
  drec[i * N.rec + N.last_off + k] = act_bwd(s, rec_s[i * N.rec + N.last_off + k], N.act);
}
__global__ void __launch_bounds__(256) k_hyb_heads_wgrad(HybNet N, const float* __restrict__ rec_s, const float* __restrict__ drec, long long B, float* __restrict__ grads) {
  pdl_enter();
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int a = t / N.last_len, k = t - a * N.last_len;
  if (a >= N.NH) return;
  float s = 0.f, sb = 0.f;
  for (long long i = 0; i < B; ++i) {
    const float d = drec[i * N.rec + N.head_off + a];
    s = fmaf(d, rec_s[i * N.rec + N.last_off + k], s);
    sb += d;
  }
  const int wo = N.dueling ? (a == 0 ? N.hw_off[0] : N.hw_off[1] + (a - 1) * N.last_len) : N.hw_off[0] + a * N.last_len;
  grads[wo + k] = s;
  if (k == 0) grads[N.dueling ? (a == 0 ? N.hb_off[0] : N.hb_off[1] + a - 1) : N.hb_off[0] + a] = sb;
}

// Adam (+ Polyak / hard sync) over the flat blob, and the loss of the step
__global__ void __launch_bounds__(256) k_hyb_adam(AgentCtx C, StepScalars S, int total, int n_loss_parts, int write_loss) {
  pdl_enter();
  const int pi = blockIdx.x * 256 + threadIdx.x;
  if (pi < total) {
    const float g = (S.grads_in != nullptr) ? __ldcg(S.grads_in + pi) : __ldcg(C.grads + pi);
    adam_polyak_element(C, S, pi, g);
  }
  if (write_loss && blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int c = 0; c < n_loss_parts; ++c) s += __ldcg(C.loss_part + c);
    const float loss = s / static_cast<float>(S.Bglobal);
    C.loss[0] = loss;
    host_loss_store(C.host_loss, loss, S.epoch);
  }
}

// inference outputs from the head records: mode 0 greedy actions, 1 Q values [n][A], 2 raw heads [n][NH]
__global__ void k_hyb_outputs(HybNet N, const float* __restrict__ rec_base, long long n, long long* __restrict__ actions, float* __restrict__ q_out, int mode) {
  pdl_enter();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float h[kQLD], q[kQLD];
#pragma unroll
  for (int a = 0; a < kQLD; ++a) h[a] = rec_base[i * N.rec + N.head_off + a];
  if (mode == 2) {
#pragma unroll
    for (int a = 0; a < kQLD; ++a)
      if (a < N.NH) q_out[i * N.NH + a] = h[a];
    return;
  }
  if (mode == 0) {       // dueling: argmax of RAW advantages (network.py:110-117); plain: argmax Q; first maximum wins
    int best = 0;
    float bv = N.dueling ? h[1] : h[0];
#pragma unroll
    for (int a = 1; a < kQLD - 1; ++a) {
      const float v = N.dueling ? h[a + 1] : h[a];
      if (a < N.A && v > bv) { bv = v; best = a; }
    }
    actions[i] = best;
    return;
  }
  hyb_heads_to_q(h, N.A, N.dueling, q);
#pragma unroll
  for (int a = 0; a < kQLD; ++a)
    if (a < N.A) q_out[i * N.A + a] = q[a];
}

}  // namespace rmc
