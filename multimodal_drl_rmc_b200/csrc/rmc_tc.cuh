// rmc_tc.cuh -- tensor-core (tcgen05 / TMEM) mode of the batched act kernel for the DENSE config
// (BASELINE configs[2]: 65,536 macro-state vectors).  bf16 operands, fp32 accumulation in tensor memory.
//
// This is the "stated looser bound" mode of the north star: Q values within 1e-2 max-norm-relative of the fp32
// path (bf16 operand rounding, SURVEY 7.2 estimated 2.5e-3), greedy actions equal except near-ties.  The exact
// fp32 FFMA kernel (k_mlp_infer) stays the default and the parity reference.
//
// One CTA (256 threads) owns a 128-row tile; the three layers are chained through tensor memory:
//   X[128x16]  . W0^T -> D1 (TMEM, 256 cols) -> +b0, ReLU, bf16 -> H1 (smem, UMMA canonical K-major layout)
//   H1[128x256]. W2^T -> D2 (TMEM, 128 cols) -> +b2, ReLU, bf16 -> H2 (smem)
//   H2[128x128]. Wh^T -> D3 (TMEM,  16 cols) -> +bh, argmax    -> actions
// and runs TWO such tile pipelines side by side (warps 0-3 / 4-7) so that one group's epilogue overlaps the other
// group's MMAs.
// tcgen05.mma is issued by one thread; accumulators are read back with tcgen05.ld (32 lanes x 32 columns per
// warp and instruction).  Operands use the no-swizzle canonical layout: 8x8-element core matrices of 128
// contiguous bytes, K-adjacent cores LBO = 128 B apart, 8-row groups SBO = (K/8)*128 B apart.
#pragma once
#include <cuda_bf16.h>

#include "rmc_device.cuh"

namespace rmc {

constexpr int kTcRows = 128;         // rows per tile = UMMA M
constexpr int kTcK1 = 16;            // layer-1 K (obs_dim padded to one UMMA_K step)
constexpr int kTcNH = 16;            // heads padded to the minimum N for M = 128
// packed bf16 blob (element offsets, bf16 units)
constexpr int kTcOffW0 = 0;                            // [256 rows][K = 16]
constexpr int kTcOffW2 = kTcOffW0 + kH1 * kTcK1;       // [128 rows][K = 256]
constexpr int kTcOffWh = kTcOffW2 + kH2 * kH1;         // [ 16 rows][K = 128]
// Biases ride on the tensor core (the hidden epilogues are instruction-issue bound; a bias add per element plus the bias
// loads were a third of their instructions): b0 sits in column 15 of the W0 image and the X tile carries 1.0 there
// (obs_dim <= 15), b2 is one more K step of layer 2 -- A = a constant [128][16] tile with 1.0 in column 0, B = [128][16]
// with b2 in column 0.  The biases enter as bf16 like every other operand of this mode (stated bound unchanged).
constexpr int kTcOffB2T = kTcOffWh + kTcNH * kH2;      // [128 rows = j][K = 16]: (j, 0) = b2[j]
constexpr int kTcOffOnes = kTcOffB2T + kH2 * kTcK1;    // [128 rows][K = 16]:     (r, 0) = 1
constexpr int kTcBf16Elems = kTcOffOnes + kTcRows * kTcK1;   // 43,008 bf16 = 86,016 B
constexpr int kTcBiasFloats = kH1 + kH2 + kTcNH;       // 400 floats (heads; hidden layers when obs_dim = 16 / fallback)
constexpr int kTcBlobBytes = kTcBf16Elems * 2 + kTcBiasFloats * 4;   // 87,616 B (multiple of 16)
__host__ __device__ __forceinline__ bool tc_bias_folded_l1(int D) { return D <= kTcK1 - 1; }

// element offset of (row r, k) inside a canonical K-major no-swizzle operand with K columns
__host__ __device__ __forceinline__ int tc_off(int r, int k, int K) {
  return ((r >> 3) * (K >> 3) + (k >> 3)) * 64 + (r & 7) * 8 + (k & 7);
}

// fp32 device-layout parameter blob -> packed bf16 operands + fp32 biases
__global__ void k_tc_pack(const float* __restrict__ blob, NetLayout L, unsigned char* __restrict__ out) {
  pdl_enter();
  __nv_bfloat16* w = reinterpret_cast<__nv_bfloat16*>(out);
  float* bias = reinterpret_cast<float*>(out + kTcBf16Elems * 2);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kH1 * kTcK1) {                                  // W0[i][d]
    const int i = t / kTcK1, d = t % kTcK1;
    const float v = d < L.D ? blob[L.off_w0t + d * kH1 + i] : ((d == kTcK1 - 1 && tc_bias_folded_l1(L.D)) ? blob[L.off_b0 + i] : 0.f);
    w[kTcOffW0 + tc_off(i, d, kTcK1)] = __float2bfloat16_rn(v);
  }
  if (t < kH2 * kTcK1) {                                  // bias tile of layer 2 and the constant ones tile
    const int j = t / kTcK1, k = t % kTcK1;
    w[kTcOffB2T + tc_off(j, k, kTcK1)] = __float2bfloat16_rn(k == 0 ? blob[L.off_b2 + j] : 0.f);
    w[kTcOffOnes + tc_off(j, k, kTcK1)] = __float2bfloat16_rn(k == 0 ? 1.f : 0.f);
  }
  if (t < kH2 * kH1) {                                    // W2[j][k]
    const int j = t / kH1, k = t % kH1;
    w[kTcOffW2 + tc_off(j, k, kH1)] = __float2bfloat16_rn(blob[L.off_w2t + k * kW2LD + j]);
  }
  if (t < kTcNH * kH2) {                                  // Wh[a][j]
    const int a = t / kH2, j = t % kH2;
    w[kTcOffWh + tc_off(a, j, kH2)] = __float2bfloat16_rn(a < L.NH ? blob[L.off_wh + a * kH2 + j] : 0.f);
  }
  if (t < kH1) bias[t] = blob[L.off_b0 + t];
  if (t < kH2) bias[kH1 + t] = blob[L.off_b2 + t];
  if (t < kTcNH) bias[kH1 + kH2 + t] = (t < L.NH) ? blob[L.off_bh + t] : 0.f;
}

// ---- tcgen05 wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t tc_smem_desc(const void* smem_ptr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3fff);                // start address        bits [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;     // leading byte offset  bits [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;     // stride byte offset   bits [32,46)
  d |= 1ull << 46;                                                  // descriptor version 1 (sm_100)
  return d;                                                         // layout_type = 0: no swizzle
}
__device__ __forceinline__ uint32_t tc_idesc_bf16(int M, int N) {
  return (1u << 4)                                  // accumulator format f32
         | (1u << 7) | (1u << 10)                   // A, B formats bf16
         | (static_cast<uint32_t>(N >> 3) << 17)    // N / 8
         | (static_cast<uint32_t>(M >> 4) << 24);   // M / 16 ; A and B K-major, no negate, dense
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits)
  return *reinterpret_cast<const uint32_t*>(&v);
}

// epilogue of one hidden layer: TMEM columns [0, 32*n32) of this warp's 32 lanes -> +bias, ReLU, bf16 ->
// canonical K-major operand `dst` (K = Kdst) for the next layer.  row = TMEM lane.
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {      // no wait: pair with tc_ld_wait()
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// wait for the outstanding tensor-memory loads; the registers are in/out operands so that no use of them can be scheduled
// above the wait
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// 32 accumulator columns of this thread's row -> +bias, ReLU, bf16 -> operand tile (+ global copy).
// The epilogue is instruction-issue bound (profiles/r1/tc_fwd_stages.txt), so it runs on packed operations: one
// add.f32x2 per two bias adds (sm_100 packed fp32), one cvt per two elements, and ReLU as one max.bf16x2 on the rounded
// pair -- max(round(x), 0) == round(max(x, 0)) because rounding is monotonic and keeps the sign.
// act = 1: ELU(alpha = 1), the repo-HEAD activation (env/dqn_config.py:175): z > 0 ? z : exp(z) - 1 in fp32 (ex2-based
// __expf: its 2-ulp error is far below the bf16 rounding that follows), then rounded to bf16 like the ReLU output.
template <int ACT>
__device__ __forceinline__ uint32_t tc_bias_relu_pack(uint32_t a0, uint32_t a1, float b0, float b1) {
  const float2 s = __fadd2_rn(make_float2(__uint_as_float(a0), __uint_as_float(a1)), make_float2(b0, b1));
  if (ACT != 0) {
    const float x = (s.x > 0.f) ? s.x : __expf(s.x) - 1.f, y = (s.y > 0.f) ? s.y : __expf(s.y) - 1.f;
    const __nv_bfloat162 r = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r = __hmax2(__floats2bfloat162_rn(s.x, s.y), __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <int ACT>
__device__ __forceinline__ void tc_hidden_chunk(const uint32_t (&v)[32], int row, int col, const float* __restrict__ bias,
                                                __nv_bfloat16* __restrict__ dst, int Kdst) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {              // 4 cores of 8 columns
    const float4 b0 = *reinterpret_cast<const float4*>(bias + col + 8 * c), b1 = *reinterpret_cast<const float4*>(bias + col + 8 * c + 4);
    uint4 q;
    q.x = tc_bias_relu_pack<ACT>(v[8 * c + 0], v[8 * c + 1], b0.x, b0.y);
    q.y = tc_bias_relu_pack<ACT>(v[8 * c + 2], v[8 * c + 3], b0.z, b0.w);
    q.z = tc_bias_relu_pack<ACT>(v[8 * c + 4], v[8 * c + 5], b1.x, b1.y);
    q.w = tc_bias_relu_pack<ACT>(v[8 * c + 6], v[8 * c + 7], b1.z, b1.w);
    *reinterpret_cast<uint4*>(dst + tc_off(row, col + 8 * c, Kdst)) = q;
  }
}
// The same with the bias already inside the accumulator (see kTcOffB2T): ReLU is the .relu modifier of the conversion, so
// a pair of elements costs ONE instruction.
__device__ __forceinline__ uint32_t tc_cvt_relu_bf16x2(uint32_t lo, uint32_t hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return d;
}
template <int ACT>
__device__ __forceinline__ uint32_t tc_act_pack(uint32_t a0, uint32_t a1) {
  if (ACT == 0) return tc_cvt_relu_bf16x2(a0, a1);
  const float z0 = __uint_as_float(a0), z1 = __uint_as_float(a1);
  return pack_bf16x2((z0 > 0.f) ? z0 : __expf(z0) - 1.f, (z1 > 0.f) ? z1 : __expf(z1) - 1.f);
}
template <int ACT>
__device__ __forceinline__ void tc_hidden_chunk_nb(const uint32_t (&v)[32], int row, int col, __nv_bfloat16* __restrict__ dst, int Kdst) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 q;
    q.x = tc_act_pack<ACT>(v[8 * c + 0], v[8 * c + 1]);
    q.y = tc_act_pack<ACT>(v[8 * c + 2], v[8 * c + 3]);
    q.z = tc_act_pack<ACT>(v[8 * c + 4], v[8 * c + 5]);
    q.w = tc_act_pack<ACT>(v[8 * c + 6], v[8 * c + 7]);
    *reinterpret_cast<uint4*>(dst + tc_off(row, col + 8 * c, Kdst)) = q;
  }
}
// chunks [c_first, c_first + n32) of 32 columns each; the tensor-memory load of the next chunk is in flight while the
// current one is converted (n32 is even)
template <int ACT, bool BIAS>
__device__ __forceinline__ void tc_hidden_epilogue_t(uint32_t tmem_acc, int lane_base, int row, int c_first, int n32,
                                                     const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst, int Kdst) {
  const uint32_t t0 = tmem_acc + (static_cast<uint32_t>(lane_base) << 16) + 32 * c_first;
  uint32_t va[32], vb[32];
  tc_ld32_issue(t0, va);
  for (int b = 0; b < n32; b += 2) {
    tc_ld_wait(va);
    tc_ld32_issue(t0 + 32 * (b + 1), vb);
    if (BIAS) tc_hidden_chunk<ACT>(va, row, 32 * (c_first + b), bias, dst, Kdst);
    else tc_hidden_chunk_nb<ACT>(va, row, 32 * (c_first + b), dst, Kdst);
    tc_ld_wait(vb);
    if (b + 2 < n32) tc_ld32_issue(t0 + 32 * (b + 2), va);
    if (BIAS) tc_hidden_chunk<ACT>(vb, row, 32 * (c_first + b + 1), bias, dst, Kdst);
    else tc_hidden_chunk_nb<ACT>(vb, row, 32 * (c_first + b + 1), dst, Kdst);
  }
}
// the activation (and whether the bias is still to be added) is a compile-time parameter of the instruction-issue bound
// epilogue loop: one uniform branch per call.  bias = nullptr: already inside the accumulator.
__device__ __forceinline__ void tc_hidden_epilogue(uint32_t tmem_acc, int lane_base, int row, int c_first, int n32,
                                                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst, int Kdst, int act) {
  if (bias == nullptr) {
    if (act != 0) tc_hidden_epilogue_t<1, false>(tmem_acc, lane_base, row, c_first, n32, bias, dst, Kdst);
    else tc_hidden_epilogue_t<0, false>(tmem_acc, lane_base, row, c_first, n32, bias, dst, Kdst);
  } else {
    if (act != 0) tc_hidden_epilogue_t<1, true>(tmem_acc, lane_base, row, c_first, n32, bias, dst, Kdst);
    else tc_hidden_epilogue_t<0, true>(tmem_acc, lane_base, row, c_first, n32, bias, dst, Kdst);
  }
}
constexpr int kTcFwdThreads = 512;   // two tile pipelines x 8 warps: each TMEM lane quadrant is drained by two warps (column halves)
__device__ __forceinline__ void tc_group_sync(int g) {   // the 256 threads of one tile pipeline
  asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory");
}

// Two independent tile pipelines per CTA (warps 0-3 and 4-7): while one group runs its epilogue on the CUDA
// cores, the other group's UMMAs occupy the tensor core.  Per group: 256 TMEM columns (D2 aliases D1[0:128),
// D3 aliases D1[128:144) -- each is written only after its predecessor has been drained) and one 64 KB operand
// buffer (H2 overwrites H1 once layer 2 has consumed it).
// mode 0: actions (dueling: argmax of raw advantages = head columns 1..A; plain: argmax of columns 0..A-1)
// mode 2: raw head outputs [n][NH] float (diagnostics / error measurement)
// mode 3: all 16 head columns [n][16] float (training: raw heads), plus the optional row-major bf16 copies of the
//         input tile / hidden activations that the backward kernel consumes: per 128-row tile the shared-memory
//         operand images themselves (UMMA canonical core layout tc_off(r, k, K), K = 16 / 256 / 128), written with
//         bulk stores -- the backward kernel bulk-loads them and uses them as K-major and MN-major operands as they are.
// Input rows are read at obs + i*row_stride + col_off (act: row_stride = D, col_off = 0; training: the gathered rows).
struct TcFwdExtra {
  long long row_stride;
  int col_off;
  __nv_bfloat16* Xb;
  __nv_bfloat16* H1b;
  __nv_bfloat16* H2b;
  long long* dbg;          // diagnostics: clock64() of CTA 0 / pipeline 0 at the stage boundaries of its first tiles ([tile][8])
  int act;                 // hidden activation: 0 = ReLU, 1 = ELU(alpha = 1)
};
__device__ __forceinline__ void tc_fwd_body(const unsigned char* __restrict__ packed, int D, int A, int NH, int dueling,
                                            const float* __restrict__ obs, long long n, long long* __restrict__ actions,
                                            float* __restrict__ heads_out, int mode, const TcFwdExtra& X, int bid, int nb) {
  extern __shared__ __align__(128) unsigned char tsm[];
  __nv_bfloat16* sWts = reinterpret_cast<__nv_bfloat16*>(tsm);                 // packed W0 | W2 | Wh
  float* sBias = reinterpret_cast<float*>(tsm + kTcBf16Elems * 2);             // b0 | b2 | bh
  __nv_bfloat16* sXall = reinterpret_cast<__nv_bfloat16*>(tsm + kTcBlobBytes);  // [2][128][16]
  __nv_bfloat16* sHall = sXall + 2 * kTcRows * kTcK1;                           // [2][128][256]  (H1, then H2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sHall + 2 * kTcRows * kH1);      // [0] weights, [1+3g ..] MMA stages of group g
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_tiles = (n + kTcRows - 1) / kTcRows;
  const bool dbg_k = X.dbg != nullptr && bid == 0 && tid == 0;      // diagnostics: kernel entry / weights landed / exit in slots 32..34
  if (dbg_k) X.dbg[32] = clock64();
  // Prologue, ordered so that its three latencies overlap: the weight image (one TMA bulk copy, issued by the thread that
  // initialised its barrier), the first tile's input rows (global loads into registers) and the tensor-memory allocation.
  if (tid == 0) {
    for (int b = 0; b < 7; ++b) mbar_init(bars + b, 1);
    fence_mbar_init();
    mbar_expect_tx(bars + 0, kTcBlobBytes);
    bulk_g2s(tsm, packed, kTcBlobBytes, bars + 0);
  }
  const int g = warp >> 3, q = warp & 3, half = (warp >> 2) & 1, gtid = tid & 255;
  const int row = 32 * q + lane;                  // tile row = TMEM lane
  // the NEXT tile's input row is fetched into registers while this tile runs (the global-load latency of the tiny X tile
  // would otherwise sit at the head of every tile's dependent MMA -> epilogue chain)
  float xr[16];
  // 8-byte loads when rows and columns allow it: a warp's rows are 56-128 bytes apart, so every load instruction touches
  // 15-32 cache lines and the 14 scalar loads of a row cost as many tag look-ups each
  const bool vec2 = ((D & 1) == 0) && ((X.row_stride & 1) == 0) && ((X.col_off & 1) == 0) && ((reinterpret_cast<uintptr_t>(obs) & 7) == 0);
  auto load_row = [&](long long t) {        // unconditional loads from a clamped (always valid) row: nothing depends on them until
    if (half != 0) return;                  // the next iteration packs them (invalid rows / columns are zeroed there)
    const long long i = min(t * kTcRows + gtid, n - 1);
    const float* src = obs + i * X.row_stride + X.col_off;
    if (vec2) {
#pragma unroll
      for (int d = 0; d < 16; d += 2)
        if (d < D) asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(xr[d]), "=f"(xr[d + 1]) : "l"(src + d));
    } else {
#pragma unroll
      for (int d = 0; d < 16; ++d)
        if (d < D) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(xr[d]) : "l"(src + d));
    }
  };
#pragma unroll
  for (int d = 0; d < 16; ++d) xr[d] = 0.f;
  load_row(bid + static_cast<long long>(g) * nb);
  if (warp == 0) {   // tensor-memory allocation: all 512 columns, one warp, then release the permit
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(bars + 0, 0);
  if (dbg_k) X.dbg[33] = clock64();

  __nv_bfloat16* sX = sXall + g * kTcRows * kTcK1;
  __nv_bfloat16* sH = sHall + g * kTcRows * kH1;
  uint64_t* gb = bars + 1 + 3 * g;
  const uint32_t tD1 = tmem + 256 * g, tD2 = tD1, tD3 = tD1 + 128;
  const uint32_t id1 = tc_idesc_bf16(kTcRows, kH1), id2 = tc_idesc_bf16(kTcRows, kH2), id3 = tc_idesc_bf16(kTcRows, kTcNH);
  uint32_t phase = 0;
  const bool save = X.H1b != nullptr;       // training pass over the s rows: keep X / H1 / H2 for the backward kernel
  const bool fold1 = tc_bias_folded_l1(D);  // b0 rides in column 15 of the layer-1 operands
  // Stagger: the two pipelines have identical stage lengths, so started together they stay in phase -- both in their
  // epilogues (CUDA cores contended), then both in their MMAs (tensor core contended, the layer-2 wait doubles).  Pipeline 1
  // starts its first tile when pipeline 0 has finished its first hidden epilogue; from then on one computes while the other
  // converts.
  bool first_tile = true;
  int tile_no = 0;
  if (g == 1 && bid + static_cast<long long>(nb) < n_tiles) asm volatile("bar.sync 3, 512;" ::: "memory");
  for (long long tile = bid + static_cast<long long>(g) * nb; tile < n_tiles; tile += 2ll * nb) {
    // ---- X tile: fp32 obs -> bf16 canonical [128][16]; thread = row
    if (half == 0) {
      const long long i = tile * kTcRows + gtid;
      if (i >= n) {
#pragma unroll
        for (int d = 0; d < 16; ++d) xr[d] = 0.f;
      }
      if (fold1) xr[15] = 1.f;              // bias column of layer 1 (W0 image column 15 = b0)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint4 v;
        v.x = pack_bf16x2(xr[8 * c + 0], xr[8 * c + 1]); v.y = pack_bf16x2(xr[8 * c + 2], xr[8 * c + 3]);
        v.z = pack_bf16x2(xr[8 * c + 4], xr[8 * c + 5]); v.w = pack_bf16x2(xr[8 * c + 6], xr[8 * c + 7]);
        *reinterpret_cast<uint4*>(sX + tc_off(gtid, 8 * c, kTcK1)) = v;
      }
    }
    load_row(tile + 2ll * nb);
    int dslot = 0;
    const bool dbg_on = X.dbg != nullptr && bid == 0 && gtid == 0 && g == 0 && tile_no < 4;      // (no 64-bit division here: it showed up as a 1,500-cycle "gap")
    long long* dbg_row = dbg_on ? X.dbg + tile_no * 8 : nullptr;
#define TC_FWD_STAMP() do { if (dbg_on) dbg_row[dslot++] = clock64(); } while (0)
    TC_FWD_STAMP();               // 0: X packed
    fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core's async proxy
    tc_fence_before();
    tc_group_sync(g);
    // ---- layer 1: D1[128x256] = X . W0^T  (one UMMA, K = 16)
    if (gtid == 0) {
      tc_fence_after();
      tc_mma_bf16(tD1, tc_smem_desc(sX, 128, (kTcK1 / 8) * 128), tc_smem_desc(sWts + kTcOffW0, 128, (kTcK1 / 8) * 128), id1, 0u);
      tc_commit(gb + 0);
    }
    mbar_wait(gb + 0, phase);
    tc_fence_after();
    TC_FWD_STAMP();               // 1: layer-1 MMA complete
    tc_hidden_epilogue(tD1, 32 * q, row, 4 * half, 4, fold1 ? nullptr : sBias, sH, kH1, X.act);
    TC_FWD_STAMP();               // 2: epilogue 1 done (this thread)
    if (first_tile && g == 0 && bid + static_cast<long long>(nb) < n_tiles) asm volatile("bar.arrive 3, 512;" ::: "memory");
    first_tile = false;
    fence_proxy_async();
    tc_fence_before();
    tc_group_sync(g);
    // ---- layer 2: D2[128x128] = H1 . W2^T  (K = 256: 16 UMMAs, descriptors advance by 2 cores = 256 B)
    if (gtid == 0) {
      tc_fence_after();
      const uint64_t a0 = tc_smem_desc(sH, 128, (kH1 / 8) * 128), b0 = tc_smem_desc(sWts + kTcOffW2, 128, (kH1 / 8) * 128);
#pragma unroll
      for (int k = 0; k < kH1 / 16; ++k) tc_mma_bf16(tD2, a0 + static_cast<uint64_t>(k * 16), b0 + static_cast<uint64_t>(k * 16), id2, k > 0 ? 1u : 0u);
      // + b2: ones tile x bias tile (one more K step)
      tc_mma_bf16(tD2, tc_smem_desc(sWts + kTcOffOnes, 128, (kTcK1 / 8) * 128), tc_smem_desc(sWts + kTcOffB2T, 128, (kTcK1 / 8) * 128), id2, 1u);
      tc_commit(gb + 1);
      if (save) {     // the operand images of X and H1 go to HBM as they are: two bulk stores, read by the TMA engine while layer 2 runs
        bulk_s2g(X.Xb + tile * (kTcRows * kTcK1), sX, kTcRows * kTcK1 * 2);
        bulk_s2g(X.H1b + tile * (kTcRows * kH1), sH, kTcRows * kH1 * 2);
        bulk_commit();
      }
    }
    mbar_wait(gb + 1, phase);
    tc_fence_after();
    TC_FWD_STAMP();               // 3: layer-2 MMAs complete
    if (save) {                   // H2 overwrites H1: the bulk store must have finished READING the buffer
      if (gtid == 0) bulk_wait_read();
      tc_group_sync(g);
    }
    tc_hidden_epilogue(tD2, 32 * q, row, 2 * half, 2, nullptr, sH, kH2, X.act);     // H2 overwrites H1 (layer 2 has consumed it)
    TC_FWD_STAMP();               // 4: epilogue 2 done
    fence_proxy_async();
    tc_fence_before();
    tc_group_sync(g);
    // ---- heads: D3[128x16] = H2 . Wh^T  (K = 128: 8 UMMAs)
    if (gtid == 0) {
      tc_fence_after();
      const uint64_t a0 = tc_smem_desc(sH, 128, (kH2 / 8) * 128), b0 = tc_smem_desc(sWts + kTcOffWh, 128, (kH2 / 8) * 128);
#pragma unroll
      for (int k = 0; k < kH2 / 16; ++k) tc_mma_bf16(tD3, a0 + static_cast<uint64_t>(k * 16), b0 + static_cast<uint64_t>(k * 16), id3, k > 0 ? 1u : 0u);
      tc_commit(gb + 2);
      if (save) {
        bulk_s2g(X.H2b + tile * (kTcRows * kH2), sH, kTcRows * kH2 * 2);
        bulk_commit();
      }
    }
    mbar_wait(gb + 2, phase);
    tc_fence_after();
    TC_FWD_STAMP();               // 5: head MMAs complete
    if (half == 0) {
      uint32_t v[16];
      tc_ld16(tD3 + (static_cast<uint32_t>(32 * q) << 16), v);
      const long long i = tile * kTcRows + row;
      if (i < n) {
        const float* bh = sBias + kH1 + kH2;
        if (mode == 0) {
          const int first = dueling ? 1 : 0;
          int best = 0;
          float bv = __uint_as_float(v[first]) + bh[first];
#pragma unroll
          for (int a = 1; a < 15; ++a) {
            if (a < A) {
              const float x = __uint_as_float(v[first + a]) + bh[first + a];
              if (x > bv) { bv = x; best = a; }
            }
          }
          actions[i] = best;
        } else if (mode == 3) {
#pragma unroll
          for (int a = 0; a < kTcNH; a += 4)
            *reinterpret_cast<float4*>(heads_out + i * kTcNH + a) =
                make_float4(__uint_as_float(v[a]) + bh[a], __uint_as_float(v[a + 1]) + bh[a + 1],
                            __uint_as_float(v[a + 2]) + bh[a + 2], __uint_as_float(v[a + 3]) + bh[a + 3]);
        } else {
#pragma unroll
          for (int a = 0; a < kTcNH; ++a)
            if (a < NH) heads_out[i * NH + a] = __uint_as_float(v[a]) + bh[a];
        }
      }
    }
    TC_FWD_STAMP();               // 6: head epilogue done
    if (save && gtid == 0) bulk_wait_read();      // sX / sH are rewritten by the next tile
    tc_fence_before();
    tc_group_sync(g);     // this group's TMEM slot and operand buffers are free for its next tile
    TC_FWD_STAMP();               // 7: tile done
    phase ^= 1u;
    ++tile_no;
  }
  if (save && gtid == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  if (dbg_k) X.dbg[34] = clock64();
}

__global__ void __launch_bounds__(kTcFwdThreads, 1) k_mlp_infer_tc(const unsigned char* __restrict__ packed, int D, int A, int NH, int dueling,
                                                              const float* __restrict__ obs, long long n,
                                                              long long* __restrict__ actions, float* __restrict__ heads_out, int mode,
                                                              TcFwdExtra X) {
  tc_fwd_body(packed, D, A, NH, dueling, obs, n, actions, heads_out, mode, X, blockIdx.x, gridDim.x);
}

// The three forward passes of a learner step (online(s'), target(s'), online(s) + saved activations) in ONE launch: each
// job owns a contiguous range of CTAs, which stage that job's weight image and walk that job's tiles.
struct TcFwdJob {
  const unsigned char* packed;
  float* heads_out;
  TcFwdExtra X;
  int cta_begin, cta_count;
};
struct TcFwdJobs { TcFwdJob j[3]; };
__global__ void __launch_bounds__(kTcFwdThreads, 1) k_tc_fwd3(TcFwdJobs J, int D, int A, int NH, int dueling, const float* __restrict__ rows, long long n) {
  pdl_enter();
  const SpanScope span_(SPAN_FWD3);
  const int b = blockIdx.x;
  const int k = (b >= J.j[2].cta_begin) ? 2 : (b >= J.j[1].cta_begin) ? 1 : 0;
  const TcFwdJob& job = J.j[k];
  tc_fwd_body(job.packed, D, A, NH, dueling, rows, n, nullptr, job.heads_out, 3, job.X, b - job.cta_begin, job.cta_count);
}

constexpr int kTcSmemBytes = kTcBlobBytes + 2 * (kTcRows * kTcK1 + kTcRows * kH1) * 2 + 8 * 8 + 16;

}  // namespace rmc
