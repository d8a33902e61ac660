// rmc_device.cuh -- device-side building blocks shared by the kernels of librmc_b200.
// sm_100a only.  Compiled with -fmad=false: every FMA in this library is an explicit fmaf()
// so that the parity-critical scalar arithmetic (sampling values, TD target, Adam, Polyak)
// rounds exactly like the reference's un-fused numpy / ATen-CPU expressions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rmc {

constexpr int kH1 = 256;        // hidden1 (macro network_config)
constexpr int kH2 = 128;        // hidden2
constexpr int kW2LD = 132;      // padded leading dim of W2^T rows: conflict-free for both fwd and dgrad
constexpr int kTM = 4;          // batch rows per row-tile (online pass runs 2*kTM rows: s' and s)
constexpr int kR = 2 * kTM;     // rows held per tile in shared memory
constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kMaxD = 32;
constexpr int kMaxRowFloats = 68;   // round4(2*32+3)
constexpr int kQLD = 16;        // leading dim of per-sample Q rows (A <= 15, heads <= 16)
constexpr int kTopNodes = 1023;  // tree[0..1023) = the top 10 levels, cached in shared memory by the sampler
constexpr int kTreeCtaMax = 4096;  // batches up to this size get their tree write-back from one CTA

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// TMA bulk copy global -> shared (1-D, no tensor map), completion signalled on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA bulk copy shared -> global (1-D), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be reused
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // writes complete
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Watchdog of every in-kernel spin (agent barrier, hand-off words): the spins assume that all CTAs of the launch are
// co-resident.  A cooperative launch guarantees it; the default programmatic-dependent launch relies on the host side
// (grid <= #SM, one CTA per SM, fused-step launches of one device serialised across streams -- launch_step in
// rmc_b200.cu).  If that assumption is ever broken (another process sharing the GPU through MPS, ...) the spin gives up
// after kSpinTimeoutNs, records the launch's epoch in the mapped-host error word and lets the kernel terminate; the host
// reports RMC_ERR_STATE at the next rmc_learner_loss_sync / rmc_learner_status.  The timer is read once per 1024 polls
// (> 30 us of waiting), i.e. never in a healthy step.
constexpr unsigned long long kSpinTimeoutNs = 2000000000ull;
struct SpinGuard {
  unsigned long long t0 = 0;
  unsigned polls = 0;
  __device__ __forceinline__ bool expired() {
    if ((++polls & 0x3ffu) != 0u) return false;
    const unsigned long long now = global_timer_ns();
    if (t0 == 0) { t0 = now; return false; }
    return now - t0 > kSpinTimeoutNs;
  }
};
// The loss of a step goes straight to mapped pinned host memory as ONE 8-byte word {loss bits, launch epoch}: a single
// aligned 8-byte store is one PCIe write, so the host sees the value and its epoch together and no system-scope fence is
// needed between them (that fence cost 1.5-4 us at the end of the publishing CTA / kernel).
__device__ __forceinline__ void host_loss_store(volatile float* host_loss, float loss, unsigned epoch) {
  if (host_loss != nullptr)
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(host_loss), "r"(__float_as_uint(loss)), "r"(epoch) : "memory");
}
// err: mapped pinned host memory of the agent (AgentCtx::host_loss), word [2] = epoch of a launch whose spin timed out
__device__ __forceinline__ void spin_report_timeout(volatile float* err, unsigned epoch) {
  if (err != nullptr) {
    err[2] = __uint_as_float(epoch);
    __threadfence_system();
  }
}

// Barrier across the CTAs of one agent inside ONE launch (all CTAs co-resident, see SpinGuard).
// `ctr` only ever grows; `target` = value it must reach (wrap-safe signed comparison).
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void agent_barrier(unsigned* ctr, unsigned target, volatile float* err, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_add_u32(ctr, 1u);      // release: orders the CTA's writes (observed through the barrier above) before the arrival
    SpinGuard guard;
    while (static_cast<int>(ld_acquire_u32(ctr) - target) < 0) {
      __nanosleep(32);
      if (guard.expired()) { spin_report_timeout(err, epoch); break; }
    }
    __threadfence();
  }
  __syncthreads();
}

// Kernel-span recorder (diagnostic, RMC_SPANS=1): every kernel of the multi-kernel pipelines notes its first CTA's start and
// last CTA's end (%globaltimer) in a device table, slot = kernel id.  Unlike ncu's serialised cold-cache replays and unlike
// CUDA events (which break programmatic dependent launches and cannot look inside a graph launch), this shows the real
// overlap of the kernels of a graph-launched step.  g_span_table == nullptr (default): one predictable branch per CTA.
__device__ unsigned long long* g_span_table = nullptr;      // [64][2] = {min start, max end}
__device__ __forceinline__ void span_begin(int slot) {
  if (g_span_table != nullptr && threadIdx.x == 0) atomicMin(g_span_table + 2 * slot, global_timer_ns());
}
__device__ __forceinline__ void span_end(int slot) {
  if (g_span_table != nullptr && threadIdx.x == 0) atomicMax(g_span_table + 2 * slot + 1, global_timer_ns());
}
struct SpanScope {      // span of a kernel whose CTAs return from several places
  int slot;
  __device__ __forceinline__ explicit SpanScope(int s) : slot(s) { span_begin(s); }
  __device__ __forceinline__ ~SpanScope() { span_end(slot); }
};
enum SpanSlot { SPAN_SAMPLE = 0, SPAN_FWD3, SPAN_TD, SPAN_BWD, SPAN_REDUCE_ADAM, SPAN_TD_TO_PRI, SPAN_TREE_STAMP, SPAN_TREE_APPLY, SPAN_TREE_TOP,
                SPAN_EXTREMES, SPAN_PUBLISH_TD, SPAN_GATHER_TD, SPAN_PUBLISH, SPAN_COMM_REDUCE, SPAN_TREE_SMALL, SPAN_UNIFORM, SPAN_PACK, SPAN_COUNT };

// first statement of every kernel launched through launch_pdl(): let the next launch be scheduled early, then wait until
// the preceding grid has completed and its writes are visible (no-ops for ordinary launches)
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

constexpr int kDbgSlots = 32;        // timestamps per CTA in AgentCtx::dbg
constexpr int kDbgCtas = 1024;       // CTA records; the launch-gap slots follow them
#define RMC_STAMP(C, slot)                                                                         \
  do {                                                                                             \
    if ((C).dbg != nullptr && threadIdx.x == 0)                                                    \
      (C).dbg[(blockIdx.y * gridDim.x + blockIdx.x) * kDbgSlots + (slot)] = global_timer_ns();     \
  } while (0)

// ----------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
// uniform double in [0,1) with 53 random bits
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t counter, uint32_t agent, uint32_t i) {
  uint4 c = make_uint4(i, agent, static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32));
  uint2 k = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  uint4 r = philox4x32_10(c, k);
  uint64_t bits = ((static_cast<uint64_t>(r.x) << 32) | r.y) >> 11;
  return static_cast<double>(bits) * (1.0 / 9007199254740992.0);
}
// Keyed bijection on [0, n): balanced Feistel over the next even power of two + cycle walking.
// Gives `batch` DISTINCT positions for i = 0..batch-1 (sampling without replacement).
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint64_t feistel_perm(uint64_t i, uint64_t n, uint64_t seed, uint64_t counter, uint32_t agent) {
  int bits = 2;
  while ((1ull << bits) < n) bits += 2;   // even number of bits
  const int half = bits / 2;
  const uint32_t mask = (1u << half) - 1u;
  const uint32_t k0 = mix32(static_cast<uint32_t>(seed) ^ 0x9E3779B9u) ^ mix32(static_cast<uint32_t>(counter) + agent * 0x85ebca6bu);
  const uint32_t k1 = mix32(static_cast<uint32_t>(seed >> 32) + 0x632BE5ABu) ^ mix32(static_cast<uint32_t>(counter >> 32) ^ 0xc2b2ae35u);
  uint64_t x = i;
  do {
    uint32_t l = static_cast<uint32_t>(x >> half) & mask, r = static_cast<uint32_t>(x) & mask;
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      uint32_t f = mix32(r ^ (rd & 1 ? k1 : k0) ^ (0x27d4eb2fu * (rd + 1))) & mask;
      uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    x = (static_cast<uint64_t>(l) << half) | r;
  } while (x >= n);
  return x;
}

// ----------------------------------------------------------------------------- data structs
struct ReplayState {   // lives in HBM, updated by the kernels themselves
  long long size;
  long long dp;
  float max_p;       // max(leaves[:size])  (0 when empty)
  float min_p;       // min(leaves[:size])  (+inf when empty)
  long long cnt_max; // number of leaves equal to max_p
  long long cnt_min; // number of leaves equal to min_p
  float push_p;      // priority given to the rows of the current (chunked) push call
  int bad_nodes;     // entries of external index lists outside the leaf range that a write-back skipped (sticky count)
};

constexpr int kFlagWords = 16384;   // u32 words of AgentCtx::qt_flag per agent (layout: rmc_mlp.cuh)
constexpr int kTreeTeam = 8;     // CTAs that share the priority write-back of one learner step
constexpr int kStreamTilesMax = 74;  // row tiles of a launch that runs the streamed phase B (one tile per row CTA, role split)
struct TeamPart {                // one tree-team member's contribution to the extremes
  float bmax, bmin;              // extremes of the NEW values it applied
  int cnt_bmax, cnt_bmin;        // how many of its new values equal its own bmax / bmin
  int old_eq_max, old_eq_min;    // how many overwritten (in-domain) values equalled the previous global max / min
  int pad[2];
};

struct TinyRows { float v[8 * kMaxRowFloats]; };   // up to 8 packed rows carried in the kernel-argument buffer

struct ReplayDev {
  float* ring;       // [cap][row_floats]
  double* tree;      // [2*cap-1]  reference heap layout (dqn/utils/sum_tree.py:6-13)
  int* stamps;       // [cap]   last-writer election for duplicate leaves in a batch
  float* scratch_old;// [kTreeCtaMax] overwritten leaf values of the batch being applied
  TeamPart* team_part;   // [kTreeTeam]
  unsigned* team_ctr;    // arrival counter of the tree team (returns to 0 after every step)
  ReplayState* st;
  long long cap;
  int row_floats;
  int obs_dim;
  int prioritized;
  int pad;
};

struct NetLayout {     // float offsets inside one parameter blob (device layout)
  int D, A, NH, dueling;
  int off_w0t, off_b0, off_w2t, off_b2, off_wh, off_bh, total;   // total is a multiple of 4
  int act;           // hidden activation: 0 = ReLU, 1 = ELU(alpha = 1)
};

// Hidden activation and its derivative expressed through the OUTPUT h (what the step keeps):
//   ReLU: h = max(z, 0),               dh/dz = [h > 0]
//   ELU : h = z > 0 ? z : exp(z) - 1   (torch's CPU kernel: exp then subtract, not expm1),  dh/dz = h > 0 ? 1 : h + 1 (= exp(z))
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == 0) return fmaxf(z, 0.f);
  return (z > 0.f) ? z : expf(z) - 1.f;
}
__device__ __forceinline__ float act_bwd(float upstream, float h, int act) {
  if (act == 0) return (h > 0.f) ? upstream : 0.f;
  return (h > 0.f) ? upstream : upstream * (h + 1.f);
}

struct AgentCtx {
  ReplayDev rp;
  NetLayout L;
  float *online, *target, *adam_m, *adam_v, *grads;
  // per-sample products of the last step
  long long* nodes;     // [B] tree node (PER) or ring slot (uniform)
  double* leaf_p;       // [B] priority of the sampled leaf (its tree value at sampling time)
  float *is_w, *q_sa, *y, *abs_td, *hub, *pri, *gcoef;
  float *QT, *QN, *Q;   // [B][kQLD]  Q_target(s'), Q_online(s'), Q_online(s)
  float *X;             // [B][row_floats] gathered rows
  float *H1, *H2, *DZ1, *DZ2, *DH;   // saved activations / deltas of the s rows
  float* loss_part;     // [n_tiles]
  float* gpart;         // [gpart_cap][L.total] per-CTA partial gradient blobs of the batch-stationary row phase (rmc_rows_ws.cuh); zero at padding
  float* loss;          // [1]
  unsigned* barrier;
  unsigned* qt_flag;    // [kFlagWords] {epoch, payload} hand-off words of the fused step: Q_target(s') per (tile, row, action),
                        // |td| per row and the loss partial per tile (layout: rmc_mlp.cuh)
  // streamed phase B of the fused step (rmc_mlp.cuh): what the weight-gradient units need from the row CTAs, as
  // {epoch, value} words that are polled -- no barrier and no fence between the row phase and the weight gradients
  unsigned* x_words;    // [rows][kMaxD]  state columns of the s rows (zero past obs_dim)
  unsigned* hp_words;   // [rows][kH1]    hidden-1 activations
  unsigned* h2_words;   // [rows][kH2]    hidden-2 activations
  unsigned* dh_words;   // [rows][kQLD]   head deltas
  unsigned* zp_words;   // [rows][kH2]    layer-2 deltas dz2
  unsigned* z1_words;   // [rows][kH1]    layer-1 deltas dz1
  volatile float* host_loss;      // mapped pinned host memory: [0] loss of the last step, [1] its epoch (as bits), [2] epoch of a launch whose in-kernel spin timed out (0 = healthy)
  unsigned long long* dbg;   // optional per-CTA phase timestamps [G][16] (nullptr = off)
};

struct StepScalars {
  long long B;            // local batch
  long long Bglobal;      // batch of the whole job (loss / gradient scale, stratified segments)
  long long shard_off;    // global index of local sample 0
  int phases;
  int double_dqn;
  int prioritized;        // learner flavour uses IS weights + write-back
  int n_row_ctas;         // CTAs that own row tiles
  unsigned barrier_target;  // counter value after the (single) A->B barrier
  unsigned epoch;           // unique per launch of this agent / group (never 0)
  double beta;
  const double* u;        // injected uniforms (agent-major) or nullptr
  const long long* idx;   // injected positions or nullptr
  unsigned long long seed, counter;
  const float* grads_in;
  float gamma;
  float adam_w1;          // (float)(1 - beta1)
  float adam_b2;          // (float)beta2
  float adam_w2;          // (float)(1 - beta2)
  float adam_neg_step;    // (float)(-(lr / (1 - beta1^t)))
  float adam_bc2_sqrt;    // (float)sqrt(1 - beta2^t)
  float adam_eps;
  float polyak_k, polyak_1mk;
  float per_eps, per_alpha, per_pmax;
};

}  // namespace rmc
