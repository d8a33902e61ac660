// rmc_b200.cu -- host side of librmc_b200.so: the C ABI declared in include/rmc_b200.h.
// Plain CUDA runtime, no torch.  One translation unit with the kernels (rmc_tree.cuh, rmc_mlp.cuh).
#include "../../include/rmc_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <tuple>
#include <vector>

#include "rmc_device.cuh"
#include "rmc_mlp.cuh"
#include "rmc_tc.cuh"
#include "rmc_tc_train.cuh"
#include "rmc_comm.cuh"
#include "rmc_infer64.cuh"
#include "rmc_hybrid.cuh"
#include "rmc_tree.cuh"

using namespace rmc;

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int32_t fail(int32_t code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define RMC_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(RMC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
  } while (0)
#define RMC_KERNEL_OK()                                                                                  \
  do {                                                                                                   \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                                  \
    cudaError_t _e = cudaGetLastError();                                                                 \
    if (_e != cudaSuccess) return fail(RMC_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
  } while (0)


// Programmatic dependent launch for the multi-kernel pipelines (tensor-core step, hybrid network, grid-wide sampler / tree
// write-back, peer exchange): every kernel of these chains begins with pdl_enter() (griddepcontrol.launch_dependents +
// griddepcontrol.wait), so the NEXT kernel's launch latency overlaps this kernel's execution while its body still runs
// strictly after this kernel has completed and flushed.  RMC_PDL=0 launches them as ordinary stream-ordered kernels.
// Measured (B200): tensor-core step 173 -> 167 us; for the hybrid network's many multi-wave kernels it is neutral at
// B = 32 and SLOWER at B = 256 (1.63 -> 2.01 ms: early-resident dependent CTAs take SM slots from the running grid), so
// the hybrid paths switch it off with a PdlScope.
static thread_local bool g_pdl_on = true;
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(g_pdl_on) { g_pdl_on = on; }
  ~PdlScope() { g_pdl_on = prev; }
};
// Re-trace-and-patch CUDA graphs for launch-bound multi-kernel steps (the hybrid network: ~47 small kernels on two
// streams per step).  Every kernel of such a step goes through launch_pdl(), so one host routine serves three modes:
//   * ordinary        launch the kernel;
//   * capturing (1)   launch it into a capturing stream and remember the node it became, with a copy of its arguments;
//   * patching  (2)   do not launch: compare the arguments with the node's copy and, where they differ (step counters,
//                     Adam bias corrections, sampling seeds, injected pointers), update that node of the instantiated
//                     graph.  The caller then replays the whole step with ONE cudaGraphLaunch.
// A kernel / grid that does not match the recorded sequence marks the trace invalid; the caller falls back to ordinary
// launches and captures again.
struct TraceRec {
  cudaGraphNode_t node = nullptr;
  const void* func = nullptr;
  dim3 grid, block;
  size_t smem = 0;
  std::vector<unsigned char> bytes;      // the kernel arguments, back to back
};
struct StepGraph {
  long long batch = 0; int phases = 0;
  cudaGraph_t graph = nullptr;           // kept alive: the node handles of `recs` belong to it
  cudaGraphExec_t exec = nullptr;
  std::vector<TraceRec> recs;
  void destroy() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    exec = nullptr; graph = nullptr;
  }
};
struct Trace { int mode = 0; StepGraph* g = nullptr; size_t cursor = 0; bool ok = true; int patched = 0; };
static thread_local Trace* g_trace = nullptr;

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool enabled = [] { const char* e = std::getenv("RMC_PDL"); return !(e && e[0] == '0'); }();
  std::tuple<std::decay_t<KArgs>...> packed(static_cast<KArgs>(args)...);
  void* ptrs[sizeof...(KArgs) + 1];
  size_t total = 0;
  std::apply([&](auto&... a) { size_t i = 0; ((ptrs[i++] = static_cast<void*>(&a), total += sizeof(a)), ...); }, packed);
  auto pack_bytes = [&](std::vector<unsigned char>& out) {
    out.resize(total);
    size_t off = 0;
    std::apply([&](auto&... a) { ((std::memcpy(out.data() + off, &a, sizeof(a)), off += sizeof(a)), ...); }, packed);
  };
  if (g_trace != nullptr && g_trace->mode == 2) {
    Trace& T = *g_trace;
    if (!T.ok) return cudaSuccess;
    if (T.cursor >= T.g->recs.size()) { T.ok = false; return cudaSuccess; }
    TraceRec& R = T.g->recs[T.cursor++];
    if (R.func != reinterpret_cast<const void*>(kernel) || R.grid.x != grid.x || R.grid.y != grid.y || R.grid.z != grid.z || R.block.x != block.x ||
        R.smem != smem || R.bytes.size() != total) { T.ok = false; return cudaSuccess; }
    bool same = true;
    {
      size_t off = 0;
      std::apply([&](auto&... a) { ((same = same && std::memcmp(R.bytes.data() + off, &a, sizeof(a)) == 0, off += sizeof(a)), ...); }, packed);
    }
    if (same) return cudaSuccess;
    cudaKernelNodeParams np{};
    np.func = const_cast<void*>(R.func); np.gridDim = grid; np.blockDim = block; np.sharedMemBytes = static_cast<unsigned>(smem);
    np.kernelParams = ptrs; np.extra = nullptr;
    const cudaError_t e = cudaGraphExecKernelNodeSetParams(T.g->exec, R.node, &np);
    if (e != cudaSuccess) { T.ok = false; cudaGetLastError(); return cudaSuccess; }
    pack_bytes(R.bytes);
    ++T.patched;
    return cudaSuccess;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (enabled && g_pdl_on) ? 1 : 0;
  const cudaError_t err = cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), ptrs);
  if (err == cudaSuccess && g_trace != nullptr && g_trace->mode == 1 && g_trace->ok) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    const cudaGraphNode_t* deps = nullptr;
    size_t n_deps = 0;
    if (cudaStreamGetCaptureInfo_v2(st, &status, nullptr, nullptr, &deps, &n_deps) != cudaSuccess || status != cudaStreamCaptureStatusActive ||
        n_deps != 1) {
      g_trace->ok = false;
      cudaGetLastError();
    } else {
      TraceRec R;
      R.node = deps[0]; R.func = reinterpret_cast<const void*>(kernel); R.grid = grid; R.block = block; R.smem = smem;
      pack_bytes(R.bytes);
      g_trace->g->recs.push_back(std::move(R));
    }
  }
  return err;
}

static inline cudaStream_t as_stream(rmc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }
static inline int round4(int x) { return (x + 3) & ~3; }

// ------------------------------------------------------------------------------ handles
static constexpr int kStageSlots = 4;

struct rmc_replay {
  int device = 0;
  long long cap = 0, size = 0, dp = 0;   // host mirrors of the device state
  int D = 0, rf = 0, prioritized = 0;
  long long n_nodes = 0;
  ReplayDev dev{};
  long long* scratch_nodes = nullptr;
  float* scratch_pri = nullptr;
  ExtTuple* ext_parts = nullptr;     // per-block partials + arrival counter of the one-pass extremes rescan
  unsigned* ext_arrive = nullptr;
  // staging ring for host pushes
  long long stage_rows = 0;
  float* pin[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  float* dstage[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  bool ev_used[kStageSlots] = {false, false, false, false};
  int slot = 0;
  bool l2_window_set = false;
};

struct rmc_learner {
  int device = 0;
  rmc_net_spec_t spec{};
  rmc_hyper_t hyper{};
  NetLayout L{};
  long long P = 0;           // torch parameter count
  long long max_batch = 0, last_batch = 0;
  int rf = 0;
  int num_sms = 0;
  int smem_bytes = 0;
  int max_smem_optin = 0;
  bool big_ready = false;
  float* blobs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // online,target,m,v,grads
  int* map = nullptr;        // torch index -> device-layout index
  float* io = nullptr;       // [P] device staging for set/get
  AgentCtx ctx{};            // replay part filled per step
  unsigned barrier_count = 0;
  unsigned epoch = 0;
  unsigned loss_epoch = 0;          // epoch of the last launch that produced a loss
  volatile float* host_loss = nullptr;   // mapped pinned host memory (host view)
  unsigned long long* dbg_buf = nullptr;
  unsigned char* tc_packed = nullptr;   // bf16 operands of the tensor-core act mode (lazily allocated)
  unsigned long long online_version = 1, tc_packed_version = 0;   // repack only when the online weights changed
  int last_grid = 0;
  // tensor-core training mode (lazily allocated, rmc_tc_train.cuh)
  bool tct_ready = false;
  unsigned char* tc_packed_target = nullptr;
  __nv_bfloat16* tc_packed_bwd = nullptr;
  unsigned long long tc_bwd_version = 0;
  unsigned long long target_version = 1, tc_target_version = 0;
  TcTrainBufs tct{};
  cudaStream_t tc_side = nullptr;           // the PER write-back runs here beside the backward / Adam kernels
  cudaEvent_t tc_ev[2] = {};
  bool no_graph = false;                    // set around the local part of a sharded step: the caller owns the graph
  struct { rmc_comm* c = nullptr; CommView V{}; int parity = 0; bool issue_side = false; long long Bg = 0; } early;   // sharded step: (leaf,|td|) leave right after TD
  // hybrid CNN+MLP network (rmc_hybrid.cuh): per-row activation / delta records
  bool hybrid = false;
  HybNet H{};
  float *rec_on = nullptr, *rec_tg = nullptr, *drec = nullptr, *hyb_ws = nullptr, *hyb_ws2 = nullptr;
  cudaStream_t hyb_side = nullptr;          // second stream: target-network pass and the weight-gradient kernels
  cudaStream_t hyb_cap = nullptr;           // origin stream of the step-graph captures (the caller's stream may be the legacy default stream)
  std::vector<StepGraph> hyb_graphs;        // instantiated step graphs, one per (batch, phases)
  bool hyb_graph_off = false;               // a capture failed on this handle: ordinary launches from then on
  cudaEvent_t hyb_ev[16] = {};
  // act staging
  float* act_pin_obs = nullptr; long long* act_pin_out = nullptr; float* act_dev_obs = nullptr; long long* act_dev_out = nullptr;
  long long act_cap = 0;
  // per-env-step act (k_act_tiny): mapped pinned host words [0..kActTinyMax) actions, [kActTinyMax] epoch; device view; arrival counter
  volatile long long* act_map_host = nullptr; long long* act_map_dev = nullptr; unsigned* act_ctr = nullptr; unsigned act_epoch = 0;
  int act_map_state = 0;   // 0 = not tried, 1 = ready, -1 = unavailable (falls back to the copy path)
  std::vector<void*> owned;
};

struct rmc_comm {
  int device = 0, rank = 0, world = 1;
  long long global_batch_max = 0, local_max = 0;
  size_t bytes = 0;
  unsigned char* local = nullptr;               // this rank's exchange buffer (cudaMalloc: IPC-exportable)
  void* peer[kCommMaxWorld] = {nullptr};        // mapped peer buffers (own rank: == local)
  bool ipc_opened[kCommMaxWorld] = {false};
  bool connected = false;
  CommView view{};
  unsigned epoch = 0;
  unsigned* arrive = nullptr;
  unsigned* arrive_td = nullptr;
  unsigned* verdict = nullptr;                  // [2] grid-uniform outcome of the waits (comm_wait_all)
  long long* g_nodes = nullptr;                 // gathered (leaf, |td|) of the whole batch + their priorities
  float* g_td = nullptr;
  float* g_pri = nullptr;
  rmc_learner* learner = nullptr;
};

struct rmc_group {
  int n = 0;
  std::vector<rmc_learner*> learners;
  std::vector<rmc_replay*> replays;
  AgentCtx* ctx_dev = nullptr;
  unsigned* barriers = nullptr;
  unsigned* qt_flags = nullptr;
  unsigned barrier_count = 0;
  unsigned epoch = 0;
  // tensor-core mode: the members' steps run side by side on one stream per member (fork / join around them)
  std::vector<cudaStream_t> tc_streams;
  std::vector<cudaEvent_t> tc_join;
  cudaEvent_t tc_fork = nullptr;
};

static int32_t step_resident_ctas(int device, int smem_bytes, int* out);
extern "C" int32_t rmc_replay_destroy(rmc_replay_t* r);
extern "C" int32_t rmc_learner_destroy(rmc_learner_t* l);
extern "C" int32_t rmc_comm_destroy(rmc_comm_t* c);

// ------------------------------------------------------------------------------ library
extern "C" int32_t rmc_abi_version(void) { return RMC_ABI_VERSION; }
extern "C" const char* rmc_last_error(void) { return g_err.c_str(); }
extern "C" int64_t rmc_launch_count(void) { return g_launches.load(); }

static int32_t use_device(int device) {
  RMC_CUDA(cudaSetDevice(device));
  return RMC_OK;
}

template <typename T>
static int32_t dev_alloc(T** p, size_t count, bool zero = true) {
  RMC_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  if (zero) RMC_CUDA(cudaMemset(*p, 0, count * sizeof(T)));
  return RMC_OK;
}

__global__ void k_fill_f32(float* p, long long n, float v) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_init_state(ReplayState* st) {
  st->size = 0; st->dp = 0; st->max_p = 0.f; st->min_p = __int_as_float(0x7f800000); st->cnt_max = 0; st->cnt_min = 0;
  st->push_p = 1.f; st->bad_nodes = 0;
}

// ------------------------------------------------------------------------------ replay
extern "C" int32_t rmc_replay_create(rmc_replay_t** out, int64_t capacity, int32_t obs_dim, int32_t prioritized,
                                     int32_t device) {
  if (!out || capacity < 1 || obs_dim < 1 || obs_dim > 4096) return fail(RMC_ERR_ARG, "rmc_replay_create: bad capacity/obs_dim");
  if (capacity > (1ll << 30)) return fail(RMC_ERR_ARG, "rmc_replay_create: capacity too large");
  if (int32_t e = use_device(device)) return e;
  auto* r = new rmc_replay();
  r->device = device;
  r->cap = capacity;
  r->D = obs_dim;
  r->rf = round4(2 * obs_dim + 3);
  r->prioritized = prioritized ? 1 : 0;
  r->n_nodes = 2 * capacity - 1;
  ReplayDev& d = r->dev;
  d.cap = capacity; d.row_floats = r->rf; d.obs_dim = obs_dim; d.prioritized = r->prioritized;
  auto build = [&]() -> int32_t {
    int32_t e = RMC_OK;
    if ((e = dev_alloc(&d.ring, static_cast<size_t>(capacity) * r->rf))) return e;
    if ((e = dev_alloc(&d.st, 1))) return e;
    k_init_state<<<1, 1>>>(d.st);
    RMC_KERNEL_OK();
    if (r->prioritized) {
      if ((e = dev_alloc(&d.tree, static_cast<size_t>(r->n_nodes)))) return e;
      if ((e = dev_alloc(&d.stamps, static_cast<size_t>(capacity)))) return e;
      if ((e = dev_alloc(&d.scratch_old, kTreeCtaMax))) return e;
      if ((e = dev_alloc(&d.team_part, kTreeTeam))) return e;
      if ((e = dev_alloc(&d.team_ctr, 1))) return e;
    }
    if ((e = dev_alloc(&r->scratch_nodes, kTreeCtaMax))) return e;
    if ((e = dev_alloc(&r->scratch_pri, kTreeCtaMax))) return e;
    if ((e = dev_alloc(&r->ext_parts, kExtBlocks))) return e;
    if ((e = dev_alloc(&r->ext_arrive, 1))) return e;
    r->stage_rows = 16384;
    for (int s = 0; s < kStageSlots; ++s) {
      RMC_CUDA(cudaMallocHost(reinterpret_cast<void**>(&r->pin[s]), static_cast<size_t>(r->stage_rows) * r->rf * sizeof(float)));
      if ((e = dev_alloc(&r->dstage[s], static_cast<size_t>(r->stage_rows) * r->rf, false))) return e;
      RMC_CUDA(cudaEventCreateWithFlags(&r->ev[s], cudaEventDisableTiming));
    }
    RMC_CUDA(cudaDeviceSynchronize());
    return RMC_OK;
  };
  if (int32_t e = build()) {          // a partially built handle is released by its own destroy function
    const std::string msg = g_err;
    rmc_replay_destroy(r);
    return fail(e, msg);
  }
  *out = r;
  return RMC_OK;
}

extern "C" int32_t rmc_replay_destroy(rmc_replay_t* r) {
  if (!r) return RMC_OK;
  cudaSetDevice(r->device);
  cudaDeviceSynchronize();
  ReplayDev& d = r->dev;
  cudaFree(d.ring); cudaFree(d.tree); cudaFree(d.stamps); cudaFree(d.scratch_old); cudaFree(d.team_part); cudaFree(d.team_ctr); cudaFree(d.st); cudaFree(r->scratch_nodes); cudaFree(r->scratch_pri); cudaFree(r->ext_parts); cudaFree(r->ext_arrive);
  for (int s = 0; s < kStageSlots; ++s) {
    if (r->pin[s]) cudaFreeHost(r->pin[s]);
    cudaFree(r->dstage[s]);
    if (r->ev[s]) cudaEventDestroy(r->ev[s]);
  }
  delete r;
  return RMC_OK;
}

extern "C" int32_t rmc_replay_row_floats(const rmc_replay_t* r) { return r ? r->rf : 0; }

static int32_t tree_rebuild(rmc_replay* r, cudaStream_t st) {
  // internal nodes are 0 .. cap-2; rebuild level by level from the deepest
  const long long last_internal = r->cap - 2;
  for (int L = 40; L >= 0; --L) {
    const long long first = (1ll << L) - 1;
    if (first > last_internal) continue;
    const long long last = std::min((1ll << (L + 1)) - 2, last_internal);
    const long long count = last - first + 1;
    k_tree_rebuild_level<<<blocks_for(count, 256), 256, 0, st>>>(r->dev.tree, first, count);
    RMC_KERNEL_OK();
  }
  return RMC_OK;
}
static int32_t minmax_rebuild(rmc_replay* r, cudaStream_t st) {
  const unsigned grid = std::max(1u, std::min(static_cast<unsigned>(kExtBlocks), blocks_for(r->cap, 1024)));
  RMC_CUDA(launch_pdl(k_extremes_scan, dim3(grid), dim3(256), 0, st, r->dev, r->ext_parts, r->ext_arrive));
  RMC_KERNEL_OK();
  return RMC_OK;
}

// rows already packed in a device staging buffer
static int32_t push_packed_small(rmc_replay* r, const float* rows_dev, long long n, cudaStream_t st) {
  if (n <= 32 && n <= r->cap) {
    k_push_tiny<false><<<1, 32, 0, st>>>(r->dev, rows_dev, TinyRows{}, static_cast<int>(n), 1.0f);
  } else {
    k_push_small<<<1, kThreads, 0, st>>>(r->dev, rows_dev, n, r->scratch_nodes, r->scratch_pri, 1.0f);
  }
  RMC_KERNEL_OK();
  r->dp = (r->dp + n) % r->cap;
  r->size = std::min(r->size + n, r->cap);
  return RMC_OK;
}

static void pack_rows_host(float* dst, const float* obs, const int64_t* act, const float* rew, const float* done,
                           const float* nxt, long long n, int D, int rf) {
  for (long long i = 0; i < n; ++i) {
    float* row = dst + i * rf;
    std::memcpy(row, obs + i * D, sizeof(float) * D);
    std::memcpy(row + D, nxt + i * D, sizeof(float) * D);
    const int32_t a = static_cast<int32_t>(act[i]);
    std::memcpy(row + 2 * D, &a, sizeof(float));
    row[2 * D + 1] = rew[i];
    row[2 * D + 2] = done[i];
    for (int c = 2 * D + 3; c < rf; ++c) row[c] = 0.f;
  }
}

static int32_t push_impl(rmc_replay* r, const float* obs, const int64_t* act, const float* rew, const float* done,
                         const float* nxt, int64_t n, bool host, cudaStream_t st) {
  if (!r || n < 0) return fail(RMC_ERR_ARG, "rmc_replay_push: bad args");
  if (n == 0) return RMC_OK;
  if (int32_t e = use_device(r->device)) return e;
  if (host && n <= 8 && n <= r->cap && r->rf <= kMaxRowFloats) {
    // the trainer's per-env-step push: the packed rows ride in the kernel-argument buffer of the launch (no
    // staging copy, no event) -- that launch IS the host->device transfer of these n*row_floats*4 bytes
    TinyRows tr;
    pack_rows_host(tr.v, obs, act, rew, done, nxt, n, r->D, r->rf);
    k_push_tiny<true><<<1, 32, 0, st>>>(r->dev, nullptr, tr, static_cast<int>(n), 1.0f);
    RMC_KERNEL_OK();
    r->dp = (r->dp + n) % r->cap;
    r->size = std::min<long long>(r->size + n, r->cap);
    return RMC_OK;
  }
  const long long small_max = std::min<long long>(kTreeCtaMax, r->cap);
  const bool bulk = n > small_max;
  const long long chunk_max = bulk ? std::min<long long>(r->stage_rows, r->cap) : small_max;
  if (bulk) {
    k_push_begin<<<1, 1, 0, st>>>(r->dev, 1.0f);
    RMC_KERNEL_OK();
  }
  for (long long off = 0; off < n; off += chunk_max) {
    const long long m = std::min<long long>(chunk_max, n - off);
    const int s = r->slot;
    r->slot = (r->slot + 1) % kStageSlots;
    if (host) {
      if (r->ev_used[s]) RMC_CUDA(cudaEventSynchronize(r->ev[s]));
      pack_rows_host(r->pin[s], obs + off * r->D, act + off, rew + off, done + off, nxt + off * r->D, m, r->D, r->rf);
      RMC_CUDA(cudaMemcpyAsync(r->dstage[s], r->pin[s], static_cast<size_t>(m) * r->rf * sizeof(float), cudaMemcpyHostToDevice, st));
      RMC_CUDA(cudaEventRecord(r->ev[s], st));
      r->ev_used[s] = true;
    } else {
      k_pack_rows<<<blocks_for(m * r->rf, 256), 256, 0, st>>>(r->dstage[s], obs + off * r->D, reinterpret_cast<const long long*>(act + off), rew + off, done + off,
                                                             nxt + off * r->D, m, r->D, r->rf);
      RMC_KERNEL_OK();
    }
    if (!bulk) {
      if (int32_t e = push_packed_small(r, r->dstage[s], m, st)) return e;
    } else {
      k_push_rows_bulk<<<blocks_for(m * r->rf, 256), 256, 0, st>>>(r->dev, r->dstage[s], m, r->dp);
      RMC_KERNEL_OK();
      r->dp = (r->dp + m) % r->cap;
      r->size = std::min(r->size + m, r->cap);
    }
  }
  if (bulk) {
    k_push_end<<<1, 1, 0, st>>>(r->dev, r->dp, r->size);
    RMC_KERNEL_OK();
    if (r->prioritized) {
      if (int32_t e = tree_rebuild(r, st)) return e;
      if (int32_t e = minmax_rebuild(r, st)) return e;
    }
  }
  return RMC_OK;
}

extern "C" int32_t rmc_replay_push(rmc_replay_t* r, const float* obs_dev, const int64_t* act_dev, const float* rew_dev,
                                   const float* done_dev, const float* next_obs_dev, int64_t n, rmc_stream_t s) {
  return push_impl(r, obs_dev, act_dev, rew_dev, done_dev, next_obs_dev, n, false, as_stream(s));
}
extern "C" int32_t rmc_replay_push_host(rmc_replay_t* r, const float* obs_host, const int64_t* act_host, const float* rew_host,
                                        const float* done_host, const float* next_obs_host, int64_t n, rmc_stream_t s) {
  return push_impl(r, obs_host, act_host, rew_host, done_host, next_obs_host, n, true, as_stream(s));
}

extern "C" int32_t rmc_replay_set_priorities(rmc_replay_t* r, const float* pri_dev, int64_t n, rmc_stream_t s) {
  if (!r || !r->prioritized || n < 0 || n > r->size) return fail(RMC_ERR_ARG, "rmc_replay_set_priorities: bad args");
  if (int32_t e = use_device(r->device)) return e;
  cudaStream_t st = as_stream(s);
  if (n > 0) {
    k_set_leaves<<<blocks_for(n, 256), 256, 0, st>>>(r->dev, pri_dev, n);
    RMC_KERNEL_OK();
  }
  if (int32_t e = tree_rebuild(r, st)) return e;
  return minmax_rebuild(r, st);
}

// Exact-resume side-car: put back a replay that rmc_replay_read_rows_sync / read_tree_sync saved.  Rows go to
// slots [0, size) through the pinned staging ring, the float32-exact leaf priorities to the leaves, then the inner nodes and
// the extremes are rebuilt (sums of float32-exact values are exact in float64, so the rebuilt tree equals the saved one bit
// for bit) and (size, data_pointer) are set -- the next sample / push behaves as if the process had never stopped.
__global__ void k_load_rows(ReplayDev R, const float* __restrict__ rows, long long first_slot, long long n) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t < n * R.row_floats) R.ring[first_slot * R.row_floats + t] = rows[t];
}
__global__ void k_load_leaves(ReplayDev R, const float* __restrict__ pri, long long first_slot, long long n) {
  const long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t < n) R.tree[R.cap - 1 + first_slot + t] = static_cast<double>(pri[t]);
}
extern "C" int32_t rmc_replay_load_host(rmc_replay_t* r, const float* rows_host, const float* leaf_pri_host, int64_t size, int64_t data_pointer,
                                        rmc_stream_t s) {
  if (!r || size < 0 || size > r->cap || data_pointer < 0 || data_pointer >= r->cap || (size > 0 && !rows_host))
    return fail(RMC_ERR_ARG, "rmc_replay_load_host: bad size / data_pointer");
  if (size < r->cap && data_pointer != size % r->cap) return fail(RMC_ERR_ARG, "rmc_replay_load_host: a ring that has not wrapped has data_pointer == size");
  if (r->prioritized && size > 0 && !leaf_pri_host) return fail(RMC_ERR_ARG, "rmc_replay_load_host: prioritized replay needs the leaf priorities");
  if (int32_t e = use_device(r->device)) return e;
  cudaStream_t st = as_stream(s);
  for (long long off = 0; off < size; off += r->stage_rows) {
    const long long m = std::min<long long>(r->stage_rows, size - off);
    const int sl = r->slot;
    r->slot = (r->slot + 1) % kStageSlots;
    if (r->ev_used[sl]) RMC_CUDA(cudaEventSynchronize(r->ev[sl]));
    std::memcpy(r->pin[sl], rows_host + off * r->rf, static_cast<size_t>(m) * r->rf * sizeof(float));
    RMC_CUDA(cudaMemcpyAsync(r->dstage[sl], r->pin[sl], static_cast<size_t>(m) * r->rf * sizeof(float), cudaMemcpyHostToDevice, st));
    k_load_rows<<<blocks_for(m * r->rf, 256), 256, 0, st>>>(r->dev, r->dstage[sl], off, m);
    RMC_KERNEL_OK();
    RMC_CUDA(cudaEventRecord(r->ev[sl], st));
    r->ev_used[sl] = true;
  }
  if (r->prioritized) {
    RMC_CUDA(cudaMemsetAsync(r->dev.tree, 0, static_cast<size_t>(r->n_nodes) * sizeof(double), st));
    const long long per = r->stage_rows * r->rf;      // floats per staging slot
    for (long long off = 0; off < size; off += per) {
      const long long m = std::min<long long>(per, size - off);
      const int sl = r->slot;
      r->slot = (r->slot + 1) % kStageSlots;
      if (r->ev_used[sl]) RMC_CUDA(cudaEventSynchronize(r->ev[sl]));
      std::memcpy(r->pin[sl], leaf_pri_host + off, static_cast<size_t>(m) * sizeof(float));
      RMC_CUDA(cudaMemcpyAsync(r->dstage[sl], r->pin[sl], static_cast<size_t>(m) * sizeof(float), cudaMemcpyHostToDevice, st));
      k_load_leaves<<<blocks_for(m, 256), 256, 0, st>>>(r->dev, r->dstage[sl], off, m);
      RMC_KERNEL_OK();
      RMC_CUDA(cudaEventRecord(r->ev[sl], st));
      r->ev_used[sl] = true;
    }
  }
  k_push_end<<<1, 1, 0, st>>>(r->dev, data_pointer, size);
  RMC_KERNEL_OK();
  r->dp = data_pointer;
  r->size = size;
  if (r->prioritized) {
    if (int32_t e = tree_rebuild(r, st)) return e;
    if (int32_t e = minmax_rebuild(r, st)) return e;
  }
  RMC_CUDA(cudaStreamSynchronize(st));
  return RMC_OK;
}

extern "C" int32_t rmc_replay_stats_sync(rmc_replay_t* r, rmc_replay_stats_t* out, rmc_stream_t s) {
  if (!r || !out) return fail(RMC_ERR_ARG, "rmc_replay_stats_sync: null");
  if (int32_t e = use_device(r->device)) return e;
  cudaStream_t st = as_stream(s);
  ReplayState hs{};
  double total = 0.0;
  RMC_CUDA(cudaMemcpyAsync(&hs, r->dev.st, sizeof(hs), cudaMemcpyDeviceToHost, st));
  if (r->prioritized) RMC_CUDA(cudaMemcpyAsync(&total, r->dev.tree, sizeof(double), cudaMemcpyDeviceToHost, st));
  RMC_CUDA(cudaStreamSynchronize(st));
  out->capacity = r->cap;
  out->size = hs.size;
  out->data_pointer = hs.dp;
  out->total_priority = total;
  out->max_priority = (hs.size > 0 && r->prioritized) ? static_cast<double>(hs.max_p) : 0.0;
  out->min_priority = (hs.size > 0 && r->prioritized) ? static_cast<double>(hs.min_p) : 0.0;
  out->rejected_nodes = hs.bad_nodes;
  return RMC_OK;
}

extern "C" int32_t rmc_replay_read_tree_sync(rmc_replay_t* r, double* out_host, int64_t first, int64_t n, rmc_stream_t s) {
  if (!r || !r->prioritized || first < 0 || n < 0 || first + n > r->n_nodes) return fail(RMC_ERR_ARG, "rmc_replay_read_tree_sync: range");
  if (int32_t e = use_device(r->device)) return e;
  RMC_CUDA(cudaMemcpyAsync(out_host, r->dev.tree + first, static_cast<size_t>(n) * sizeof(double), cudaMemcpyDeviceToHost, as_stream(s)));
  RMC_CUDA(cudaStreamSynchronize(as_stream(s)));
  return RMC_OK;
}
extern "C" int32_t rmc_replay_read_rows_sync(rmc_replay_t* r, float* out_host, int64_t first_slot, int64_t n, rmc_stream_t s) {
  if (!r || first_slot < 0 || n < 0 || first_slot + n > r->cap) return fail(RMC_ERR_ARG, "rmc_replay_read_rows_sync: range");
  if (int32_t e = use_device(r->device)) return e;
  RMC_CUDA(cudaMemcpyAsync(out_host, r->dev.ring + first_slot * r->rf, static_cast<size_t>(n) * r->rf * sizeof(float),
                           cudaMemcpyDeviceToHost, as_stream(s)));
  RMC_CUDA(cudaStreamSynchronize(as_stream(s)));
  return RMC_OK;
}

// PER minibatch draw by the grid-wide samplers: warp per sample (lowest latency) below kLaneSampleMin samples, lane per
// sample (whole batch in flight) above.
static constexpr long long kLaneSampleMin = 2048;
static int32_t launch_per_sample(const ReplayDev& R, long long B, long long Bglobal, long long shard_off, double beta, const double* u,
                                 unsigned long long seed, unsigned long long counter, long long* nodes, float* is_w, float* rows,
                                 double* leaf_p, cudaStream_t st) {
  if (B >= kLaneSampleMin)
    RMC_CUDA(launch_pdl(k_per_sample_lane, dim3(blocks_for(B, kLaneThreads)), dim3(kLaneThreads), 0, st, R, B, Bglobal, shard_off, beta, u, seed, counter, 0u, nodes, is_w, rows, leaf_p));
  else
    RMC_CUDA(launch_pdl(k_per_sample, dim3(blocks_for(B, kWarps)), dim3(kThreads), 0, st, R, B, Bglobal, shard_off, beta, u, seed, counter, 0u, nodes, is_w, rows, leaf_p));
  RMC_KERNEL_OK();
  return RMC_OK;
}

extern "C" int32_t rmc_per_sample(rmc_replay_t* r, int64_t batch, double beta, const double* u_dev, uint64_t seed, uint64_t counter,
                                  int64_t* out_nodes_dev, float* out_is_w_dev, float* out_rows_dev, rmc_stream_t s) {
  if (!r || !r->prioritized || batch < 1 || !out_nodes_dev) return fail(RMC_ERR_ARG, "rmc_per_sample: bad args");
  if (r->size < 1) return fail(RMC_ERR_STATE, "rmc_per_sample: empty replay");
  if (int32_t e = use_device(r->device)) return e;
  return launch_per_sample(r->dev, batch, batch, 0, beta, u_dev, seed, counter, reinterpret_cast<long long*>(out_nodes_dev), out_is_w_dev,
                           out_rows_dev, nullptr, as_stream(s));
}

extern "C" int32_t rmc_tree_get_leaf(rmc_replay_t* r, const double* v_dev, int64_t n, int64_t* out_nodes_dev, double* out_pri_dev,
                                     rmc_stream_t s) {
  if (!r || !r->prioritized || n < 1 || !v_dev || !out_nodes_dev) return fail(RMC_ERR_ARG, "rmc_tree_get_leaf: bad args");
  if (int32_t e = use_device(r->device)) return e;
  k_tree_get_leaf<<<blocks_for(n, kWarps), kThreads, 0, as_stream(s)>>>(r->dev, v_dev, n, reinterpret_cast<long long*>(out_nodes_dev), out_pri_dev);
  RMC_KERNEL_OK();
  return RMC_OK;
}

extern "C" int32_t rmc_uniform_sample(rmc_replay_t* r, int64_t batch, const int64_t* idx_dev, uint64_t seed, uint64_t counter,
                                      int64_t* out_slots_dev, float* out_rows_dev, rmc_stream_t s) {
  if (!r || batch < 1 || !out_slots_dev) return fail(RMC_ERR_ARG, "rmc_uniform_sample: bad args");
  if (r->size < batch) return fail(RMC_ERR_STATE, "rmc_uniform_sample: sample larger than population");
  if (int32_t e = use_device(r->device)) return e;
  RMC_CUDA(launch_pdl(k_uniform_sample, dim3(blocks_for(batch, kWarps)), dim3(kThreads), 0, as_stream(s), r->dev, batch, 0, reinterpret_cast<const long long*>(idx_dev),
                                                                            seed, counter, 0u,
                                                                            reinterpret_cast<long long*>(out_slots_dev), out_rows_dev));
  RMC_KERNEL_OK();
  return RMC_OK;
}

static int32_t tree_update_large(rmc_replay* r, const long long* nodes, const float* pri, long long n, bool stamps_done, cudaStream_t st) {
  if (!stamps_done) {
    RMC_CUDA(launch_pdl(k_tree_stamp, dim3(blocks_for(n, 256)), dim3(256), 0, st, r->dev, nodes, n));
    RMC_KERNEL_OK();
  }
  // ancestors below the top levels by float64 reductions (few updates per node); the contended top is rebuilt by one CTA
  int L = 0;
  while (L < 11 && (2ll << L) <= r->cap) ++L;             // largest L <= 11 with 2^L <= cap
  const int F = (L >= 3 && n >= 1024) ? (1 << L) - 1 : 0; // small batches / tiny trees: plain propagation to the root
  RMC_CUDA(launch_pdl(k_tree_apply, dim3(blocks_for(n, 256)), dim3(256), 0, st, r->dev, nodes, pri, n, static_cast<long long>(F)));
  RMC_KERNEL_OK();
  if (F > 0) {
    RMC_CUDA(launch_pdl(k_tree_rebuild_top, dim3(1), dim3(1024), 0, st, r->dev, F));
    RMC_KERNEL_OK();
  }
  return minmax_rebuild(r, st);
}

extern "C" int32_t rmc_per_update_from_td(rmc_replay_t* r, const int64_t* nodes_dev, const float* abs_td_dev, int64_t batch, float eps,
                                          float alpha, float pmax, float* out_pri_dev, rmc_stream_t s) {
  if (!r || !r->prioritized || batch < 1 || !nodes_dev || !abs_td_dev) return fail(RMC_ERR_ARG, "rmc_per_update_from_td: bad args");
  if (int32_t e = use_device(r->device)) return e;
  cudaStream_t st = as_stream(s);
  const long long* nodes = reinterpret_cast<const long long*>(nodes_dev);
  if (batch <= kTreeCtaMax) {
    float* pri = out_pri_dev ? out_pri_dev : r->scratch_pri;
    RMC_CUDA(launch_pdl(k_tree_update_small, dim3(1), dim3(kThreads), 0, st, r->dev, nodes, nullptr, abs_td_dev, pri, batch, eps, alpha, pmax));
    RMC_KERNEL_OK();
    return RMC_OK;
  }
  if (!out_pri_dev) return fail(RMC_ERR_ARG, "rmc_per_update_from_td: out_pri_dev required for batch > 4096");
  RMC_CUDA(launch_pdl(k_td_to_pri, dim3(blocks_for(batch, 256)), dim3(256), 0, st, abs_td_dev, out_pri_dev, batch, eps, alpha, pmax));
  RMC_KERNEL_OK();
  return tree_update_large(r, nodes, out_pri_dev, batch, false, st);
}

extern "C" int32_t rmc_per_update(rmc_replay_t* r, const int64_t* nodes_dev, const float* pri_dev, int64_t batch, rmc_stream_t s) {
  if (!r || !r->prioritized || batch < 1 || !nodes_dev || !pri_dev) return fail(RMC_ERR_ARG, "rmc_per_update: bad args");
  if (int32_t e = use_device(r->device)) return e;
  cudaStream_t st = as_stream(s);
  const long long* nodes = reinterpret_cast<const long long*>(nodes_dev);
  if (batch <= kTreeCtaMax) {
    RMC_CUDA(launch_pdl(k_tree_update_small, dim3(1), dim3(kThreads), 0, st, r->dev, nodes, pri_dev, nullptr, nullptr, batch, 0.f, 0.f, 0.f));
    RMC_KERNEL_OK();
    return RMC_OK;
  }
  return tree_update_large(r, nodes, pri_dev, batch, false, st);
}

// ------------------------------------------------------------------------------ learner
static NetLayout make_layout(const rmc_net_spec_t& sp) {
  NetLayout L{};
  L.D = sp.obs_dim; L.A = sp.n_actions; L.dueling = sp.dueling ? 1 : 0;
  L.NH = L.dueling ? L.A + 1 : L.A;
  L.act = (sp.activation == RMC_ACT_ELU) ? 1 : 0;
  int o = 0;
  L.off_w0t = o; o += L.D * kH1;
  L.off_b0 = o; o += kH1;
  L.off_w2t = o; o += kH1 * kW2LD;
  L.off_b2 = o; o += kH2;
  L.off_wh = o; o += L.NH * kH2;
  L.off_bh = o; o += round4(L.NH);
  L.total = o;
  return L;
}

static std::vector<int> make_param_map(const NetLayout& L) {
  std::vector<int> m;
  for (int i = 0; i < kH1; ++i) for (int d = 0; d < L.D; ++d) m.push_back(L.off_w0t + d * kH1 + i);   // net.0.weight [H1][D]
  for (int i = 0; i < kH1; ++i) m.push_back(L.off_b0 + i);                                              // net.0.bias
  for (int j = 0; j < kH2; ++j) for (int k = 0; k < kH1; ++k) m.push_back(L.off_w2t + k * kW2LD + j);  // net.2.weight [H2][H1]
  for (int j = 0; j < kH2; ++j) m.push_back(L.off_b2 + j);                                              // net.2.bias
  if (L.dueling) {
    for (int j = 0; j < kH2; ++j) m.push_back(L.off_wh + j);                                            // fc_val.weight [1][H2]
    m.push_back(L.off_bh);                                                                              // fc_val.bias
    for (int a = 0; a < L.A; ++a) for (int j = 0; j < kH2; ++j) m.push_back(L.off_wh + (1 + a) * kH2 + j);   // fc_adv.weight
    for (int a = 0; a < L.A; ++a) m.push_back(L.off_bh + 1 + a);                                        // fc_adv.bias
  } else {
    for (int a = 0; a < L.A; ++a) for (int j = 0; j < kH2; ++j) m.push_back(L.off_wh + a * kH2 + j);   // fc_out.weight
    for (int a = 0; a < L.A; ++a) m.push_back(L.off_bh + a);                                            // fc_out.bias
  }
  return m;
}

template <typename T>
static int32_t owned_alloc(rmc_learner* l, T** p, size_t count) {
  if (int32_t e = dev_alloc(p, count)) return e;
  l->owned.push_back(*p);
  return RMC_OK;
}

static int32_t learner_build_mlp(rmc_learner* l, const rmc_net_spec_t* spec, const rmc_hyper_t* hyper, int64_t max_batch, int32_t device) {
  l->device = device;
  l->spec = *spec;
  l->hyper = *hyper;
  l->L = make_layout(*spec);
  l->max_batch = max_batch;
  l->rf = round4(2 * spec->obs_dim + 3);
  RMC_CUDA(cudaDeviceGetAttribute(&l->num_sms, cudaDevAttrMultiProcessorCount, device));
  const SmemPlan plan = make_smem_plan(l->L.total);
  l->smem_bytes = std::max(plan.total_floats, kGemmSmemFloats) * 4;
  int max_optin = 0;
  RMC_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  l->max_smem_optin = max_optin;
  if (l->smem_bytes > max_optin) return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create: parameter blob does not fit shared memory");
  RMC_CUDA(cudaFuncSetAttribute(k_learner_step<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, l->smem_bytes));
  RMC_CUDA(cudaFuncSetAttribute(k_learner_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, l->smem_bytes));
  RMC_CUDA(cudaFuncSetAttribute(k_learner_step<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, l->smem_bytes));
  RMC_CUDA(cudaFuncSetAttribute(k_mlp_infer, cudaFuncAttributeMaxDynamicSharedMemorySize, l->smem_bytes));
  {
    int resident = 0;
    if (int32_t e2 = step_resident_ctas(device, l->smem_bytes, &resident)) return e2;
    if (resident < l->num_sms) return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create: the fused step's grid (one CTA per SM) would not be co-resident on this device");
  }
  const std::vector<int> map = make_param_map(l->L);
  l->P = static_cast<long long>(map.size());
  int32_t e = RMC_OK;
  for (int k = 0; k < 5; ++k)
    if ((e = owned_alloc(l, &l->blobs[k], static_cast<size_t>(l->L.total)))) return e;
  if ((e = owned_alloc(l, &l->map, map.size()))) return e;
  RMC_CUDA(cudaMemcpy(l->map, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice));
  if ((e = owned_alloc(l, &l->io, map.size()))) return e;
  AgentCtx& c = l->ctx;
  c.L = l->L;
  c.online = l->blobs[0]; c.target = l->blobs[1]; c.adam_m = l->blobs[2]; c.adam_v = l->blobs[3]; c.grads = l->blobs[4];
  const size_t B = static_cast<size_t>(max_batch);
  if ((e = owned_alloc(l, &c.nodes, B))) return e;
  if ((e = owned_alloc(l, &c.leaf_p, B))) return e;
  float** per_sample[] = {&c.is_w, &c.q_sa, &c.y, &c.abs_td, &c.hub, &c.pri, &c.gcoef};
  for (float** p : per_sample)
    if ((e = owned_alloc(l, p, B))) return e;
  if ((e = owned_alloc(l, &c.QT, B * kQLD))) return e;
  if ((e = owned_alloc(l, &c.QN, B * kQLD))) return e;
  if ((e = owned_alloc(l, &c.Q, B * kQLD))) return e;
  if ((e = owned_alloc(l, &c.X, B * l->rf))) return e;
  if ((e = owned_alloc(l, &c.H1, B * kH1))) return e;
  if ((e = owned_alloc(l, &c.DZ1, B * kH1))) return e;
  if ((e = owned_alloc(l, &c.H2, B * kH2))) return e;
  if ((e = owned_alloc(l, &c.DZ2, B * kH2))) return e;
  if ((e = owned_alloc(l, &c.DH, B * kQLD))) return e;
  if ((e = owned_alloc(l, &c.loss_part, 1024))) return e;
  if ((e = owned_alloc(l, &c.loss, 1))) return e;
  if ((e = owned_alloc(l, &c.barrier, 1))) return e;
  if ((e = owned_alloc(l, &c.qt_flag, kFlagWords))) return e;
  {   // per-CTA partial gradient blobs of the batch-stationary row phase (rmc_rows_ws.cuh): one per row CTA, tiles of >= 8 rows
    const size_t parts = std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(l->num_sms), (B + 7) / 8));
    if ((e = owned_alloc(l, &c.gpart, parts * static_cast<size_t>(l->L.total)))) return e;
  }
  {   // streamed phase B of the fused step: {epoch, value} word arrays (rmc_mlp.cuh, StreamPlan)
    const size_t rows = std::min<size_t>(B, static_cast<size_t>(kStreamTilesMax) * kTM);
    if ((e = owned_alloc(l, &c.x_words, 2 * rows * kMaxD))) return e;
    if ((e = owned_alloc(l, &c.hp_words, 2 * rows * kH1))) return e;
    if ((e = owned_alloc(l, &c.h2_words, 2 * rows * kH2))) return e;
    if ((e = owned_alloc(l, &c.dh_words, 2 * rows * kQLD))) return e;
    if ((e = owned_alloc(l, &c.zp_words, 2 * rows * kH2))) return e;
    if ((e = owned_alloc(l, &c.z1_words, 2 * rows * kH1))) return e;
  }
  {
    float* hp = nullptr;
    float* dp = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&hp), 64, cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&dp), hp, 0) == cudaSuccess) {
      for (int k = 0; k < 16; ++k) hp[k] = 0.f;
      l->host_loss = hp;
      c.host_loss = dp;
    } else {
      cudaGetLastError();
      c.host_loss = nullptr;
    }
  }
  if ((e = owned_alloc(l, &l->dbg_buf, kDbgCtas * kDbgSlots + 128))) return e;
  RMC_CUDA(cudaDeviceSynchronize());
  return RMC_OK;
}


extern "C" int32_t rmc_learner_create(rmc_learner_t** out, const rmc_net_spec_t* spec, const rmc_hyper_t* hyper, int64_t max_batch,
                                      int32_t device) {
  if (!out || !spec || !hyper || max_batch < 1) return fail(RMC_ERR_ARG, "rmc_learner_create: null/bad args");
  if (spec->hidden1 != kH1 || spec->hidden2 != kH2)
    return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create: only the macro MLP body 256-128 is built (no fallback path)");
  if (spec->activation != RMC_ACT_RELU && spec->activation != RMC_ACT_ELU)
    return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create: hidden activation must be ReLU or ELU(alpha=1)");
  if (spec->obs_dim < 1 || spec->obs_dim > kMaxD || spec->n_actions < 1 || spec->n_actions > 15)
    return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create: obs_dim must be 1..32 and n_actions 1..15");
  if (int32_t e = use_device(device)) return e;
  auto* l = new rmc_learner();
  if (int32_t e = learner_build_mlp(l, spec, hyper, max_batch, device)) {     // a partially built handle is released by its destroy function
    const std::string msg = g_err;
    rmc_learner_destroy(l);
    return fail(e, msg);
  }
  *out = l;
  return RMC_OK;
}

extern "C" int32_t rmc_learner_destroy(rmc_learner_t* l) {
  if (!l) return RMC_OK;
  cudaSetDevice(l->device);
  cudaDeviceSynchronize();
  for (void* p : l->owned) cudaFree(p);
  if (l->act_pin_obs) cudaFreeHost(l->act_pin_obs);
  if (l->act_pin_out) cudaFreeHost(l->act_pin_out);
  if (l->host_loss) cudaFreeHost(const_cast<float*>(l->host_loss));
  if (l->act_map_host) cudaFreeHost(const_cast<long long*>(l->act_map_host));
  if (l->act_ctr) cudaFree(l->act_ctr);
  if (l->tc_side) { cudaStreamDestroy(l->tc_side); for (auto& ev : l->tc_ev) if (ev) cudaEventDestroy(ev); }
  if (l->hyb_side) { cudaStreamDestroy(l->hyb_side); for (auto& ev : l->hyb_ev) if (ev) cudaEventDestroy(ev); }
  for (auto& g : l->hyb_graphs) g.destroy();
  if (l->hyb_cap) cudaStreamDestroy(l->hyb_cap);
  cudaFree(l->act_dev_obs);
  cudaFree(l->act_dev_out);
  delete l;
  return RMC_OK;
}

extern "C" int64_t rmc_learner_param_count(const rmc_learner_t* l) { return l ? l->P : 0; }

extern "C" int32_t rmc_learner_set_hyper(rmc_learner_t* l, const rmc_hyper_t* hyper) {
  if (!l || !hyper) return fail(RMC_ERR_ARG, "rmc_learner_set_hyper: null");
  l->hyper = *hyper;
  return RMC_OK;
}

extern "C" int32_t rmc_learner_set_params(rmc_learner_t* l, int32_t kind, const float* src, int64_t n, int32_t src_is_host, rmc_stream_t s) {
  if (!l || kind < 0 || kind > 4 || !src || n != l->P) return fail(RMC_ERR_ARG, "rmc_learner_set_params: bad kind/size");
  if (int32_t e = use_device(l->device)) return e;
  cudaStream_t st = as_stream(s);
  const float* dsrc = src;
  if (src_is_host) {
    RMC_CUDA(cudaMemcpyAsync(l->io, src, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice, st));
    dsrc = l->io;
  }
  k_params_scatter<<<blocks_for(n, 256), 256, 0, st>>>(l->blobs[kind], dsrc, l->map, n);
  RMC_KERNEL_OK();
  if (kind == RMC_ONLINE) ++l->online_version;
  if (kind == RMC_TARGET) ++l->target_version;
  if (src_is_host) RMC_CUDA(cudaStreamSynchronize(st));
  return RMC_OK;
}

extern "C" int32_t rmc_learner_get_params(rmc_learner_t* l, int32_t kind, float* dst, int64_t n, int32_t dst_is_host, rmc_stream_t s) {
  if (!l || kind < 0 || kind > 4 || !dst || n != l->P) return fail(RMC_ERR_ARG, "rmc_learner_get_params: bad kind/size");
  if (int32_t e = use_device(l->device)) return e;
  cudaStream_t st = as_stream(s);
  float* ddst = dst_is_host ? l->io : dst;
  k_params_gather<<<blocks_for(n, 256), 256, 0, st>>>(ddst, l->blobs[kind], l->map, n);
  RMC_KERNEL_OK();
  if (dst_is_host) {
    RMC_CUDA(cudaMemcpyAsync(dst, l->io, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost, st));
    RMC_CUDA(cudaStreamSynchronize(st));
  }
  return RMC_OK;
}

static int32_t fill_scalars(const rmc_learner* l, const rmc_step_args_t* a, StepScalars* S) {
  const rmc_hyper_t& h = l->hyper;
  std::memset(S, 0, sizeof(*S));
  S->B = a->batch;
  S->Bglobal = a->global_batch > 0 ? a->global_batch : a->batch;
  S->shard_off = a->shard_offset;
  S->phases = a->phases;
  S->double_dqn = l->spec.double_dqn;
  S->prioritized = l->spec.prioritized;
  S->beta = a->per_beta;
  S->u = a->u_dev;
  S->idx = reinterpret_cast<const long long*>(a->idx_dev);
  S->seed = a->seed;
  S->counter = a->counter;
  S->grads_in = a->grads_in_dev;
  S->gamma = static_cast<float>(h.gamma);
  if (a->phases & RMC_PH_ADAM) {
    if (a->adam_t < 1) return fail(RMC_ERR_ARG, "rmc_learner_step: adam_t must be >= 1");
    // torch/optim/adam.py (_single_tensor_adam): python-float bias corrections
    const double t = static_cast<double>(a->adam_t);
    const double bc1 = 1.0 - std::pow(h.adam_beta1, t);
    const double bc2 = 1.0 - std::pow(h.adam_beta2, t);
    const double step_size = h.lr / bc1;
    const double bc2_sqrt = std::pow(bc2, 0.5);
    S->adam_w1 = static_cast<float>(1.0 - h.adam_beta1);
    S->adam_b2 = static_cast<float>(h.adam_beta2);
    S->adam_w2 = static_cast<float>(1.0 - h.adam_beta2);
    S->adam_neg_step = static_cast<float>(-step_size);
    S->adam_bc2_sqrt = static_cast<float>(bc2_sqrt);
    S->adam_eps = static_cast<float>(h.adam_eps);
  }
  S->polyak_k = static_cast<float>(h.polyak_k);
  S->polyak_1mk = static_cast<float>(1.0 - h.polyak_k);
  S->per_eps = static_cast<float>(h.per_eps);
  S->per_alpha = static_cast<float>(h.per_alpha);
  S->per_pmax = static_cast<float>(h.per_pmax);
  return RMC_OK;
}

static int32_t check_step(const rmc_learner* l, const rmc_replay* r, const rmc_step_args_t* a) {
  if (!l || !r || !a) return fail(RMC_ERR_ARG, "rmc_learner_step: null");
  if (a->batch < 1 || a->batch > l->max_batch) return fail(RMC_ERR_ARG, "rmc_learner_step: batch outside [1, max_batch]");
  if (r->D != l->spec.obs_dim) return fail(RMC_ERR_ARG, "rmc_learner_step: replay obs_dim differs from the learner's");
  if ((l->spec.prioritized != 0) != (r->prioritized != 0))
    return fail(RMC_ERR_ARG, "rmc_learner_step: prioritized learner needs a prioritized replay (and vice versa)");
  if (r->device != l->device) return fail(RMC_ERR_ARG, "rmc_learner_step: replay and learner on different devices");
  if (a->phases & RMC_PH_SAMPLE) {
    if (r->size < 1) return fail(RMC_ERR_STATE, "rmc_learner_step: empty replay");
    if (!r->prioritized && a->batch > r->size && a->idx_dev == nullptr)
      return fail(RMC_ERR_STATE, "rmc_learner_step: sample larger than population");
  }
  return RMC_OK;
}

// One launch of the fused step.  Its in-kernel agent barrier and hand-off words need every CTA of the grid co-resident.
//   pdl (default)  programmatic dependent launch: the next step's launch latency and prologue overlap this step's tail
//                  (idle gap between steps 4.9 -> 2.2 us).  Co-residency is established by construction instead of by
//                  the driver: the grid never exceeds the number of CTAs that fit the device at once (checked against the
//                  occupancy calculator when the learner / group is created), steps of one stream are stream-ordered, and
//                  fused-step launches of ONE device on DIFFERENT streams are serialised here with an event edge -- so two
//                  grids of this kernel never share the device half-scheduled (the only way the barrier could wait for a
//                  CTA that cannot start).  Kernels of other libraries on other streams finish by themselves and only
//                  delay a step.  What remains uncovered is a second PROCESS running this kernel on the same GPU through
//                  MPS: the in-kernel watchdog (SpinGuard) turns that into RMC_ERR_STATE instead of a hang; use
//                  RMC_LAUNCH=coop there.
//   coop           cooperative launch (driver-guaranteed co-residency; no overlap with the previous step's tail)
//   plain / pdlcoop  diagnostics
static int launch_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("RMC_LAUNCH");
    mode = (e && std::strcmp(e, "plain") == 0) ? 1 : (e && std::strcmp(e, "coop") == 0) ? 0 : (e && std::strcmp(e, "pdlcoop") == 0) ? 3 : 2;
  }
  return mode;
}
static constexpr int kMaxDevices = 64;
struct StepStreamGuard {            // per device: the stream the last fused-step launch went to
  std::mutex mu;
  bool any = false;
  cudaStream_t last = nullptr;
  cudaEvent_t ev = nullptr;
};
static StepStreamGuard g_step_guard[kMaxDevices];

static int32_t launch_step(int device, dim3 grid, void** args, size_t smem, cudaStream_t st, int path) {
  const int mode = launch_mode();
  void* fn = path == 1 ? reinterpret_cast<void*>(k_learner_step<1>) : path == 2 ? reinterpret_cast<void*>(k_learner_step<2>) : reinterpret_cast<void*>(k_learner_step<0>);
  if (mode == 0) {
    RMC_CUDA(cudaLaunchCooperativeKernel(fn, grid, dim3(kThreads, 1, 1), args, smem, st));
  } else {
    StepStreamGuard& G = g_step_guard[device >= 0 && device < kMaxDevices ? device : 0];
    std::lock_guard<std::mutex> lock(G.mu);
    if (G.any && G.last != st) {      // stream switch: this grid starts only after everything queued on the previous stream
      if (G.ev == nullptr) RMC_CUDA(cudaEventCreateWithFlags(&G.ev, cudaEventDisableTiming));
      if (cudaEventRecord(G.ev, G.last) == cudaSuccess) {
        RMC_CUDA(cudaStreamWaitEvent(st, G.ev, 0));
      } else {                        // the previous stream no longer exists: its work is ordered by a device-wide wait
        cudaGetLastError();
        RMC_CUDA(cudaDeviceSynchronize());
      }
    }
    G.any = true; G.last = st;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = (mode == 2) ? 1 : (mode == 3) ? 2 : 0;
    RMC_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return RMC_OK;
}
// co-residency by construction: how many CTAs of the fused step fit the device at once
static int32_t step_resident_ctas(int device, int smem_bytes, int* out) {
  int per_sm_a = 0, per_sm_b = 0, per_sm_c = 0, sms = 0;
  RMC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  RMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, k_learner_step<1>, kThreads, static_cast<size_t>(smem_bytes)));
  RMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_learner_step<0>, kThreads, static_cast<size_t>(smem_bytes)));
  RMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, k_learner_step<2>, kThreads, static_cast<size_t>(smem_bytes)));
  *out = std::min(per_sm_a, std::min(per_sm_b, per_sm_c)) * sms;
  return RMC_OK;
}

// The one-tile instantiation keeps the sampled rows in shared memory between SAMPLE and FORWARD: a FORWARD-only launch
// (rows already in X from an earlier launch) takes the general instantiation, which reads X.
static bool one_tile_ok(const StepScalars& S, long long n_tiles) {
  return n_tiles <= S.n_row_ctas && ((S.phases & (RMC_PH_SAMPLE | RMC_PH_FORWARD)) != RMC_PH_FORWARD);
}
// Which instantiation of k_learner_step runs a launch: 1 = one 4-row tile per row CTA (the default batches), 2 = the
// batch-stationary phases of rmc_rows_ws.cuh once a row CTA owns at least two 16-row tiles, 0 = several 4-row tiles per CTA.
static int step_path(const StepScalars& S, long long n_tiles, const AgentCtx& c) {
  if (one_tile_ok(S, n_tiles)) return 1;
  if ((S.phases & RMC_PH_FORWARD) && c.gpart != nullptr && (S.B + kWR - 1) / kWR >= 2ll * S.n_row_ctas) return 2;
  return 0;
}

static int grid_for(const rmc_learner* l, long long B, int max_ctas) {
  const long long n_tiles = (B + kTM - 1) / kTM;
  const int want = static_cast<int>(std::min<long long>(max_ctas, std::max<long long>(n_tiles, 154)));
  return std::max(1, std::min(want, max_ctas));
}

// RMC_L2_PERSIST=1 (diagnostic): pin the sum tree in the persisting part of L2 for the launching stream.
static void maybe_persist_tree(rmc_replay* r, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = std::getenv("RMC_L2_PERSIST"); enabled = (e && e[0] == '1') ? 1 : 0; }
  if (!enabled || !r->prioritized || r->l2_window_set) return;
  r->l2_window_set = true;
  int max_persist = 0, max_window = 0;
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, r->device);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, r->device);
  const size_t bytes = static_cast<size_t>(r->n_nodes) * sizeof(double);
  const size_t want = std::min<size_t>(bytes, static_cast<size_t>(std::min(max_persist, max_window)));
  if (want == 0) return;
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
  cudaStreamAttrValue v{};
  v.accessPolicyWindow.base_ptr = r->dev.tree;
  v.accessPolicyWindow.num_bytes = want;
  v.accessPolicyWindow.hitRatio = 1.0f;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

// ---- tensor-core (bf16 / tcgen05) learner step: the dense large-batch mode (rmc_tc_train.cuh) ------------------
static int32_t tc_train_setup(rmc_learner* l) {
  if (l->tct_ready) return RMC_OK;
  const size_t B = static_cast<size_t>(l->max_batch);
  int32_t e = RMC_OK;
  if (l->tc_packed == nullptr) {
    if ((e = owned_alloc(l, &l->tc_packed, static_cast<size_t>(kTcBlobBytes)))) return e;
  }
  if ((e = owned_alloc(l, &l->tc_packed_target, static_cast<size_t>(kTcBlobBytes)))) return e;
  if ((e = owned_alloc(l, &l->tc_packed_bwd, static_cast<size_t>(kTcBwdElems)))) return e;
  TcTrainBufs& T = l->tct;
  if ((e = owned_alloc(l, &T.heads_n, B * kTcNH))) return e;
  if ((e = owned_alloc(l, &T.heads_t, B * kTcNH))) return e;
  if ((e = owned_alloc(l, &T.heads_s, B * kTcNH))) return e;
  const size_t Bt = ((B + kTcRows - 1) / kTcRows) * kTcRows;        // whole 128-row tile images
  if ((e = owned_alloc(l, &T.Xb, Bt * kTcK1))) return e;
  if ((e = owned_alloc(l, &T.H1b, Bt * kH1))) return e;
  if ((e = owned_alloc(l, &T.H2b, Bt * kH2))) return e;
  if ((e = owned_alloc(l, &T.DHb, Bt * kTcNH))) return e;
  if ((e = owned_alloc(l, &T.partials, static_cast<size_t>(l->num_sms) * l->L.total))) return e;
  RMC_CUDA(cudaFuncSetAttribute(k_mlp_infer_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  RMC_CUDA(cudaFuncSetAttribute(k_tc_fwd3, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  RMC_CUDA(cudaStreamCreateWithFlags(&l->tc_side, cudaStreamNonBlocking));
  for (auto& ev : l->tc_ev) RMC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  RMC_CUDA(cudaFuncSetAttribute(k_tc_bwd_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcBwdFusedSmemBytes));
  l->tct_ready = true;
  return RMC_OK;
}

static int32_t tree_update_large(rmc_replay* r, const long long* nodes, const float* pri, long long n, bool stamps_done, cudaStream_t st);
// sharded tensor-core step: gather every rank's (leaf, |td|) slice and apply the replicated write-back of the GLOBAL batch on
// the side stream (fork from `st`; tc_ev[1] marks its end)
static int32_t comm_side_writeback(rmc_learner* l, rmc_replay* r, cudaStream_t st) {
  rmc_comm* c = l->early.c;
  const long long Bg = l->early.Bg;
  cudaStream_t ts = l->tc_side;
  const bool patching = g_trace != nullptr && g_trace->mode == 2;      // replaying a graph: the fork / join edges are in it
  if (!patching) {
    RMC_CUDA(cudaEventRecord(l->tc_ev[0], st));
    RMC_CUDA(cudaStreamWaitEvent(ts, l->tc_ev[0], 0));
  }
  RMC_CUDA(launch_pdl(k_comm_gather_td, dim3(static_cast<unsigned>(std::min<long long>(256, (Bg + 255) / 256))), dim3(256), 0, ts, l->early.V, l->early.parity, c->epoch,
                      c->g_nodes, c->g_td));
  RMC_KERNEL_OK();
  const float eps = static_cast<float>(l->hyper.per_eps), alpha = static_cast<float>(l->hyper.per_alpha), pmax = static_cast<float>(l->hyper.per_pmax);
  if (Bg <= kTreeCtaMax) {
    RMC_CUDA(launch_pdl(k_tree_update_small, dim3(1), dim3(kThreads), 0, ts, r->dev, c->g_nodes, nullptr, c->g_td, c->g_pri, Bg, eps, alpha, pmax));
    RMC_KERNEL_OK();
  } else {
    RMC_CUDA(launch_pdl(k_td_to_pri, dim3(blocks_for(Bg, 256)), dim3(256), 0, ts, c->g_td, c->g_pri, Bg, eps, alpha, pmax));
    RMC_KERNEL_OK();
    if (int32_t e = tree_update_large(r, c->g_nodes, c->g_pri, Bg, false, ts)) return e;
  }
  if (!patching) RMC_CUDA(cudaEventRecord(l->tc_ev[1], ts));
  return RMC_OK;
}

// RMC_TC_EVENTS=1 (diagnostic): CUDA events between the kernels of the tensor-core step on the main stream; the durations of
// the PREVIOUS step are printed to stderr at the start of the next one (warm caches, real stream order -- unlike ncu's
// serialised cold-cache replays).  Disables the step graph.
struct TcEvents {
  bool on = false, armed = false;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
  size_t n = 0;
  void mark(const char* name, cudaStream_t st) {
    if (!on) return;
    if (n == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); names.push_back(name); }
    names[n] = name;
    cudaEventRecord(ev[n++], st);
  }
  void report() {
    if (!on || !armed || n < 2) return;
    cudaEventSynchronize(ev[n - 1]);
    std::string line = "[rmc] tc step (us):";
    for (size_t k = 1; k < n; ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[k - 1], ev[k]);
      char buf[96];
      std::snprintf(buf, sizeof buf, " %s %.1f", names[k], 1e3f * ms);
      line += buf;
    }
    float tot = 0.f;
    cudaEventElapsedTime(&tot, ev[0], ev[n - 1]);
    std::fprintf(stderr, "%s | total %.1f\n", line.c_str(), 1e3f * tot);
  }
};
static TcEvents g_tc_ev;
static bool tc_events_on() {
  static const bool on = [] { const char* e = std::getenv("RMC_TC_EVENTS"); return e && e[0] == '1'; }();
  return on;
}

static int32_t step_tc(rmc_learner* l, rmc_replay* r, const rmc_step_args_t* a, StepScalars& S, cudaStream_t st) {
  g_tc_ev.on = tc_events_on();
  g_tc_ev.report();
  g_tc_ev.n = 0; g_tc_ev.armed = true;
  g_tc_ev.mark("start", st);
  if (l->L.D > kTcK1 - 1) return fail(RMC_ERR_UNSUPPORTED, "tensor-core learner mode: obs_dim must be <= 15 (column 15 of the X tile carries the bias-gradient ones)");
  if ((a->phases & (RMC_PH_FORWARD | RMC_PH_BACKWARD)) != (RMC_PH_FORWARD | RMC_PH_BACKWARD))
    return fail(RMC_ERR_UNSUPPORTED, "tensor-core learner mode needs RMC_PH_FORWARD and RMC_PH_BACKWARD in one step");
  if (a->grads_in_dev != nullptr) return fail(RMC_ERR_ARG, "tensor-core learner mode: grads_in_dev belongs to an Adam-only step");
  if (int32_t e = tc_train_setup(l)) return e;
  const long long B = a->batch;
  AgentCtx& C = l->ctx;
  TcTrainBufs T = l->tct;
  if (a->phases & RMC_PH_SAMPLE) {
    if (r->prioritized) {
      if (int32_t e = launch_per_sample(r->dev, B, S.Bglobal, S.shard_off, S.beta, S.u, S.seed, S.counter, C.nodes, C.is_w, C.X, C.leaf_p, st)) return e;
    } else {
      RMC_CUDA(launch_pdl(k_uniform_sample, dim3(blocks_for(B, kWarps)), dim3(kThreads), 0, st, r->dev, B, S.shard_off, S.idx, S.seed, S.counter, 0u, C.nodes, C.X));
      RMC_KERNEL_OK();
    }
  }
  g_tc_ev.mark("sample", st);
  // bf16 operand images of the online net (forward + backward forms) and of the target net
  if (l->tc_packed_version != l->online_version) {
    RMC_CUDA(launch_pdl(k_tc_pack, dim3(blocks_for(kH2 * kH1, 256)), dim3(256), 0, st, l->blobs[RMC_ONLINE], l->L, l->tc_packed));
    RMC_KERNEL_OK();
    l->tc_packed_version = l->online_version;
  }
  if (l->tc_bwd_version != l->online_version) {
    RMC_CUDA(launch_pdl(k_tc_pack_bwd, dim3(blocks_for(kH1 * kH2, 256)), dim3(256), 0, st, l->blobs[RMC_ONLINE], l->L, l->tc_packed_bwd));
    RMC_KERNEL_OK();
    l->tc_bwd_version = l->online_version;
  }
  if (l->tc_target_version != l->target_version) {
    RMC_CUDA(launch_pdl(k_tc_pack, dim3(blocks_for(kH2 * kH1, 256)), dim3(256), 0, st, l->blobs[RMC_TARGET], l->L, l->tc_packed_target));
    RMC_KERNEL_OK();
    l->tc_target_version = l->target_version;
  }
  const long long n_tiles = (B + kTcRows - 1) / kTcRows;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(n_tiles, l->num_sms));
  // the three forwards in ONE launch: online(s'), target(s'), online(s) (+ saved bf16 activations) on disjoint CTA ranges
  TcFwdJobs J{};
  J.j[0].packed = l->tc_packed;        J.j[0].heads_out = T.heads_n;
  J.j[1].packed = l->tc_packed_target; J.j[1].heads_out = T.heads_t;
  J.j[2].packed = l->tc_packed;        J.j[2].heads_out = T.heads_s;
  J.j[0].X.row_stride = J.j[1].X.row_stride = J.j[2].X.row_stride = l->rf;
  J.j[0].X.act = J.j[1].X.act = J.j[2].X.act = l->L.act;
  J.j[0].X.col_off = J.j[1].X.col_off = l->L.D;      // s' rows: next_obs sits D floats into the gathered row
  J.j[2].X.col_off = 0; J.j[2].X.Xb = T.Xb; J.j[2].X.H1b = T.H1b; J.j[2].X.H2b = T.H2b;
  int c0, c2;
  if (3 * n_tiles <= l->num_sms) { c0 = c2 = static_cast<int>(n_tiles); }
  else { c0 = std::max(1, l->num_sms / 3); c2 = l->num_sms - 2 * c0; }
  J.j[0].cta_begin = 0;      J.j[0].cta_count = c0;
  J.j[1].cta_begin = c0;     J.j[1].cta_count = c0;
  J.j[2].cta_begin = 2 * c0; J.j[2].cta_count = c2;
  g_tc_ev.mark("pack", st);
  RMC_CUDA(launch_pdl(k_tc_fwd3, dim3(2 * c0 + c2), dim3(kTcFwdThreads), kTcSmemBytes, st, J, l->L.D, l->L.A, l->L.NH, l->L.dueling, C.X, B));
  RMC_KERNEL_OK();
  g_tc_ev.mark("fwd3", st);
  l->ctx.rp = r->dev;
  const unsigned td_blocks = blocks_for(B, kTdThreads);
  if (td_blocks > 1024) return fail(RMC_ERR_UNSUPPORTED, "tensor-core learner mode: batch above 131,072");
  RMC_CUDA(launch_pdl(k_tc_td, dim3(td_blocks), dim3(kTdThreads), 0, st, l->ctx, S, T));
  RMC_KERNEL_OK();
  g_tc_ev.mark("td", st);
  // The priority write-back needs only |td| and the sampled leaves: it runs on a side stream beside the backward and
  // Adam kernels (latency-bound tree kernels next to tensor-core CTAs) and joins before the step returns.
  static const bool side_tree = [] { const char* e = std::getenv("RMC_TC_SIDE_TREE"); return !(e && e[0] == '0'); }();
  const bool write_back = (a->phases & RMC_PH_PRIORITY) && l->spec.prioritized;
  if (l->early.c != nullptr) {      // sharded step: this rank's (leaf, |td|) slice leaves now, long before its gradients
    RMC_CUDA(launch_pdl(k_comm_publish_td, dim3(static_cast<unsigned>(std::max<long long>(1, std::min<long long>(148, (B + 1023) / 1024)))), dim3(256), 0, st, l->early.V,
                        l->early.parity, l->early.c->epoch, C.nodes, C.abs_td, B, l->early.c->arrive_td));
    RMC_KERNEL_OK();
    if (l->early.issue_side)
      if (int32_t e = comm_side_writeback(l, r, st)) return e;
  }
  if (write_back) {
    cudaStream_t ts = side_tree ? l->tc_side : st;
    const bool patching = g_trace != nullptr && g_trace->mode == 2;      // replaying a graph: the fork / join edges are in it
    if (ts != st && !patching) {
      RMC_CUDA(cudaEventRecord(l->tc_ev[0], st));
      RMC_CUDA(cudaStreamWaitEvent(ts, l->tc_ev[0], 0));
    }
    if (B <= kTreeCtaMax) {
      RMC_CUDA(launch_pdl(k_tree_update_small, dim3(1), dim3(kThreads), 0, ts, r->dev, C.nodes, nullptr, C.abs_td, C.pri, B, S.per_eps, S.per_alpha, S.per_pmax));
      RMC_KERNEL_OK();
    } else {
      RMC_CUDA(launch_pdl(k_td_to_pri, dim3(blocks_for(B, 256)), dim3(256), 0, ts, C.abs_td, C.pri, B, S.per_eps, S.per_alpha, S.per_pmax));
      RMC_KERNEL_OK();
      if (int32_t e = tree_update_large(r, C.nodes, C.pri, B, false, ts)) return e;
    }
    if (ts != st && !patching) RMC_CUDA(cudaEventRecord(l->tc_ev[1], ts));
  }
  // backward: dgrad chain + weight gradients fused per 128-row tile
  T.n_part = static_cast<int>(grid);
  g_tc_ev.mark("fork", st);
  RMC_CUDA(launch_pdl(k_tc_bwd_fused, dim3(grid), dim3(kThreads), kTcBwdFusedSmemBytes, st, l->ctx, reinterpret_cast<const unsigned char*>(l->tc_packed_bwd), B, T));
  RMC_KERNEL_OK();
  g_tc_ev.mark("bwd", st);
  l->epoch = (l->epoch >= 0x7fffffffu) ? 1u : l->epoch + 1u;
  S.epoch = l->epoch;
  // the Adam kernel refreshes the bf16 operand images element by element: no pack kernels on the next step
  const TcPackOut P{l->tc_packed, l->tc_packed_bwd, l->tc_packed_target};
  RMC_CUDA(launch_pdl(k_tc_reduce_adam, dim3(blocks_for(l->L.total, 128)), dim3(256), 0, st, l->ctx, S, T, static_cast<int>(td_blocks), P, 0));
  RMC_KERNEL_OK();
  g_tc_ev.mark("reduce_adam", st);
  if (g_tc_ev.on) {      // diagnostic: the same kernel again on warm inputs (how much of its time is the first touch of the partials?)
    static const int rep = [] { const char* e = std::getenv("RMC_TC_REDUCE_REPEAT"); return e ? std::atoi(e) : 0; }();
    static const int skip = [] { const char* e = std::getenv("RMC_TC_REDUCE_SKIP"); return e ? std::atoi(e) : 0; }();
    for (int k = 0; k < rep; ++k) {
      RMC_CUDA(launch_pdl(k_tc_reduce_adam, dim3(blocks_for(l->L.total, 128)), dim3(256), 0, st, l->ctx, S, T, static_cast<int>(td_blocks), P, skip));
      RMC_KERNEL_OK();
    }
    if (rep > 0) g_tc_ev.mark("reduce_adam_again_xN", st);
  }
  l->loss_epoch = S.epoch;
  if (a->phases & RMC_PH_ADAM) l->tc_packed_version = l->tc_bwd_version = ++l->online_version;
  if (a->phases & (RMC_PH_POLYAK | RMC_PH_HARDSYNC)) l->tc_target_version = ++l->target_version;
  if (write_back && side_tree && !(g_trace != nullptr && g_trace->mode == 2)) RMC_CUDA(cudaStreamWaitEvent(st, l->tc_ev[1], 0));
  g_tc_ev.mark("join_tree", st);
  return RMC_OK;
}

// ------------------------------------------------------------------------------ hybrid CNN + MLP learner
static constexpr long long kHybWsFloats = 4ll << 20;      // split-K workspace (16 MB)
extern "C" int32_t rmc_learner_create_hybrid(rmc_learner_t** out, const rmc_hybrid_spec_t* sp, const rmc_hyper_t* hyper, int64_t max_batch,
                                             int32_t device) {
  if (!out || !sp || !hyper || max_batch < 1) return fail(RMC_ERR_ARG, "rmc_learner_create_hybrid: null/bad args");
  if (sp->n_conv < 1 || sp->n_conv > kHybMaxConv || sp->n_dense < 1 || sp->n_dense > kHybMaxDense || sp->n_actions < 1 || sp->n_actions > 15 ||
      sp->macro_len < 0 || sp->grid_c < 1 || sp->grid_h < 1 || sp->grid_w < 1)
    return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create_hybrid: 1..4 conv layers, 1..3 dense layers, n_actions 1..15");
  if (sp->activation != RMC_ACT_RELU && sp->activation != RMC_ACT_ELU) return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create_hybrid: activation must be ReLU or ELU(alpha=1)");
  if (max_batch > 4096) return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create_hybrid: max_batch above 4096");
  if (int32_t e = use_device(device)) return e;
  HybNet N{};
  N.n_conv = sp->n_conv; N.n_dense = sp->n_dense;
  N.macro_len = sp->macro_len; N.grid_len = sp->grid_c * sp->grid_h * sp->grid_w; N.D = N.macro_len + N.grid_len;
  N.A = sp->n_actions; N.dueling = sp->dueling ? 1 : 0; N.NH = N.dueling ? N.A + 1 : N.A; N.act = (sp->activation == RMC_ACT_ELU) ? 1 : 0;
  int po = 0, ro = 0;                       // parameter offset (torch state_dict order), record offset
  int ic = sp->grid_c, ih = sp->grid_h, iw = sp->grid_w, prev_out = -1;
  for (int i = 0; i < N.n_conv; ++i) {     // net.cnn_stream.{2i}.weight / .bias  (env/dqn_config.py:84-93: 3x3, padding 1)
    HybConv& c = N.conv[i];
    c.ic = ic; c.ih = ih; c.iw = iw; c.oc = sp->conv_out[i]; c.sh = sp->conv_sh[i]; c.sw = sp->conv_sw[i];
    if (c.oc < 4 || (c.oc & 3) || c.sh < 1 || c.sw < 1 || (i > 0 && (c.ic & 3))) return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create_hybrid: conv channels must be multiples of 4");
    c.oh = (ih + 2 - 3) / c.sh + 1; c.ow = (iw + 2 - 3) / c.sw + 1;
    c.w_off = po; po += c.oc * c.ic * 9;
    c.b_off = po; po += c.oc;
    c.in_off = prev_out;
    c.out_off = ro;
    if (c.ic * c.ih * c.iw > 11000 || c.oc * c.oh * c.ow > 11000) return fail(RMC_ERR_UNSUPPORTED, "rmc_learner_create_hybrid: feature map too large for the shared-memory staging");
    prev_out = ro;
    ro += c.oc * c.oh * c.ow;
    ic = c.oc; ih = c.oh; iw = c.ow;
  }
  N.conv_flat = ic * ih * iw;
  N.feat_off = N.conv[N.n_conv - 1].out_off;
  N.feat_len = N.conv_flat + N.macro_len;
  ro = round4(N.feat_off + N.feat_len);
  int in_len = N.feat_len, in_off = N.feat_off;
  for (int i = 0; i < N.n_dense; ++i) {    // net.dense_stream.{2i}.weight / .bias
    HybDense& d = N.dense[i];
    d.in = in_len; d.out = sp->dense_out[i];
    if (d.out < 1) return fail(RMC_ERR_ARG, "rmc_learner_create_hybrid: dense width");
    d.w_off = po; po += d.out * d.in;
    d.b_off = po; po += d.out;
    d.in_off = in_off; d.out_off = ro;
    in_off = ro; in_len = d.out;
    ro = round4(ro + d.out);
  }
  N.last_off = in_off; N.last_len = in_len;
  N.head_off = ro; ro += kQLD;
  N.rec = round4(ro);
  if (N.dueling) {                          // fc_val.weight, fc_val.bias, fc_adv.weight, fc_adv.bias (network.py:81-82)
    N.hw_off[0] = po; po += N.last_len; N.hb_off[0] = po; po += 1;
    N.hw_off[1] = po; po += N.A * N.last_len; N.hb_off[1] = po; po += N.A;
  } else {                                  // fc_out.weight, fc_out.bias (network.py:54)
    N.hw_off[0] = po; po += N.A * N.last_len; N.hb_off[0] = po; po += N.A;
    N.hw_off[1] = N.hb_off[1] = -1;
  }
  const int P = po;
  N.total = round4(P);
  auto* l = new rmc_learner();
  auto build = [&]() -> int32_t {
  l->device = device; l->hybrid = true; l->H = N; l->hyper = *hyper; l->max_batch = max_batch;
  l->spec.obs_dim = N.D; l->spec.n_actions = N.A; l->spec.dueling = N.dueling; l->spec.double_dqn = sp->double_dqn; l->spec.prioritized = sp->prioritized;
  l->spec.activation = sp->activation; l->spec.hidden1 = 0; l->spec.hidden2 = 0;
  l->L = NetLayout{}; l->L.D = N.D; l->L.A = N.A; l->L.NH = N.NH; l->L.dueling = N.dueling; l->L.total = N.total; l->L.act = N.act;
  l->P = P;
  l->rf = round4(2 * N.D + 3);
  RMC_CUDA(cudaDeviceGetAttribute(&l->num_sms, cudaDevAttrMultiProcessorCount, device));
  int32_t e = RMC_OK;
  for (int k = 0; k < 5; ++k)
    if ((e = owned_alloc(l, &l->blobs[k], static_cast<size_t>(N.total)))) return e;
  std::vector<int> map(static_cast<size_t>(P));
  for (int i = 0; i < P; ++i) map[i] = i;  // the device layout IS the torch order
  if ((e = owned_alloc(l, &l->map, map.size()))) return e;
  RMC_CUDA(cudaMemcpy(l->map, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice));
  if ((e = owned_alloc(l, &l->io, map.size()))) return e;
  AgentCtx& c = l->ctx;
  c.L = l->L;
  c.online = l->blobs[0]; c.target = l->blobs[1]; c.adam_m = l->blobs[2]; c.adam_v = l->blobs[3]; c.grads = l->blobs[4];
  const size_t B = static_cast<size_t>(max_batch);
  if ((e = owned_alloc(l, &c.nodes, B))) return e;
  if ((e = owned_alloc(l, &c.leaf_p, B))) return e;
  float** per_sample[] = {&c.is_w, &c.q_sa, &c.y, &c.abs_td, &c.hub, &c.pri, &c.gcoef};
  for (float** p : per_sample)
    if ((e = owned_alloc(l, p, B))) return e;
  if ((e = owned_alloc(l, &c.QT, B * kQLD))) return e;
  if ((e = owned_alloc(l, &c.QN, B * kQLD))) return e;
  if ((e = owned_alloc(l, &c.Q, B * kQLD))) return e;
  if ((e = owned_alloc(l, &c.DH, B * kQLD))) return e;
  if ((e = owned_alloc(l, &c.X, B * l->rf))) return e;
  if ((e = owned_alloc(l, &c.loss_part, 1024))) return e;
  if ((e = owned_alloc(l, &c.loss, 1))) return e;
  if ((e = owned_alloc(l, &c.barrier, 1))) return e;
  if ((e = owned_alloc(l, &l->rec_on, 2 * B * N.rec))) return e;
  if ((e = owned_alloc(l, &l->rec_tg, B * N.rec))) return e;
  if ((e = owned_alloc(l, &l->drec, B * N.rec))) return e;
  if ((e = owned_alloc(l, &l->hyb_ws, static_cast<size_t>(kHybWsFloats)))) return e;
  if ((e = owned_alloc(l, &l->hyb_ws2, static_cast<size_t>(kHybWsFloats)))) return e;
  RMC_CUDA(cudaStreamCreateWithFlags(&l->hyb_side, cudaStreamNonBlocking));
  RMC_CUDA(cudaStreamCreateWithFlags(&l->hyb_cap, cudaStreamNonBlocking));
  for (auto& ev : l->hyb_ev) RMC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  {
    float* hp = nullptr;
    float* dp = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&hp), 64, cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&dp), hp, 0) == cudaSuccess) {
      for (int k = 0; k < 16; ++k) hp[k] = 0.f;
      l->host_loss = hp;
      c.host_loss = dp;
    } else {
      cudaGetLastError();
      c.host_loss = nullptr;
    }
  }
  RMC_CUDA(cudaDeviceSynchronize());
  return RMC_OK;
  };
  if (int32_t e = build()) {
    const std::string msg = g_err;
    rmc_learner_destroy(l);
    return fail(e, msg);
  }
  *out = l;
  return RMC_OK;
}

template <int MODE>
static int32_t hyb_gemm_mode(rmc_learner* l, HybGemm G, cudaStream_t st) {
  // skinny problems (few output tiles, long K) are split along K so that the operand streams of one GEMM are spread
  // over the SMs; the partial sums are combined in split order by a second kernel (deterministic)
  const long long tiles = static_cast<long long>(blocks_for(G.N, 64)) * blocks_for(G.M, 64);
  int splits = 1;
  if (tiles < 96 && G.K >= 128) {
    splits = static_cast<int>(std::min<long long>(64, std::min<long long>((2 * 148 + tiles - 1) / tiles, G.K / 64)));
    while (splits > 1 && static_cast<long long>(splits) * G.M * G.N > kHybWsFloats - 16 * 1024) --splits;
  }
  G.splits = splits; G.ws = (st == l->hyb_side) ? l->hyb_ws2 : l->hyb_ws;        // one split-K workspace per stream
  G.k_chunk = ((G.K + splits - 1) / splits + 15) / 16 * 16;
  if (splits > 1) G.splits = (G.K + G.k_chunk - 1) / G.k_chunk;
  RMC_CUDA(launch_pdl(k_hyb_gemm<MODE>, dim3(blocks_for(G.N, 64), blocks_for(G.M, 64), static_cast<unsigned>(G.splits)), dim3(256), 0, st, G));
  RMC_KERNEL_OK();
  if (G.splits > 1) {
    RMC_CUDA(launch_pdl(k_hyb_splitk_reduce<MODE>, dim3(blocks_for(static_cast<long long>(G.M) * G.N, 256)), dim3(256), 0, st, G));
    RMC_KERNEL_OK();
  }
  return RMC_OK;
}
static int32_t hyb_gemm(rmc_learner* l, const HybGemm& G, cudaStream_t st) { return hyb_gemm_mode<0>(l, G, st); }

// one forward pass of R rows through the net with parameters P into the records `rec`
static int32_t hybrid_forward(rmc_learner* l, const float* P, const HybSrc& src, float* rec, long long R, cudaStream_t st) {
  const HybNet& N = l->H;
  for (int i = 0; i < N.n_conv; ++i) {
    const HybConv& c = N.conv[i];
    if (i == 0 || c.ic * 9 < 64) {          // tiny K (first layer: 2 input channels): the direct kernel
      const unsigned slices = blocks_for(static_cast<long long>(c.oc / 4) * c.oh * c.ow, 256);
      RMC_CUDA(launch_pdl(k_hyb_conv_fwd, dim3(static_cast<unsigned>(R), slices, 1), dim3(256), static_cast<size_t>(c.ic) * c.ih * c.iw * sizeof(float), st, N, i, P, src, rec));
      RMC_KERNEL_OK();
      continue;
    }
    HybGemm G{};                             // implicit GEMM: [R*pixels x ic*9] . [ic*9 x oc]
    G.cv = c; G.img = rec + c.in_off; G.img_stride = N.rec;
    G.B = P + c.w_off; G.b_sk = 1; G.b_sn = c.ic * 9;
    G.C = rec + c.out_off; G.c_sm = N.rec; G.bias = P + c.b_off;
    G.M = static_cast<int>(R) * c.oh * c.ow; G.N = c.oc; G.K = c.ic * 9; G.epi = 0; G.act = N.act;
    if (int32_t e = hyb_gemm_mode<1>(l, G, st)) return e;
    if (i == N.n_conv - 1) {                 // features = [flattened last conv output | macro]
      RMC_CUDA(launch_pdl(k_hyb_copy_macro, dim3(blocks_for(R * N.macro_len, 128)), dim3(128), 0, st, N, src, rec, R));
      RMC_KERNEL_OK();
    }
  }
  for (int i = 0; i < N.n_dense; ++i) {
    const HybDense& d = N.dense[i];
    HybGemm G{};
    G.A = rec + d.in_off; G.a_sm = N.rec; G.a_sk = 1;
    G.B = P + d.w_off; G.b_sk = 1; G.b_sn = d.in;
    G.C = rec + d.out_off; G.c_sm = N.rec;
    G.bias = P + d.b_off; G.M = static_cast<int>(R); G.N = d.out; G.K = d.in; G.epi = 0; G.act = N.act;
    if (int32_t e = hyb_gemm(l, G, st)) return e;
  }
  RMC_CUDA(launch_pdl(k_hyb_heads_fwd, dim3(blocks_for(R, 8)), dim3(256), 0, st, N, P, rec, R));
  RMC_KERNEL_OK();
  return RMC_OK;
}

static int32_t hybrid_step_body(rmc_learner* l, rmc_replay* r, const rmc_step_args_t* a, StepScalars& S, cudaStream_t st) {
  const PdlScope no_pdl(false);
  const HybNet& N = l->H;
  AgentCtx& C = l->ctx;
  const long long B = a->batch;
  const int ph = a->phases;
  if (a->precision != RMC_PREC_FP32) return fail(RMC_ERR_UNSUPPORTED, "hybrid network: the exact fp32 path is the only one built");
  if ((ph & RMC_PH_BACKWARD) && !(ph & RMC_PH_FORWARD)) return fail(RMC_ERR_ARG, "hybrid network: BACKWARD needs FORWARD in the same step");
  if (ph & RMC_PH_SAMPLE) {
    if (r->prioritized) {
      if (int32_t e = launch_per_sample(r->dev, B, S.Bglobal, S.shard_off, S.beta, S.u, S.seed, S.counter, C.nodes, C.is_w, C.X, C.leaf_p, st)) return e;
    } else {
      RMC_CUDA(launch_pdl(k_uniform_sample, dim3(blocks_for(B, kWarps)), dim3(kThreads), 0, st, r->dev, B, S.shard_off, S.idx, S.seed, S.counter, 0u, C.nodes, C.X));
      RMC_KERNEL_OK();
    }
  }
  const HybSrc src_on{C.X, l->rf, B, N.D, 0};            // online pass: rows [0,B) = s', rows [B,2B) = s
  const unsigned td_blocks = blocks_for(B, 128);
  // Two streams: the target-network pass runs beside the online pass, and the weight-gradient kernels of a layer beside
  // the data-gradient chain below it (both need only that layer's deltas).  The kernels of this network are small, so the
  // overlap roughly halves the step.  RMC_HYB_STREAMS=0 keeps everything on the caller's stream.
  static const bool two_streams = [] { const char* e = std::getenv("RMC_HYB_STREAMS"); return !(e && e[0] == '0'); }();
  cudaStream_t side = two_streams ? l->hyb_side : st;
  int ev_next = 0;
  auto hand_over = [&](cudaStream_t from, cudaStream_t to) -> int32_t {     // `to` continues after everything queued on `from`
    if (from == to) return RMC_OK;
    if (g_trace != nullptr && g_trace->mode == 2) return RMC_OK;             // patching a graph: the edges are already in it
    cudaEvent_t ev = l->hyb_ev[ev_next++ & 15];
    RMC_CUDA(cudaEventRecord(ev, from));
    RMC_CUDA(cudaStreamWaitEvent(to, ev, 0));
    return RMC_OK;
  };
  if (ph & RMC_PH_FORWARD) {
    const HybSrc src_tg{C.X, l->rf, B, N.D, N.D};
    if (int32_t e = hand_over(st, side)) return e;
    if (int32_t e = hybrid_forward(l, C.target, src_tg, l->rec_tg, B, side)) return e;
    if (int32_t e = hybrid_forward(l, C.online, src_on, l->rec_on, 2 * B, st)) return e;
    if (int32_t e = hand_over(side, st)) return e;
    RMC_CUDA(launch_pdl(k_hyb_td, dim3(td_blocks), dim3(128), 0, st, C, S, N, l->rec_on, l->rec_tg, l->drec));
    RMC_KERNEL_OK();
  }
  if (ph & RMC_PH_BACKWARD) {
    // main stream: the data-gradient chain (heads -> dense -> conv); side stream: every layer's weight / bias gradients, each
    // released as soon as that layer's deltas exist
    const float* rec_s = l->rec_on + B * N.rec;           // records of the s rows
    RMC_CUDA(launch_pdl(k_hyb_heads_dgrad, dim3(blocks_for(B * N.last_len, 256)), dim3(256), 0, st, N, C.online, rec_s, l->drec, B));
    RMC_KERNEL_OK();
    if (int32_t e = hand_over(st, side)) return e;
    RMC_CUDA(launch_pdl(k_hyb_heads_wgrad, dim3(blocks_for(static_cast<long long>(N.NH) * N.last_len, 256)), dim3(256), 0, side, N, rec_s, l->drec, B, C.grads));
    RMC_KERNEL_OK();
    for (int i = N.n_dense - 1; i >= 0; --i) {
      const HybDense& d = N.dense[i];
      HybGemm W{};                                        // dW[n][k] = sum_r dZ[r][n] X[r][k]
      W.A = l->drec + d.out_off; W.a_sm = 1; W.a_sk = N.rec;
      W.B = rec_s + d.in_off; W.b_sk = N.rec; W.b_sn = 1;
      W.C = C.grads + d.w_off; W.c_sm = d.in; W.M = d.out; W.N = d.in; W.K = static_cast<int>(B); W.epi = 2;
      if (int32_t e = hyb_gemm(l, W, side)) return e;
      RMC_CUDA(launch_pdl(k_hyb_colsum, dim3(blocks_for(d.out, 128)), dim3(128), 0, side, l->drec + d.out_off, N.rec, static_cast<int>(B), d.out, C.grads + d.b_off));
      RMC_KERNEL_OK();
      HybGemm G{};                                        // dX[r][k] = (sum_n dZ[r][n] W[n][k]) * act'(X[r][k])
      G.A = l->drec + d.out_off; G.a_sm = N.rec; G.a_sk = 1;
      G.B = C.online + d.w_off; G.b_sk = d.in; G.b_sn = 1;
      G.C = l->drec + d.in_off; G.c_sm = N.rec; G.H = rec_s + d.in_off; G.h_sm = N.rec;
      G.M = static_cast<int>(B); G.N = (i == 0) ? N.conv_flat : d.in; G.K = d.out; G.epi = 1; G.act = N.act;
      if (int32_t e = hyb_gemm(l, G, st)) return e;
      if (int32_t e = hand_over(st, side)) return e;
    }
    for (int i = N.n_conv - 1; i >= 0; --i) {
      const HybConv& c = N.conv[i];
      const int npix = c.oh * c.ow;
      HybGemm W{};                           // dW[oc][(ic,ky,kx)] = sum_(row,pix) dZ . im2col   (split along K = rows x pixels)
      W.cv = c; W.dz = l->drec + c.out_off; W.dz_stride = N.rec;
      if (i == 0) { W.img = nullptr; W.src = src_on; W.src_row0 = B; W.macro_len = N.macro_len; }
      else { W.img = rec_s + c.in_off; W.img_stride = N.rec; }
      W.C = C.grads + c.w_off; W.c_sm = c.ic * 9; W.M = c.oc; W.N = c.ic * 9; W.K = static_cast<int>(B) * npix; W.epi = 2;
      if (int32_t e = hyb_gemm_mode<3>(l, W, side)) return e;
      {   // bias gradient in two fixed-order stages: (oc, 16 row slices) partial sums, then one thread per channel
        float* part = ((side == l->hyb_side) ? l->hyb_ws2 : l->hyb_ws) + kHybWsFloats - 16 * 1024;     // tail of the stream's workspace
        RMC_CUDA(launch_pdl(k_hyb_conv_bias_grad, dim3(static_cast<unsigned>(c.oc), 16, 1), dim3(256), 0, side, l->drec + c.out_off, N.rec, npix, B, part));
        RMC_KERNEL_OK();
        RMC_CUDA(launch_pdl(k_hyb_colsum, dim3(blocks_for(c.oc, 128)), dim3(128), 0, side, static_cast<const float*>(part), static_cast<long long>(c.oc), 16, c.oc, C.grads + c.b_off));
        RMC_KERNEL_OK();
      }
      if (i > 0) {                           // delta of the layer below: (gathered dZ . W) * act'(input activation)
        HybGemm G{};
        G.cv = c; G.dz = l->drec + c.out_off; G.dz_stride = N.rec;
        G.B = C.online + c.w_off;
        G.C = l->drec + c.in_off; G.c_sm = N.rec; G.H = rec_s + c.in_off;
        G.M = static_cast<int>(B) * c.ih * c.iw; G.N = c.ic; G.K = c.oc * 9; G.epi = 1; G.act = N.act;
        if (int32_t e = hyb_gemm_mode<2>(l, G, st)) return e;
        if (int32_t e = hand_over(st, side)) return e;
      }
    }
    if (int32_t e = hand_over(side, st)) return e;         // every gradient is in place before Adam
  }
  if (ph & (RMC_PH_FORWARD | RMC_PH_ADAM | RMC_PH_POLYAK | RMC_PH_HARDSYNC)) {
    l->epoch = (l->epoch >= 0x7fffffffu) ? 1u : l->epoch + 1u;
    S.epoch = l->epoch;
    const int write_loss = (ph & RMC_PH_FORWARD) ? 1 : 0;
    RMC_CUDA(launch_pdl(k_hyb_adam, dim3(blocks_for(N.total, 256)), dim3(256), 0, st, C, S, N.total, static_cast<int>(td_blocks), write_loss));
    RMC_KERNEL_OK();
    if (write_loss) l->loss_epoch = S.epoch;
    if (ph & RMC_PH_ADAM) ++l->online_version;
    if (ph & (RMC_PH_POLYAK | RMC_PH_HARDSYNC)) ++l->target_version;
  }
  if ((ph & RMC_PH_PRIORITY) && l->spec.prioritized) {
    if (B <= kTreeCtaMax) {
      RMC_CUDA(launch_pdl(k_tree_update_small, dim3(1), dim3(kThreads), 0, st, r->dev, C.nodes, nullptr, C.abs_td, C.pri, B, S.per_eps, S.per_alpha, S.per_pmax));
      RMC_KERNEL_OK();
    } else {
      RMC_CUDA(launch_pdl(k_td_to_pri, dim3(blocks_for(B, 256)), dim3(256), 0, st, C.abs_td, C.pri, B, S.per_eps, S.per_alpha, S.per_pmax));
      RMC_KERNEL_OK();
      if (int32_t e = tree_update_large(r, C.nodes, C.pri, B, false, st)) return e;
    }
  }
  return RMC_OK;
}

// The training step of the hybrid network as ONE graph launch (see launch_pdl): the first full step of a (batch, phases)
// shape is captured on an internal stream -- both streams of the body and their event edges included -- and instantiated;
// later steps re-trace the body in patch mode (no launches: only the nodes whose arguments changed are updated) and replay
// the graph on the caller's stream.  RMC_HYB_GRAPH=0 keeps ordinary launches.  Any mismatch or capture error falls back to
// ordinary launches (the body's host-side counters may then advance twice for that step, which only invalidates caches).
template <typename Body>
static int32_t run_step_graph(rmc_learner* l, long long key_batch, int key_phases, cudaStream_t st, Body body) {
  static const bool graphs = [] { const char* e = std::getenv("RMC_HYB_GRAPH"); return !(e && e[0] == '0'); }();
  if (!graphs || l->hyb_graph_off) return body(st);
  StepGraph* g = nullptr;
  for (auto& c : l->hyb_graphs)
    if (c.batch == key_batch && c.phases == key_phases) g = &c;
  if (g != nullptr) {
    Trace T; T.mode = 2; T.g = g;
    g_trace = &T;
    const int32_t e = body(st);
    g_trace = nullptr;
    static const bool dbg = std::getenv("RMC_HYB_GRAPH_DEBUG") != nullptr;
    if (dbg) std::fprintf(stderr, "[rmc] hybrid graph replay: %zu kernels, %d nodes patched, trace %s\n", g->recs.size(), T.patched, T.ok ? "ok" : "MISMATCH");
    if (e == RMC_OK && T.ok && T.cursor == g->recs.size()) {
      RMC_CUDA(cudaGraphLaunch(g->exec, st));
      return RMC_OK;
    }
    g->destroy();                                                   // the step changed shape: forget the graph, launch normally
    l->hyb_graphs.erase(l->hyb_graphs.begin() + (g - l->hyb_graphs.data()));
    if (e != RMC_OK) return e;
    return body(st);
  }
  if (l->hyb_graphs.size() >= 16) return body(st);
  if (l->hyb_cap == nullptr && cudaStreamCreateWithFlags(&l->hyb_cap, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    l->hyb_graph_off = true;
    return body(st);
  }
  StepGraph fresh; fresh.batch = key_batch; fresh.phases = key_phases;
  Trace T; T.mode = 1; T.g = &fresh;
  if (cudaStreamBeginCapture(l->hyb_cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    l->hyb_graph_off = true;
    return body(st);
  }
  g_trace = &T;
  const int32_t e = body(l->hyb_cap);
  g_trace = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(l->hyb_cap, &fresh.graph);
  bool ok = (e == RMC_OK) && T.ok && ce == cudaSuccess && fresh.graph != nullptr;
  if (ok) ok = cudaGraphInstantiate(&fresh.exec, fresh.graph, 0) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    fresh.destroy();
    l->hyb_graph_off = true;
    return body(st);
  }
  l->hyb_graphs.push_back(std::move(fresh));
  RMC_CUDA(cudaGraphLaunch(l->hyb_graphs.back().exec, st));
  return RMC_OK;
}

static int32_t hybrid_step(rmc_learner* l, rmc_replay* r, const rmc_step_args_t* a, StepScalars& S, cudaStream_t st) {
  const int full = RMC_PH_FORWARD | RMC_PH_BACKWARD | RMC_PH_ADAM;
  if ((a->phases & full) != full || a->precision != RMC_PREC_FP32) return hybrid_step_body(l, r, a, S, st);
  return run_step_graph(l, a->batch, a->phases, st, [&](cudaStream_t s_) { return hybrid_step_body(l, r, a, S, s_); });
}

// act / Q values / raw heads of n states [n][D] through the hybrid net (chunks of at most 2 * max_batch rows)
static int32_t hybrid_infer_body(rmc_learner* l, const float* params, const float* obs_dev, long long n, long long* actions, float* q, int mode, cudaStream_t st);
static int32_t hybrid_infer(rmc_learner* l, const float* params, const float* obs_dev, long long n, long long* actions, float* q, int mode, cudaStream_t st) {
  // the per-env-step act of the trainer (a handful of rows, ~11 small kernels): one graph launch per call shape
  if (n > 64) return hybrid_infer_body(l, params, obs_dev, n, actions, q, mode, st);
  return run_step_graph(l, n, 0x40000000 | (mode << 2) | (actions ? 1 : 0) | (q ? 2 : 0), st,
                        [&](cudaStream_t s_) { return hybrid_infer_body(l, params, obs_dev, n, actions, q, mode, s_); });
}
static int32_t hybrid_infer_body(rmc_learner* l, const float* params, const float* obs_dev, long long n, long long* actions, float* q, int mode, cudaStream_t st) {
  const PdlScope no_pdl(false);
  const HybNet& N = l->H;
  const long long cap = 2 * l->max_batch;
  const int per_row = (mode == 1) ? N.A : N.NH;
  for (long long off = 0; off < n; off += cap) {
    const long long m = std::min(cap, n - off);
    const HybSrc src{obs_dev + off * N.D, N.D, m, 0, 0};
    if (int32_t e = hybrid_forward(l, params, src, l->rec_on, m, st)) return e;
    RMC_CUDA(launch_pdl(k_hyb_outputs, dim3(blocks_for(m, 128)), dim3(128), 0, st, N, l->rec_on, m, actions ? actions + off : nullptr, q ? q + off * per_row : nullptr, mode));
    RMC_KERNEL_OK();
  }
  return RMC_OK;
}

extern "C" int32_t rmc_learner_step(rmc_learner_t* l, rmc_replay_t* r, const rmc_step_args_t* a, rmc_stream_t s) {
  if (int32_t e = check_step(l, r, a)) return e;
  if (int32_t e = use_device(l->device)) return e;
  cudaStream_t st = as_stream(s);
  maybe_persist_tree(r, st);
  StepScalars S;
  if (int32_t e = fill_scalars(l, a, &S)) return e;
  l->ctx.rp = r->dev;
  l->last_batch = a->batch;
  if (l->hybrid) return hybrid_step(l, r, a, S, st);
  if (a->precision == RMC_PREC_BF16_TC) {
    // single-GPU tensor-core step (13 launches on two streams): replayed as one graph once the operand images are current
    // (the first step, and any step after an un-fused target update, still packs them and launches normally); the sharded
    // step keeps ordinary launches -- its exchange kernels spin on peer flags.
    const int full = RMC_PH_FORWARD | RMC_PH_BACKWARD | RMC_PH_ADAM;
    const bool images_current = l->tct_ready && l->tc_packed_version == l->online_version && l->tc_bwd_version == l->online_version &&
                                l->tc_target_version == l->target_version;
    if (l->early.c == nullptr && !l->no_graph && images_current && (a->phases & full) == full && a->grads_in_dev == nullptr && !tc_events_on())
      return run_step_graph(l, a->batch, 0x20000000 | a->phases, st, [&](cudaStream_t s_) { return step_tc(l, r, a, S, s_); });
    return step_tc(l, r, a, S, st);
  }
  if (a->precision != RMC_PREC_FP32) return fail(RMC_ERR_ARG, "rmc_learner_step: unknown precision");
  const int G = grid_for(l, a->batch, l->num_sms);
  const long long n_tiles = (a->batch + kTM - 1) / kTM;
  S.n_row_ctas = static_cast<int>(std::min<long long>(G, n_tiles));
  const bool rows = (a->phases & (RMC_PH_SAMPLE | RMC_PH_FORWARD)) != 0;
  const bool phase_b = (a->phases & (RMC_PH_PRIORITY | RMC_PH_BACKWARD | RMC_PH_ADAM | RMC_PH_POLYAK | RMC_PH_HARDSYNC)) != 0;
  // Batches with several row tiles per CTA: draw the minibatch with the full-occupancy grid-wide samplers first
  // (the fused kernel runs 8 warps per SM, far too few to hide the prefix search's dependent round trips).
  if ((a->phases & RMC_PH_SAMPLE) && n_tiles > G) {
    if (r->prioritized) {
      if (int32_t e = launch_per_sample(r->dev, a->batch, S.Bglobal, S.shard_off, S.beta, S.u, S.seed, S.counter, l->ctx.nodes, l->ctx.is_w, l->ctx.X,
                                        l->ctx.leaf_p, st)) return e;
    } else {
      RMC_CUDA(launch_pdl(k_uniform_sample, dim3(blocks_for(a->batch, kWarps)), dim3(kThreads), 0, st, r->dev, a->batch, S.shard_off, S.idx, S.seed, S.counter, 0u, l->ctx.nodes, l->ctx.X));
      RMC_KERNEL_OK();
    }
    S.phases &= ~RMC_PH_SAMPLE;
  }
  if (rows && phase_b) S.barrier_target = l->barrier_count + static_cast<unsigned>(G);
  l->epoch = (l->epoch >= 0x7fffffffu) ? 1u : l->epoch + 1u;
  S.epoch = l->epoch;
  AgentCtx single = l->ctx;
  l->last_grid = G;
  const AgentCtx* many = nullptr;
  void* args[] = {&single, &many, &S};
  if (int32_t e = launch_step(l->device, dim3(G, 1, 1), args, static_cast<size_t>(l->smem_bytes), st, step_path(S, n_tiles, l->ctx))) return e;
  if (rows && phase_b) l->barrier_count = S.barrier_target;
  if ((a->phases & RMC_PH_FORWARD) && phase_b) l->loss_epoch = S.epoch;
  if (a->phases & RMC_PH_ADAM) ++l->online_version;
  if (a->phases & (RMC_PH_POLYAK | RMC_PH_HARDSYNC)) ++l->target_version;
  if ((a->phases & RMC_PH_PRIORITY) && l->spec.prioritized && a->batch > kTreeCtaMax) {
    RMC_CUDA(launch_pdl(k_td_to_pri, dim3(blocks_for(a->batch, 256)), dim3(256), 0, st, l->ctx.abs_td, l->ctx.pri, a->batch, S.per_eps, S.per_alpha, S.per_pmax));
    RMC_KERNEL_OK();
    if (int32_t e = tree_update_large(r, l->ctx.nodes, l->ctx.pri, a->batch, false, st)) return e;
  }
  return RMC_OK;
}

// ------------------------------------------------------------------------------ sharded learner: peer-memory exchange
static void shard_range_c(long long batch, int rank, int world, long long* lo, long long* hi) {   // = parallel.shard_range
  const long long base = batch / world, rem = batch % world;
  *lo = rank * base + std::min<long long>(rank, rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}

extern "C" int32_t rmc_comm_create(rmc_comm_t** out, rmc_learner_t* l, int32_t rank, int32_t world, int64_t global_batch_max) {
  if (!out || !l || world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || global_batch_max < world)
    return fail(RMC_ERR_ARG, "rmc_comm_create: bad args (world must be 1..8)");
  if (l->hybrid) return fail(RMC_ERR_UNSUPPORTED, "rmc_comm_create: the sharded step is built for the macro MLP only");
  if (int32_t e = use_device(l->device)) return e;
  auto* c = new rmc_comm();
  c->device = l->device; c->rank = rank; c->world = world; c->learner = l;
  c->global_batch_max = global_batch_max;
  c->local_max = (global_batch_max + world - 1) / world;
  if (c->local_max > l->max_batch) { delete c; return fail(RMC_ERR_ARG, "rmc_comm_create: local shard exceeds the learner's max_batch"); }
  auto up = [](long long x) { return (x + 255) & ~255ll; };
  CommView& V = c->view;
  V.rank = rank; V.world = world;
  V.off_loss = up(static_cast<long long>(l->L.total) * 4);
  V.off_nodes = V.off_loss + 256;
  V.off_td = V.off_nodes + up(c->local_max * 8);
  V.slot_bytes = V.off_td + up(c->local_max * 4);
  c->bytes = static_cast<size_t>(kCommHeaderBytes + 2 * V.slot_bytes);
  {
    const char* t = std::getenv("RMC_COMM_TIMEOUT_MS");
    const double ms = t ? std::atof(t) : 10000.0;
    V.timeout_ns = static_cast<unsigned long long>((ms > 0 ? ms : 10000.0) * 1e6);
  }
  auto build = [&]() -> int32_t {
    RMC_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->local), c->bytes));
    RMC_CUDA(cudaMemset(c->local, 0, c->bytes));
    int32_t e = RMC_OK;
    if ((e = dev_alloc(&c->arrive, 1))) return e;
    if ((e = dev_alloc(&c->arrive_td, 1))) return e;
    if ((e = dev_alloc(&c->verdict, 2))) return e;
    if ((e = dev_alloc(&c->g_nodes, static_cast<size_t>(global_batch_max)))) return e;
    if ((e = dev_alloc(&c->g_td, static_cast<size_t>(global_batch_max)))) return e;
    if ((e = dev_alloc(&c->g_pri, static_cast<size_t>(global_batch_max)))) return e;
    RMC_CUDA(cudaDeviceSynchronize());
    return RMC_OK;
  };
  if (int32_t e = build()) { rmc_comm_destroy(c); return e; }      // partially built handle: released by its destroy function
  V.verdict = c->verdict;
  *out = c;
  return RMC_OK;
}

extern "C" int32_t rmc_comm_export(rmc_comm_t* c, void* handle64_out, void** local_ptr_out) {
  if (!c) return fail(RMC_ERR_ARG, "rmc_comm_export: null");
  if (int32_t e = use_device(c->device)) return e;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (handle64_out) {
    cudaIpcMemHandle_t h;
    RMC_CUDA(cudaIpcGetMemHandle(&h, c->local));
    std::memcpy(handle64_out, &h, 64);
  }
  if (local_ptr_out) *local_ptr_out = c->local;
  return RMC_OK;
}

extern "C" int32_t rmc_comm_connect(rmc_comm_t* c, const void* handles64, void* const* same_process_ptrs) {
  if (!c || (!handles64 && !same_process_ptrs)) return fail(RMC_ERR_ARG, "rmc_comm_connect: null");
  if (c->connected) return fail(RMC_ERR_STATE, "rmc_comm_connect: already connected");
  if (int32_t e = use_device(c->device)) return e;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peer[r] = c->local; continue; }
    if (same_process_ptrs) {                      // ranks emulated inside one process (tests): plain device pointers
      c->peer[r] = same_process_ptrs[r];
      cudaPointerAttributes at{};
      RMC_CUDA(cudaPointerGetAttributes(&at, c->peer[r]));
      if (at.device != c->device) {
        int can = 0;
        RMC_CUDA(cudaDeviceCanAccessPeer(&can, c->device, at.device));
        if (!can) return fail(RMC_ERR_UNSUPPORTED, "rmc_comm_connect: no peer access between the devices");
        cudaError_t pe = cudaDeviceEnablePeerAccess(at.device, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) RMC_CUDA(pe);
        cudaGetLastError();
      }
    } else {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, static_cast<const unsigned char*>(handles64) + 64 * r, 64);
      RMC_CUDA(cudaIpcOpenMemHandle(&c->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
      c->ipc_opened[r] = true;
    }
  }
  for (int r = 0; r < kCommMaxWorld; ++r) c->view.base[r] = static_cast<unsigned char*>(r < c->world ? c->peer[r] : nullptr);
  c->connected = true;
  return RMC_OK;
}

extern "C" int32_t rmc_comm_destroy(rmc_comm_t* c) {
  if (!c) return RMC_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->ipc_opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->local); cudaFree(c->arrive); cudaFree(c->arrive_td); cudaFree(c->verdict); cudaFree(c->g_nodes); cudaFree(c->g_td); cudaFree(c->g_pri);
  delete c;
  return RMC_OK;
}

/* 0 = healthy; otherwise the epoch of the step whose exchange timed out (a peer never published). Synchronises. */
extern "C" int32_t rmc_comm_status_sync(rmc_comm_t* c, uint32_t* timed_out_epoch, rmc_stream_t s) {
  if (!c || !timed_out_epoch) return fail(RMC_ERR_ARG, "rmc_comm_status_sync: null");
  if (int32_t e = use_device(c->device)) return e;
  RMC_CUDA(cudaMemcpyAsync(timed_out_epoch, c->local + 2 * kCommMaxWorld * sizeof(unsigned), sizeof(unsigned), cudaMemcpyDeviceToHost, as_stream(s)));
  RMC_CUDA(cudaStreamSynchronize(as_stream(s)));
  return RMC_OK;
}

static int32_t sharded_body(rmc_learner* l, rmc_replay* r, rmc_comm* c, const rmc_step_args_t* a, int32_t stages, cudaStream_t st);
extern "C" int32_t rmc_learner_step_sharded(rmc_learner_t* l, rmc_replay_t* r, rmc_comm_t* c, const rmc_step_args_t* a, int32_t stages, rmc_stream_t s) {
  if (!l || !r || !c || !a) return fail(RMC_ERR_ARG, "rmc_learner_step_sharded: null");
  if (!c->connected || c->learner != l) return fail(RMC_ERR_STATE, "rmc_learner_step_sharded: comm not connected to this learner");
  const long long Bg = a->global_batch > 0 ? a->global_batch : a->batch;
  if (Bg > c->global_batch_max || Bg < c->world) return fail(RMC_ERR_ARG, "rmc_learner_step_sharded: global batch outside the comm's range");
  long long lo, hi;
  shard_range_c(Bg, c->rank, c->world, &lo, &hi);
  if (a->shard_offset != lo || a->batch != hi - lo) return fail(RMC_ERR_ARG, "rmc_learner_step_sharded: batch/shard_offset differ from shard_range(global_batch, rank, world)");
  const int want = RMC_PH_FORWARD | RMC_PH_BACKWARD | RMC_PH_ADAM;
  if ((a->phases & want) != want) return fail(RMC_ERR_ARG, "rmc_learner_step_sharded: needs FORWARD, BACKWARD and ADAM");
  if (stages < 1 || stages > 3) return fail(RMC_ERR_ARG, "rmc_learner_step_sharded: stages must be 1, 2 or 3");
  cudaStream_t st = as_stream(s);
  if (stages & 1) c->epoch = (c->epoch >= 0x7fffffffu) ? 1u : c->epoch + 1u;      // once per step, outside the (re-traceable) body: ranks stay in lockstep
  // The tensor-core step is ~16 short kernels on two streams: replayed as ONE graph launch per step once the operand images
  // are current (re-trace and patch, see launch_pdl; the epoch / parity arguments of the exchange kernels are patched each step).
  // A kernel of the graph that waits for a peer's flag simply keeps the graph's later nodes waiting, like stream order would.
  const bool images = l->tct_ready && l->tc_packed_version == l->online_version && l->tc_bwd_version == l->online_version &&
                      l->tc_target_version == l->target_version;
  static const bool sharded_graph = [] { const char* e = std::getenv("RMC_SHARDED_GRAPH"); return !(e && e[0] == '0'); }();
  if (sharded_graph && stages == 3 && a->precision == RMC_PREC_BF16_TC && !l->hybrid && images && a->grads_in_dev == nullptr && !tc_events_on())
    return run_step_graph(l, a->batch, 0x10000000 | a->phases, st, [&](cudaStream_t s_) { return sharded_body(l, r, c, a, stages, s_); });
  return sharded_body(l, r, c, a, stages, st);
}

static int32_t sharded_body(rmc_learner* l, rmc_replay* r, rmc_comm* c, const rmc_step_args_t* a, int32_t stages, cudaStream_t st) {
  rmc_stream_t s = reinterpret_cast<rmc_stream_t>(st);
  const long long Bg = a->global_batch > 0 ? a->global_batch : a->batch;
  const bool per = l->spec.prioritized != 0 && (a->phases & RMC_PH_PRIORITY);
  const int parity = static_cast<int>(c->epoch & 1u);
  CommView V = c->view;
  for (int q = 0; q <= c->world; ++q) {
    long long qlo, qhi;
    shard_range_c(Bg, std::min(q, c->world - 1), c->world, &qlo, &qhi);
    V.shard_lo[q] = (q < c->world) ? qlo : qhi;
  }
  const long long n_local = a->batch;
  // tensor-core mode: |td| exists right after the TD kernel -> the (leaf, |td|) slices are exchanged early and the replicated
  // write-back runs on a side stream beside the backward pass and the gradient exchange
  const bool early = per && a->precision == RMC_PREC_BF16_TC && !l->hybrid;
  if (stages & 1) {
  // 1. local shard: sample (global strata), forward, TD, backward -> local gradient blob scaled by 1/B_global
  rmc_step_args_t a1 = *a;
  a1.phases = a->phases & (RMC_PH_SAMPLE | RMC_PH_FORWARD | RMC_PH_BACKWARD);
  a1.grads_in_dev = nullptr;
  if (early) { l->early.c = c; l->early.V = V; l->early.parity = parity; l->early.Bg = Bg; l->early.issue_side = (stages == 3); }
  l->no_graph = true;
  const int32_t e1 = rmc_learner_step(l, r, &a1, s);
  l->no_graph = false;
  l->early.c = nullptr;
  if (e1) return e1;
  // 2. publish: gradient blob, loss partial, (leaf, |td|) slice -> own exchange buffer, then one flag per rank
  const unsigned pub_blocks = std::max(1u, std::min(148u, blocks_for(std::max<long long>(l->L.total / 4, (per && !early) ? n_local : 0), 1024)));
  RMC_CUDA(launch_pdl(k_comm_publish, dim3(pub_blocks), dim3(256), 0, st, V, parity, c->epoch, l->ctx.grads, l->L.total, l->ctx.loss, (per && !early) ? l->ctx.nodes : nullptr, l->ctx.abs_td,
                                            n_local, c->arrive));
  RMC_KERNEL_OK();
  }
  if (!(stages & 2)) return RMC_OK;
  // 3. reduce over the ranks (peer loads, rank order) fused with Adam (+ Polyak)
  rmc_step_args_t a2 = *a;
  a2.phases = a->phases & (RMC_PH_ADAM | RMC_PH_POLYAK | RMC_PH_HARDSYNC);
  StepScalars S;
  if (int32_t e = fill_scalars(l, &a2, &S)) return e;
  l->epoch = (l->epoch >= 0x7fffffffu) ? 1u : l->epoch + 1u;
  S.epoch = l->epoch;
  if (early && stages != 3) {      // ranks emulated on one GPU: the side-stream part is issued here instead of right after TD
    l->early.c = c; l->early.V = V; l->early.parity = parity; l->early.Bg = Bg;
    const int32_t e2 = comm_side_writeback(l, r, st);
    l->early.c = nullptr;
    if (e2) return e2;
  }
  const int param_blocks = static_cast<int>(blocks_for(l->L.total, 256));
  const int gather_blocks = (per && !early) ? static_cast<int>(std::min<long long>(256, (Bg + 255) / 256)) : 0;
  const bool images = l->tct_ready && l->tc_packed_version == l->online_version && l->tc_bwd_version == l->online_version &&
                      l->tc_target_version == l->target_version;      // keep the bf16 operand images current (tensor-core mode)
  const TcPackOut P{images ? l->tc_packed : nullptr, images ? l->tc_packed_bwd : nullptr, images ? l->tc_packed_target : nullptr};
  RMC_CUDA(launch_pdl(k_comm_reduce_adam, dim3(param_blocks + gather_blocks), dim3(256), 0, st, l->ctx, S, V, parity, c->epoch, param_blocks, c->g_nodes, c->g_td, (per && !early) ? 1 : 0, P));
  RMC_KERNEL_OK();
  l->loss_epoch = S.epoch;
  ++l->online_version;
  if (a->phases & (RMC_PH_POLYAK | RMC_PH_HARDSYNC)) ++l->target_version;
  if (images) {
    l->tc_packed_version = l->tc_bwd_version = l->online_version;
    if (a->phases & (RMC_PH_POLYAK | RMC_PH_HARDSYNC)) l->tc_target_version = l->target_version;
  }
  // 4. PER: the full write-back of the GLOBAL batch on every replica, in global batch order (trees stay identical)
  if (early) {
    if (!(g_trace != nullptr && g_trace->mode == 2)) RMC_CUDA(cudaStreamWaitEvent(st, l->tc_ev[1], 0));      // the side-stream write-back joins here
  } else if (per) {
    if (Bg <= kTreeCtaMax) {
      RMC_CUDA(launch_pdl(k_tree_update_small, dim3(1), dim3(kThreads), 0, st, r->dev, c->g_nodes, nullptr, c->g_td, c->g_pri, Bg, static_cast<float>(l->hyper.per_eps),
                                                 static_cast<float>(l->hyper.per_alpha), static_cast<float>(l->hyper.per_pmax)));
      RMC_KERNEL_OK();
    } else {
      RMC_CUDA(launch_pdl(k_td_to_pri, dim3(blocks_for(Bg, 256)), dim3(256), 0, st, c->g_td, c->g_pri, Bg, static_cast<float>(l->hyper.per_eps), static_cast<float>(l->hyper.per_alpha),
                                                      static_cast<float>(l->hyper.per_pmax)));
      RMC_KERNEL_OK();
      if (int32_t e = tree_update_large(r, c->g_nodes, c->g_pri, Bg, false, st)) return e;
    }
  }
  return RMC_OK;
}

// store_transitions of the trainer's per-env-step rows + the learner step, issued back to back from ONE host call
// (train.py:91-101 calls them in this order; the step kernel then follows the push kernel by one launch latency instead
// of a host round trip through the interpreter).
extern "C" int32_t rmc_learner_step_push(rmc_learner_t* l, rmc_replay_t* r, const rmc_step_args_t* a, const float* obs_host, const int64_t* act_host,
                                         const float* rew_host, const float* done_host, const float* next_obs_host, int64_t n, rmc_stream_t s) {
  if (n > 0)
    if (int32_t e = push_impl(r, obs_host, act_host, rew_host, done_host, next_obs_host, n, true, as_stream(s))) return e;
  return rmc_learner_step(l, r, a, s);
}

extern "C" int32_t rmc_learner_output(rmc_learner_t* l, const char* name, void** dev_ptr, int64_t* n_elems) {
  if (!l || !name || !dev_ptr || !n_elems) return fail(RMC_ERR_ARG, "rmc_learner_output: null");
  const AgentCtx& c = l->ctx;
  const long long B = l->last_batch;
  struct Ent { const char* n; void* p; long long cnt; };
  const Ent table[] = {{"nodes", c.nodes, B}, {"is_w", c.is_w, B}, {"q_sa", c.q_sa, B}, {"y", c.y, B}, {"abs_td", c.abs_td, B},
                       {"huber", c.hub, B}, {"pri", c.pri, B}, {"gcoef", c.gcoef, B}, {"loss", c.loss, 1},
                       {"q_next_tgt", c.QT, B * kQLD}, {"q_next_on", c.QN, B * kQLD}, {"q", c.Q, B * kQLD},
                       {"rows", c.X, B * l->rf}, {"h1", c.H1, B * kH1}, {"h2", c.H2, B * kH2}, {"dz1", c.DZ1, B * kH1},
                       {"dz2", c.DZ2, B * kH2}, {"dh", c.DH, B * kQLD}, {"leaf_p", c.leaf_p, B},
                       {"grads_blob", c.grads, l->L.total}, {"loss_part", c.loss_part, 1024}};
  for (const Ent& e : table)
    if (std::strcmp(e.n, name) == 0) {
      *dev_ptr = e.p;
      *n_elems = e.cnt;
      return RMC_OK;
    }
  return fail(RMC_ERR_ARG, std::string("rmc_learner_output: unknown output ") + name);
}

// 0, or RMC_ERR_STATE when an in-kernel spin of one of this learner's launches timed out (device-side watchdog, SpinGuard)
static int32_t step_health(const rmc_learner* l) {
  if (l->host_loss == nullptr) return RMC_OK;
  const float bits = l->host_loss[2];
  unsigned bad = 0;
  std::memcpy(&bad, &bits, sizeof(bad));
  if (bad == 0) return RMC_OK;
  return fail(RMC_ERR_STATE, "fused learner step: an in-kernel wait timed out (launch epoch " + std::to_string(bad & 0x7fffffffu) +
                             "): the grid was not co-resident -- another process is sharing this GPU; set RMC_LAUNCH=coop");
}
extern "C" int32_t rmc_learner_status(rmc_learner_t* l, uint32_t* failed_epoch) {
  if (!l) return fail(RMC_ERR_ARG, "rmc_learner_status: null");
  unsigned bad = 0;
  if (l->host_loss != nullptr) { const float bits = l->host_loss[2]; std::memcpy(&bad, &bits, sizeof(bad)); }
  if (failed_epoch) *failed_epoch = bad;
  return bad ? step_health(l) : RMC_OK;
}

extern "C" int32_t rmc_learner_loss_sync(rmc_learner_t* l, float* out_host, rmc_stream_t s) {
  if (!l || !out_host) return fail(RMC_ERR_ARG, "rmc_learner_loss_sync: null");
  if (int32_t e = use_device(l->device)) return e;
  if (int32_t e = step_health(l)) return e;
  if (l->host_loss != nullptr && l->loss_epoch != 0) {
    // the kernel stores {loss, epoch} as one 8-byte word into mapped host memory: wait for this launch's epoch
    const unsigned want = l->loss_epoch;
    const volatile unsigned long long* word = reinterpret_cast<const volatile unsigned long long*>(l->host_loss);
    auto take = [&](unsigned long long w) {
      const unsigned lo = static_cast<unsigned>(w & 0xffffffffull);
      std::memcpy(out_host, &lo, sizeof(float));
    };
    for (long long spin = 0; spin < 200000000ll; ++spin) {
      const unsigned long long w = *word;
      if (static_cast<unsigned>(w >> 32) == want) { take(w); return step_health(l); }
      if ((spin & 0xfff) == 0xfff && cudaStreamQuery(as_stream(s)) != cudaErrorNotReady) break;   // finished or faulted
    }
    RMC_CUDA(cudaStreamSynchronize(as_stream(s)));
    const unsigned long long w = *word;
    if (static_cast<unsigned>(w >> 32) == want) { take(w); return RMC_OK; }
  }
  RMC_CUDA(cudaMemcpyAsync(out_host, l->ctx.loss, sizeof(float), cudaMemcpyDeviceToHost, as_stream(s)));
  RMC_CUDA(cudaStreamSynchronize(as_stream(s)));
  return RMC_OK;
}

static int32_t infer_launch(rmc_learner* l, const float* params, const float* obs_dev, long long n, long long* actions, float* q, int mode,
                            cudaStream_t st) {
  if (l->hybrid) return hybrid_infer(l, params, obs_dev, n, actions, q, mode, st);
  // large batches: 64-row tiles (every weight fetched from shared memory serves 64 rows; rmc_infer64.cuh)
  const int big_bytes = big_smem_floats(l->L.total) * 4;
  if (n >= 1024 && l->L.D <= 16 && big_bytes <= l->max_smem_optin) {
    if (!l->big_ready) {
      RMC_CUDA(cudaFuncSetAttribute(k_mlp_infer64, cudaFuncAttributeMaxDynamicSharedMemorySize, big_bytes));
      l->big_ready = true;
    }
    const long long big_tiles = (n + kBigRows - 1) / kBigRows;
    k_mlp_infer64<<<static_cast<unsigned>(std::min<long long>(big_tiles, l->num_sms)), kBigThreads, static_cast<size_t>(big_bytes), st>>>(
        l->L, params, obs_dev, n, actions, q, mode);
    RMC_KERNEL_OK();
    return RMC_OK;
  }
  const long long n_tiles = (n + kR - 1) / kR;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(n_tiles, l->num_sms));
  k_mlp_infer<<<grid, kThreads, static_cast<size_t>(l->smem_bytes), st>>>(l->L, params, obs_dev, n, actions, q, mode);
  RMC_KERNEL_OK();
  return RMC_OK;
}

extern "C" int32_t rmc_learner_q_values(rmc_learner_t* l, int32_t which, const float* obs_dev, int64_t n, float* q_out_dev, rmc_stream_t s) {
  if (!l || !obs_dev || !q_out_dev || n < 1 || (which != RMC_ONLINE && which != RMC_TARGET)) return fail(RMC_ERR_ARG, "rmc_learner_q_values: bad args");
  if (int32_t e = use_device(l->device)) return e;
  return infer_launch(l, l->blobs[which], obs_dev, n, nullptr, q_out_dev, 1, as_stream(s));
}

extern "C" int32_t rmc_learner_heads(rmc_learner_t* l, int32_t which, const float* obs_dev, int64_t n, float* heads_out_dev, rmc_stream_t s) {
  if (!l || !obs_dev || !heads_out_dev || n < 1 || (which != RMC_ONLINE && which != RMC_TARGET)) return fail(RMC_ERR_ARG, "rmc_learner_heads: bad args");
  if (int32_t e = use_device(l->device)) return e;
  return infer_launch(l, l->blobs[which], obs_dev, n, nullptr, heads_out_dev, 2, as_stream(s));
}

extern "C" int32_t rmc_learner_act(rmc_learner_t* l, const float* obs_dev, int64_t n, int64_t* actions_dev, rmc_stream_t s) {
  if (!l || !obs_dev || !actions_dev || n < 1) return fail(RMC_ERR_ARG, "rmc_learner_act: bad args");
  if (int32_t e = use_device(l->device)) return e;
  return infer_launch(l, l->blobs[RMC_ONLINE], obs_dev, n, reinterpret_cast<long long*>(actions_dev), nullptr, 0, as_stream(s));
}

// tensor-core (tcgen05, bf16 operands / fp32 accumulate) mode of act / heads: looser, stated bound (see rmc_tc.cuh)
static int32_t infer_tc(rmc_learner* l, const float* obs_dev, long long n, long long* actions, float* heads, int mode, cudaStream_t st) {
  if (l->hybrid) return fail(RMC_ERR_UNSUPPORTED, "tensor-core act mode: built for the macro MLP only");
  if (l->L.D > kTcK1) return fail(RMC_ERR_UNSUPPORTED, "tensor-core act mode: obs_dim must be <= 16");
  if (l->tc_packed == nullptr) {
    if (int32_t e = owned_alloc(l, &l->tc_packed, static_cast<size_t>(kTcBlobBytes))) return e;
    RMC_CUDA(cudaFuncSetAttribute(k_mlp_infer_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  }
  if (l->tc_packed_version != l->online_version) {
    RMC_CUDA(launch_pdl(k_tc_pack, dim3(blocks_for(kH2 * kH1, 256)), dim3(256), 0, st, l->blobs[RMC_ONLINE], l->L, l->tc_packed));
    RMC_KERNEL_OK();
    l->tc_packed_version = l->online_version;
  }
  const long long n_tiles = (n + kTcRows - 1) / kTcRows;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(n_tiles, l->num_sms));
  TcFwdExtra ex{};
  ex.row_stride = l->L.D;
  ex.act = l->L.act;
  if (std::getenv("RMC_TC_FWD_DBG") != nullptr) ex.dbg = reinterpret_cast<long long*>(l->dbg_buf);   // diagnostics: stage clocks of CTA 0
  k_mlp_infer_tc<<<grid, kTcFwdThreads, kTcSmemBytes, st>>>(l->tc_packed, l->L.D, l->L.A, l->L.NH, l->L.dueling, obs_dev, n, actions, heads, mode, ex);
  RMC_KERNEL_OK();
  return RMC_OK;
}
extern "C" int32_t rmc_learner_act_tc(rmc_learner_t* l, const float* obs_dev, int64_t n, int64_t* actions_dev, rmc_stream_t s) {
  if (!l || !obs_dev || !actions_dev || n < 1) return fail(RMC_ERR_ARG, "rmc_learner_act_tc: bad args");
  if (int32_t e = use_device(l->device)) return e;
  return infer_tc(l, obs_dev, n, reinterpret_cast<long long*>(actions_dev), nullptr, 0, as_stream(s));
}
extern "C" int32_t rmc_learner_heads_tc(rmc_learner_t* l, const float* obs_dev, int64_t n, float* heads_out_dev, rmc_stream_t s) {
  if (!l || !obs_dev || !heads_out_dev || n < 1) return fail(RMC_ERR_ARG, "rmc_learner_heads_tc: bad args");
  if (int32_t e = use_device(l->device)) return e;
  return infer_tc(l, obs_dev, n, nullptr, heads_out_dev, 2, as_stream(s));
}

static int32_t act_host_impl(rmc_learner_t* l, const float* obs_host, int64_t n, int64_t* actions_host, float eps, uint64_t seed, uint64_t counter, rmc_stream_t s);
extern "C" int32_t rmc_learner_act_host_sync(rmc_learner_t* l, const float* obs_host, int64_t n, int64_t* actions_host, rmc_stream_t s) {
  return act_host_impl(l, obs_host, n, actions_host, -1.f, 0, 0, s);
}
extern "C" int32_t rmc_learner_act_eps_host_sync(rmc_learner_t* l, const float* obs_host, int64_t n, int64_t* actions_host, float epsilon,
                                                 uint64_t seed, uint64_t counter, rmc_stream_t s) {
  if (!(epsilon >= 0.f && epsilon <= 1.f)) return fail(RMC_ERR_ARG, "rmc_learner_act_eps_host_sync: epsilon outside [0, 1]");
  return act_host_impl(l, obs_host, n, actions_host, epsilon, seed, counter, s);
}
static int32_t act_host_impl(rmc_learner_t* l, const float* obs_host, int64_t n, int64_t* actions_host, float eps, uint64_t seed, uint64_t counter, rmc_stream_t s) {
  if (!l || !obs_host || !actions_host || n < 1) return fail(RMC_ERR_ARG, "rmc_learner_act_host_sync: bad args");
  if (int32_t e = use_device(l->device)) return e;
  cudaStream_t st = as_stream(s);
  if (n <= kActTinyMax && !l->hybrid && l->L.D <= 16 && l->act_map_state >= 0) {
    if (l->act_map_state == 0) {
      long long* hp = nullptr; long long* dp = nullptr;
      if (cudaHostAlloc(reinterpret_cast<void**>(&hp), (kActTinyMax + 1) * sizeof(long long), cudaHostAllocMapped) == cudaSuccess &&
          cudaHostGetDevicePointer(reinterpret_cast<void**>(&dp), hp, 0) == cudaSuccess && dev_alloc(&l->act_ctr, 1) == RMC_OK &&
          cudaFuncSetAttribute(k_act_tiny, cudaFuncAttributeMaxDynamicSharedMemorySize, l->smem_bytes) == cudaSuccess) {
        std::memset(hp, 0, (kActTinyMax + 1) * sizeof(long long));
        l->act_map_host = hp; l->act_map_dev = dp; l->act_map_state = 1;
      } else {
        cudaGetLastError();
        if (hp) cudaFreeHost(hp);
        l->act_map_state = -1;
      }
    }
    if (l->act_map_state == 1) {
      ActTinyObs X;
      std::memcpy(X.v, obs_host, static_cast<size_t>(n) * l->L.D * sizeof(float));
      l->act_epoch = (l->act_epoch >= 0x7fffffffu) ? 1u : l->act_epoch + 1u;
      const unsigned want = l->act_epoch;
      volatile unsigned* host_epoch = reinterpret_cast<volatile unsigned*>(l->act_map_host + kActTinyMax);
      k_act_tiny<<<blocks_for(n, kR), kThreads, static_cast<size_t>(l->smem_bytes), st>>>(
          l->L, l->blobs[RMC_ONLINE], X, static_cast<int>(n), eps, seed, counter, l->act_map_dev,
          reinterpret_cast<unsigned*>(l->act_map_dev + kActTinyMax), l->act_ctr, want);
      RMC_KERNEL_OK();
      bool got = false;
      for (long long spin = 0; spin < 200000000ll && !got; ++spin) {
        got = (*host_epoch == want);
        if (!got && (spin & 0xfff) == 0xfff && cudaStreamQuery(st) != cudaErrorNotReady) break;   // finished or faulted
      }
      if (!got) {
        RMC_CUDA(cudaStreamSynchronize(st));
        if (*host_epoch != want) return fail(RMC_ERR_CUDA, "rmc_learner_act_host_sync: the act kernel did not publish its result");
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      for (int64_t i = 0; i < n; ++i) actions_host[i] = l->act_map_host[i];
      return RMC_OK;
    }
  }
  if (n > l->act_cap) {
    RMC_CUDA(cudaStreamSynchronize(st));
    if (l->act_pin_obs) cudaFreeHost(l->act_pin_obs);
    if (l->act_pin_out) cudaFreeHost(l->act_pin_out);
    cudaFree(l->act_dev_obs);
    cudaFree(l->act_dev_out);
    const long long cap = std::max<long long>(n, 1024);
    RMC_CUDA(cudaMallocHost(reinterpret_cast<void**>(&l->act_pin_obs), static_cast<size_t>(cap) * l->L.D * sizeof(float)));
    RMC_CUDA(cudaMallocHost(reinterpret_cast<void**>(&l->act_pin_out), static_cast<size_t>(cap) * sizeof(long long)));
    RMC_CUDA(cudaMalloc(reinterpret_cast<void**>(&l->act_dev_obs), static_cast<size_t>(cap) * l->L.D * sizeof(float)));
    RMC_CUDA(cudaMalloc(reinterpret_cast<void**>(&l->act_dev_out), static_cast<size_t>(cap) * sizeof(long long)));
    l->act_cap = cap;
  }
  std::memcpy(l->act_pin_obs, obs_host, static_cast<size_t>(n) * l->L.D * sizeof(float));
  RMC_CUDA(cudaMemcpyAsync(l->act_dev_obs, l->act_pin_obs, static_cast<size_t>(n) * l->L.D * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int32_t e = infer_launch(l, l->blobs[RMC_ONLINE], l->act_dev_obs, n, l->act_dev_out, nullptr, 0, st)) return e;
  if (eps >= 0.f) {
    k_eps_greedy<<<blocks_for(n, 256), 256, 0, st>>>(l->act_dev_out, n, eps, l->spec.n_actions, seed, counter);
    RMC_KERNEL_OK();
  }
  RMC_CUDA(cudaMemcpyAsync(l->act_pin_out, l->act_dev_out, static_cast<size_t>(n) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  RMC_CUDA(cudaStreamSynchronize(st));
  std::memcpy(actions_host, l->act_pin_out, static_cast<size_t>(n) * sizeof(long long));
  return RMC_OK;
}

// ------------------------------------------------------------------------------ groups (ensembles)
extern "C" int32_t rmc_group_create(rmc_group_t** out, rmc_learner_t* const* learners, rmc_replay_t* const* replays, int32_t n_agents) {
  if (!out || !learners || !replays || n_agents < 1) return fail(RMC_ERR_ARG, "rmc_group_create: bad args");
  rmc_learner* l0 = learners[0];
  if (n_agents > l0->num_sms) return fail(RMC_ERR_ARG, "rmc_group_create: more agents than SMs");
  for (int i = 0; i < n_agents; ++i) {
    rmc_learner* l = learners[i];
    rmc_replay* r = replays[i];
    if (!l || !r) return fail(RMC_ERR_ARG, "rmc_group_create: null member");
    if (l->hybrid) return fail(RMC_ERR_UNSUPPORTED, "rmc_group_create: ensemble launches are built for the macro MLP only");
    if (std::memcmp(&l->spec, &l0->spec, sizeof(rmc_net_spec_t)) != 0 || l->max_batch != l0->max_batch || l->device != l0->device)
      return fail(RMC_ERR_ARG, "rmc_group_create: members must share spec, max_batch and device");
    if (std::memcmp(&l->hyper, &l0->hyper, sizeof(rmc_hyper_t)) != 0)      // one launch carries ONE set of scalars (lr, gamma, tau, PER constants)
      return fail(RMC_ERR_ARG, "rmc_group_create: members must share the hyper-parameters (rmc_hyper_t)");
    if (r->D != l->spec.obs_dim || (r->prioritized != 0) != (l->spec.prioritized != 0) || r->device != l->device)
      return fail(RMC_ERR_ARG, "rmc_group_create: replay/learner mismatch");
  }
  if (int32_t e = use_device(l0->device)) return e;
  auto* g = new rmc_group();
  g->n = n_agents;
  g->learners.assign(learners, learners + n_agents);
  g->replays.assign(replays, replays + n_agents);
  int32_t e = RMC_OK;
  if ((e = dev_alloc(&g->ctx_dev, static_cast<size_t>(n_agents), false))) return e;
  if ((e = dev_alloc(&g->barriers, static_cast<size_t>(n_agents)))) return e;
  if ((e = dev_alloc(&g->qt_flags, static_cast<size_t>(n_agents) * kFlagWords))) return e;
  std::vector<AgentCtx> h(n_agents);
  for (int i = 0; i < n_agents; ++i) {
    h[i] = learners[i]->ctx;
    h[i].rp = replays[i]->dev;
    h[i].barrier = g->barriers + i;
    h[i].qt_flag = g->qt_flags + static_cast<size_t>(i) * kFlagWords;
  }
  RMC_CUDA(cudaMemcpy(g->ctx_dev, h.data(), sizeof(AgentCtx) * n_agents, cudaMemcpyHostToDevice));
  *out = g;
  return RMC_OK;
}

// Agent.store_transitions of every member for one env step (dqn/agent.py:70-73), n rows each: ONE launch (block = member).
// Falls back to one small push per member when the packed rows do not fit the kernel-argument buffer.
extern "C" int32_t rmc_group_push_host(rmc_group_t* g, const float* const* obs_host, const int64_t* const* act_host, const float* const* rew_host,
                                       const float* const* done_host, const float* const* next_obs_host, int64_t n, rmc_stream_t s) {
  if (!g || !obs_host || !act_host || !rew_host || !done_host || !next_obs_host || n < 0) return fail(RMC_ERR_ARG, "rmc_group_push_host: bad args");
  if (n == 0) return RMC_OK;
  rmc_replay* r0 = g->replays[0];
  if (int32_t e = use_device(r0->device)) return e;
  cudaStream_t st = as_stream(s);
  const long long per_member = n * r0->rf;
  bool one_launch = n <= 8 && per_member * g->n <= kGroupRowFloats;
  for (int i = 0; i < g->n && one_launch; ++i) one_launch = g->replays[i]->rf == r0->rf && n <= g->replays[i]->cap;
  if (!one_launch) {
    for (int i = 0; i < g->n; ++i)
      if (int32_t e = push_impl(g->replays[i], obs_host[i], act_host[i], rew_host[i], done_host[i], next_obs_host[i], n, true, st)) return e;
    return RMC_OK;
  }
  GroupRows rows;
  for (int i = 0; i < g->n; ++i)
    pack_rows_host(rows.v + i * per_member, obs_host[i], act_host[i], rew_host[i], done_host[i], next_obs_host[i], n, r0->D, r0->rf);
  k_push_tiny_group<<<g->n, 32, 0, st>>>(g->ctx_dev, rows, static_cast<int>(n), r0->rf, 1.0f);
  RMC_KERNEL_OK();
  for (int i = 0; i < g->n; ++i) {
    rmc_replay* r = g->replays[i];
    r->dp = (r->dp + n) % r->cap;
    r->size = std::min<long long>(r->size + n, r->cap);
  }
  return RMC_OK;
}

extern "C" int32_t rmc_group_destroy(rmc_group_t* g) {
  if (!g) return RMC_OK;
  cudaSetDevice(g->learners[0]->device);
  cudaDeviceSynchronize();
  cudaFree(g->ctx_dev);
  cudaFree(g->barriers);
  cudaFree(g->qt_flags);
  for (auto& st : g->tc_streams) if (st) cudaStreamDestroy(st);
  for (auto& ev : g->tc_join) if (ev) cudaEventDestroy(ev);
  if (g->tc_fork) cudaEventDestroy(g->tc_fork);
  delete g;
  return RMC_OK;
}

extern "C" int32_t rmc_group_step(rmc_group_t* g, const rmc_step_args_t* a, rmc_stream_t s) {
  if (!g || !a) return fail(RMC_ERR_ARG, "rmc_group_step: null");
  rmc_learner* l0 = g->learners[0];
  for (int i = 0; i < g->n; ++i) {
    if (int32_t e = check_step(g->learners[i], g->replays[i], a)) return e;
    if (std::memcmp(&g->learners[i]->hyper, &l0->hyper, sizeof(rmc_hyper_t)) != 0)
      return fail(RMC_ERR_ARG, "rmc_group_step: a member's hyper-parameters were changed after rmc_group_create");
  }
  if (int32_t e = use_device(l0->device)) return e;
  cudaStream_t st = as_stream(s);
  if (a->precision == RMC_PREC_BF16_TC) {
    // Tensor-core mode of the ensemble (north star: "tensor cores ... for the large-batch and agent-ensemble configs"): every
    // member runs its own tcgen05 step (step_tc: per-member bf16 operand images, one graph launch per member once they are
    // current) and the members run SIDE BY SIDE, one internal stream each, forked from and joined back into the caller's
    // stream.  Dense GEMMs need >= one 128-row UMMA tile per CTA: this mode pays for ensembles of large-batch members
    // (B >= ~2048 each); at B = 256 per member the fused fp32 launch above is the faster path and stays the default.
    if (g->tc_streams.empty()) {
      g->tc_streams.resize(g->n); g->tc_join.resize(g->n);
      for (int i = 0; i < g->n; ++i) {
        RMC_CUDA(cudaStreamCreateWithFlags(&g->tc_streams[i], cudaStreamNonBlocking));
        RMC_CUDA(cudaEventCreateWithFlags(&g->tc_join[i], cudaEventDisableTiming));
      }
      RMC_CUDA(cudaEventCreateWithFlags(&g->tc_fork, cudaEventDisableTiming));
    }
    RMC_CUDA(cudaEventRecord(g->tc_fork, st));
    for (int i = 0; i < g->n; ++i) {
      rmc_step_args_t ai = *a;
      if (a->u_dev) ai.u_dev = a->u_dev + static_cast<size_t>(i) * a->batch;
      if (a->idx_dev) ai.idx_dev = a->idx_dev + static_cast<size_t>(i) * a->batch;
      ai.seed = a->seed + 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>(i);      // independent sampling streams per member
      RMC_CUDA(cudaStreamWaitEvent(g->tc_streams[i], g->tc_fork, 0));
      if (int32_t e = rmc_learner_step(g->learners[i], g->replays[i], &ai, reinterpret_cast<rmc_stream_t>(g->tc_streams[i]))) return e;
      RMC_CUDA(cudaEventRecord(g->tc_join[i], g->tc_streams[i]));
      RMC_CUDA(cudaStreamWaitEvent(st, g->tc_join[i], 0));
    }
    return RMC_OK;
  }
  if (a->batch > kTreeCtaMax && l0->spec.prioritized && (a->phases & RMC_PH_PRIORITY))
    return fail(RMC_ERR_UNSUPPORTED, "rmc_group_step: PER batches above 4096 are stepped per agent");
  StepScalars S;
  if (int32_t e = fill_scalars(l0, a, &S)) return e;
  const int per_agent_max = std::max(1, l0->num_sms / g->n);
  const int G = grid_for(l0, a->batch, per_agent_max);
  const long long n_tiles = (a->batch + kTM - 1) / kTM;
  S.n_row_ctas = static_cast<int>(std::min<long long>(G, n_tiles));
  const bool rows = (a->phases & (RMC_PH_SAMPLE | RMC_PH_FORWARD)) != 0;
  const bool phase_b = (a->phases & (RMC_PH_PRIORITY | RMC_PH_BACKWARD | RMC_PH_ADAM | RMC_PH_POLYAK | RMC_PH_HARDSYNC)) != 0;
  if (rows && phase_b) S.barrier_target = g->barrier_count + static_cast<unsigned>(G);
  g->epoch = (g->epoch >= 0x7fffffffu) ? 1u : g->epoch + 1u;
  S.epoch = 0x80000000u | g->epoch;      // group launches and single-agent launches never share an epoch value
  for (int i = 0; i < g->n; ++i) {
    g->learners[i]->last_batch = a->batch;
    if (a->phases & RMC_PH_ADAM) ++g->learners[i]->online_version;
    if (a->phases & (RMC_PH_POLYAK | RMC_PH_HARDSYNC)) ++g->learners[i]->target_version;
    if ((a->phases & RMC_PH_FORWARD) && phase_b) g->learners[i]->loss_epoch = S.epoch;
  }
  AgentCtx single = l0->ctx;
  const AgentCtx* many = g->ctx_dev;
  void* args[] = {&single, &many, &S};
  if (int32_t e = launch_step(l0->device, dim3(G, g->n, 1), args, static_cast<size_t>(l0->smem_bytes), st, step_path(S, n_tiles, l0->ctx))) return e;
  if (rows && phase_b) g->barrier_count = S.barrier_target;
  return RMC_OK;
}

// ------------------------------------------------------------------------------ debug timing
// Kernel-span recorder (see rmc_device.cuh): enable = 1 allocates / resets the table and points the kernels at it, 0
// detaches it.  rmc_debug_spans_read_sync copies {first CTA start, last CTA end} (ns, %globaltimer) of the 64 slots and
// resets them; names_out (optional) receives the slot names, comma separated.
static unsigned long long* g_span_dev[kMaxDevices] = {nullptr};
extern "C" int32_t rmc_debug_spans(int32_t device, int32_t enable) {
  if (device < 0 || device >= kMaxDevices) return fail(RMC_ERR_ARG, "rmc_debug_spans: device");
  if (int32_t e = use_device(device)) return e;
  RMC_CUDA(cudaDeviceSynchronize());
  unsigned long long* ptr = nullptr;
  if (enable) {
    if (g_span_dev[device] == nullptr) RMC_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_span_dev[device]), 128 * sizeof(unsigned long long)));
    std::vector<unsigned long long> init(128);
    for (int k = 0; k < 64; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
    RMC_CUDA(cudaMemcpy(g_span_dev[device], init.data(), 128 * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    ptr = g_span_dev[device];
  }
  RMC_CUDA(cudaMemcpyToSymbol(g_span_table, &ptr, sizeof(ptr)));
  return RMC_OK;
}
extern "C" int32_t rmc_debug_spans_read_sync(int32_t device, uint64_t* out128_host, char* names_out, int32_t names_cap) {
  if (device < 0 || device >= kMaxDevices || !out128_host || g_span_dev[device] == nullptr) return fail(RMC_ERR_ARG, "rmc_debug_spans_read_sync: not enabled");
  if (int32_t e = use_device(device)) return e;
  RMC_CUDA(cudaDeviceSynchronize());
  RMC_CUDA(cudaMemcpy(out128_host, g_span_dev[device], 128 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  std::vector<unsigned long long> init(128);
  for (int k = 0; k < 64; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
  RMC_CUDA(cudaMemcpy(g_span_dev[device], init.data(), 128 * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  if (names_out && names_cap > 0) {
    const char* names = "sample,fwd3,td,bwd,reduce_adam,td_to_pri,tree_stamp,tree_apply,tree_top,extremes,publish_td,gather_td,publish,comm_reduce,tree_small,uniform,pack";
    std::strncpy(names_out, names, static_cast<size_t>(names_cap) - 1);
    names_out[names_cap - 1] = 0;
  }
  return RMC_OK;
}

extern "C" int32_t rmc_learner_debug_timing(rmc_learner_t* l, int32_t enable) {
  if (!l) return fail(RMC_ERR_ARG, "rmc_learner_debug_timing: null");
  l->ctx.dbg = enable ? l->dbg_buf : nullptr;
  if (enable) {   // [min start, max end] slots of the launch-gap diagnostic
    std::vector<unsigned long long> init(128);
    for (int k = 0; k < 64; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
    RMC_CUDA(cudaMemcpy(l->dbg_buf + kDbgCtas * kDbgSlots, init.data(), 128 * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  }
  return RMC_OK;
}
extern "C" int32_t rmc_learner_debug_gaps_sync(rmc_learner_t* l, uint64_t* out128_host, rmc_stream_t s) {
  if (!l || !out128_host) return fail(RMC_ERR_ARG, "rmc_learner_debug_gaps_sync: null");
  if (int32_t e = use_device(l->device)) return e;
  RMC_CUDA(cudaMemcpyAsync(out128_host, l->dbg_buf + kDbgCtas * kDbgSlots, 128 * sizeof(uint64_t), cudaMemcpyDeviceToHost, as_stream(s)));
  RMC_CUDA(cudaStreamSynchronize(as_stream(s)));
  return RMC_OK;
}
extern "C" int32_t rmc_learner_debug_read_sync(rmc_learner_t* l, uint64_t* out_host, int32_t max_ctas, int32_t* n_ctas, rmc_stream_t s) {
  if (!l || !out_host || !n_ctas) return fail(RMC_ERR_ARG, "rmc_learner_debug_read_sync: null");
  if (int32_t e = use_device(l->device)) return e;
  const int n = std::min(max_ctas, l->last_grid);
  RMC_CUDA(cudaMemcpyAsync(out_host, l->dbg_buf, static_cast<size_t>(n) * kDbgSlots * sizeof(uint64_t), cudaMemcpyDeviceToHost, as_stream(s)));
  RMC_CUDA(cudaStreamSynchronize(as_stream(s)));
  *n_ctas = n;
  return RMC_OK;
}
