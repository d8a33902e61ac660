/*
 * rmc_b200.h -- C ABI of the B200-native DQN learner hot path (librmc_b200.so).
 *
 * This is the drop-in boundary for ONE path of youcefMehamlia/Multimodal-DRL-RMC: the
 * replay-minibatch update that dqn/agent.py runs every train step, plus batched greedy
 * action selection.  The reference has no FFI of its own (it is pure Python); each entry
 * point below names the reference interface it replaces (paths relative to the reference
 * root).  The Python mirror of the reference classes (multimodal_drl_rmc_b200/) binds these
 * symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain C types only; no torch / CUDA types in any signature.  `rmc_stream_t` is a
 *    cudaStream_t passed as void* (NULL = legacy default stream).
 *  - every function returns 0 on success, a negative rmc_status on failure; the message is
 *    available from rmc_last_error() (thread local).  Nothing throws, nothing calls exit().
 *  - pointers named *_dev are device pointers owned by the caller (e.g. torch tensors);
 *    *_host are host pointers.  Functions never synchronise unless their name ends in
 *    `_sync` or their documentation says so (host-buffer variants copy through an internal
 *    pinned staging ring and are asynchronous w.r.t. the device as far as CUDA allows).
 *  - one handle is used from one host thread at a time (the reference is single threaded).
 *  - there is NO CPU fallback: every compute entry fails with RMC_ERR_CUDA if no sm_100
 *    device is usable.
 */
#ifndef RMC_B200_H
#define RMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMC_ABI_VERSION 3

typedef void* rmc_stream_t;
typedef struct rmc_replay rmc_replay_t;   /* GPU-resident ring buffer (+ sum tree when prioritized) */
typedef struct rmc_learner rmc_learner_t; /* online/target nets, Adam state, scratch                */
typedef struct rmc_group rmc_group_t;     /* N (replay, learner) pairs stepped by ONE launch        */
typedef struct rmc_comm rmc_comm_t;       /* peer-memory exchange buffers of the sharded learner    */

enum rmc_status {
  RMC_OK = 0,
  RMC_ERR_ARG = -1,      /* bad argument / unsupported configuration */
  RMC_ERR_CUDA = -2,     /* CUDA runtime error (message has the cudaError string) */
  RMC_ERR_STATE = -3,    /* call not valid in the handle's current state */
  RMC_ERR_UNSUPPORTED = -4
};

enum rmc_param_kind { RMC_ONLINE = 0, RMC_TARGET = 1, RMC_ADAM_M = 2, RMC_ADAM_V = 3, RMC_GRADS = 4 };
enum rmc_activation { RMC_ACT_RELU = 0, RMC_ACT_ELU = 1 /* nn.ELU(alpha=1): env/dqn_config.py:175 */ };

/* Network + algorithm shape.  Mirrors what dqn/network.py:12-19 obtains from the user's
 * nn_conf_func and what dqn/agent.py:275-320 wires per agent flavour. */
typedef struct {
  int32_t obs_dim;     /* D  (14 macro-lane, 8 macro-no-lane)                      */
  int32_t hidden1;     /* 256 (only value built in this round)                     */
  int32_t hidden2;     /* 128 (only value built in this round)                     */
  int32_t n_actions;   /* A  (8)                                                   */
  int32_t dueling;     /* 1: fc_val + fc_adv (network.py:77-117); 0: fc_out (:50-74) */
  int32_t double_dqn;  /* 1: agent.py:209-216 ; 0: agent.py:171-175 (max over target) */
  int32_t prioritized; /* 1: IS-weighted loss + |td| write-back (agent.py:245-272)  */
  int32_t activation;  /* rmc_activation                                           */
} rmc_net_spec_t;

/* Hyper-parameters: env/dqn_config.py:26-56 values + the constants that are implicit in
 * torch.optim.Adam / nn.SmoothL1Loss and hard-coded in dqn/replay_memory.py:49-54. */
typedef struct {
  double lr;             /* 1e-4  */
  double adam_beta1;     /* 0.9   */
  double adam_beta2;     /* 0.999 */
  double adam_eps;       /* 1e-8  */
  double gamma;          /* 0.99  */
  double polyak_k;       /* tau * n_env (agent.py:105-110); used when soft_update        */
  double per_eps;        /* 1e-4  (replay_memory.py:49) */
  double per_alpha;      /* 0.6   (replay_memory.py:50) */
  double per_pmax;       /* 1.0   (replay_memory.py:54) */
} rmc_hyper_t;

typedef struct {
  int64_t capacity;
  int64_t size;          /* sum_tree.py: self.size                                   */
  int64_t data_pointer;  /* sum_tree.py: self.data_pointer                           */
  double total_priority; /* sum_tree.py:63-65  tree[0]                               */
  double max_priority;   /* sum_tree.py:67-69  == max(leaves[:size])                 */
  double min_priority;   /* sum_tree.py:71-73  == min(leaves[:size])                 */
  int64_t rejected_nodes;/* entries of caller-supplied tree-index lists outside the leaf range [cap-1, 2cap-2] that
                            rmc_per_update* skipped so far (the reference would raise IndexError there) */
} rmc_replay_stats_t;

/* What one learner step does (bit mask).  The default Agent.learn() uses
 * SAMPLE|FORWARD|PRIORITY|BACKWARD|ADAM; Agent.update_target_network() is RMC_PH_POLYAK
 * (or fused into the same launch by passing it here). */
enum rmc_phase {
  RMC_PH_SAMPLE = 1,    /* draw the minibatch (PER descent or uniform) and gather rows        */
  RMC_PH_FORWARD = 2,   /* target/online forwards, TD target, |td|, Huber, dgrad              */
  RMC_PH_PRIORITY = 4,  /* |td| -> priority, sum-tree write-back (PER only)                   */
  RMC_PH_BACKWARD = 8,  /* weight/bias gradients (deterministic order) -> RMC_GRADS           */
  RMC_PH_ADAM = 16,     /* Adam update of the online net from RMC_GRADS                       */
  RMC_PH_POLYAK = 32,   /* soft target update with the post-Adam weights                      */
  RMC_PH_HARDSYNC = 64  /* target <- online (agent.py:102-103)                                */
};
enum rmc_precision { RMC_PREC_FP32 = 0, RMC_PREC_BF16_TC = 1 };
#define RMC_PH_LEARN (RMC_PH_SAMPLE | RMC_PH_FORWARD | RMC_PH_PRIORITY | RMC_PH_BACKWARD | RMC_PH_ADAM)

/* The repo-HEAD network (env/dqn_config.py:66-193 TwoStreamHybridNetwork + network_config): the state is
 * [macro vector | grid flattened C,H,W]; the grid goes through n_conv Conv2d(3x3, padding 1, stride (sh, sw)) + act
 * layers, is flattened and concatenated with the macro vector (flatten first, macro last), then n_dense
 * Linear + act layers feed the heads (fc_val/fc_adv or fc_out, dqn/network.py:50-117).  Parameters travel in
 * torch state_dict() order: net.cnn_stream.{0,2,..}.{weight,bias}, net.dense_stream.{0,2,..}.{weight,bias}, heads. */
typedef struct {
  int32_t macro_len;                 /* 14                                   */
  int32_t grid_c, grid_h, grid_w;    /* 2, 27, 5                             */
  int32_t n_conv;                    /* 3 (<= 4)                             */
  int32_t conv_out[4], conv_sh[4], conv_sw[4];   /* 32,64,64 ; (1,1),(2,1),(2,2) */
  int32_t n_dense;                   /* 2 (<= 3)                             */
  int32_t dense_out[3];              /* 512, 256                             */
  int32_t n_actions, dueling, double_dqn, prioritized, activation;
} rmc_hybrid_spec_t;

/* Per-step inputs of a learner step. */
typedef struct {
  int64_t batch;            /* B                                                            */
  int32_t phases;           /* rmc_phase mask                                               */
  int32_t precision;        /* rmc_precision: 0 = exact fp32 FFMA path (parity path, default);
                               1 = tcgen05/TMEM tensor-core path, bf16 operands with fp32 accumulation
                               (dense large-batch / ensemble configs; stated looser bound: gradients
                               within 1e-1 max-norm-relative of the fp32 path, measured <= 5e-2).  Needs FORWARD and
                               BACKWARD in `phases` and obs_dim <= 16.                              */
  double per_beta;          /* np.interp(step,[0,eps_dec],[0.4,1.0]) (replay_memory.py:74)  */
  const double* u_dev;      /* PER: B injected uniforms in [0,1) (float64) or NULL -> Philox */
  const int64_t* idx_dev;   /* uniform replay: B injected deque positions (0 = oldest) or NULL
                               -> on-device keyed permutation (sampling WITHOUT replacement,
                               like random.sample, replay_memory.py:38-39)                   */
  uint64_t seed;            /* device RNG key   (used when u_dev / idx_dev is NULL)          */
  uint64_t counter;         /* device RNG counter, normally the learner step                 */
  int64_t adam_t;           /* optimizer step count t >= 1 (bias corrections, float64 host)  */
  const float* grads_in_dev;/* optional: with RMC_PH_ADAM and without RMC_PH_BACKWARD, apply Adam
                               to these gradients (device layout = RMC_GRADS order of
                               rmc_learner_get_params); used by the sharded large-batch path
                               after the NCCL all-reduce.  NULL -> use the handle's RMC_GRADS */
  int64_t shard_offset;     /* sharded step: first global sample index of this rank          */
  int64_t global_batch;     /* sharded step: B of the whole job (0 -> = batch)               */
} rmc_step_args_t;

/* ---------------------------------------------------------------- library / errors --- */
int32_t rmc_abi_version(void);
const char* rmc_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t rmc_launch_count(void);

/* ---------------------------------------------------------------- replay ring + tree -- */
/* SumTree.__init__ (dqn/utils/sum_tree.py:6-13), ReplayMemoryNaive/Prioritized.__init__
 * (dqn/replay_memory.py:24-28,43-54).  One row per transition, AoS, 16-byte aligned:
 * [obs D | next_obs D | action(i32 bits) | reward | done | pad].  prioritized=1 adds the
 * float64 heap-layout tree with the reference's exact indexing (2*cap-1 nodes). */
int32_t rmc_replay_create(rmc_replay_t** out, int64_t capacity, int32_t obs_dim, int32_t prioritized,
                          int32_t device);
int32_t rmc_replay_destroy(rmc_replay_t* r);

/* ReplayMemory*.store_transitions (replay_memory.py:30-36,56-67) + SumTree.add (:34-40):
 * append n transitions at data_pointer (ring), new leaves get max_priority (1.0 if it is 0),
 * read once per call.  actions are int64, done is 0/1 float. */
int32_t rmc_replay_push(rmc_replay_t* r, const float* obs_dev, const int64_t* act_dev, const float* rew_dev,
                        const float* done_dev, const float* next_obs_dev, int64_t n, rmc_stream_t s);
int32_t rmc_replay_push_host(rmc_replay_t* r, const float* obs_host, const int64_t* act_host,
                             const float* rew_host, const float* done_host, const float* next_obs_host,
                             int64_t n, rmc_stream_t s);

/* Overwrite the priorities of the first n leaves (data order) and rebuild the tree bottom
 * up -- used to seed non-degenerate trees for tests / benchmarks. */
int32_t rmc_replay_set_priorities(rmc_replay_t* r, const float* pri_dev, int64_t n, rmc_stream_t s);

/* total/max/min_priority, size, data_pointer (sum_tree.py:63-73).  Synchronises the stream. */
int32_t rmc_replay_stats_sync(rmc_replay_t* r, rmc_replay_stats_t* out, rmc_stream_t s);
/* copy tree nodes [first, first+n) (float64) / ring rows to host -- parity tests, SumTree.tree */
int32_t rmc_replay_read_tree_sync(rmc_replay_t* r, double* out_host, int64_t first, int64_t n, rmc_stream_t s);
int32_t rmc_replay_read_rows_sync(rmc_replay_t* r, float* out_host, int64_t first_slot, int64_t n, rmc_stream_t s);
int32_t rmc_replay_row_floats(const rmc_replay_t* r);
/* Exact-resume side-car (not a reference interface; the reference refills 100 k env steps on resume, train.py:63-81):
 * restore a replay saved with the two read calls above -- rows_host [size][row_floats] in slot order, leaf_pri_host [size]
 * = the float32-exact leaf priorities (NULL for uniform replay), and the ring cursor.  The inner nodes and the extremes are
 * rebuilt from the leaves (exact sums), so tree, total/max/min and every later sample equal the saved run's.  Synchronises. */
int32_t rmc_replay_load_host(rmc_replay_t* r, const float* rows_host, const float* leaf_pri_host, int64_t size,
                             int64_t data_pointer, rmc_stream_t s);

/* ReplayMemoryPrioritized.sample_transitions (replay_memory.py:69-92) + SumTree.get_leaf
 * (sum_tree.py:42-61): stratified proportional sampling.  out_nodes = tree indices (bit-exact
 * for injected u), out_is_w = IS weights (float32), out_rows = gathered rows [B][row_floats]. */
int32_t rmc_per_sample(rmc_replay_t* r, int64_t batch, double beta, const double* u_dev, uint64_t seed,
                       uint64_t counter, int64_t* out_nodes_dev, float* out_is_w_dev, float* out_rows_dev,
                       rmc_stream_t s);
/* SumTree.get_leaf (sum_tree.py:42-61) for n explicit prefix values v (float64): tree index and
 * leaf priority per value. */
int32_t rmc_tree_get_leaf(rmc_replay_t* r, const double* v_dev, int64_t n, int64_t* out_nodes_dev,
                          double* out_pri_dev, rmc_stream_t s);
/* ReplayMemoryNaive.sample_transitions (replay_memory.py:38-39). */
int32_t rmc_uniform_sample(rmc_replay_t* r, int64_t batch, const int64_t* idx_dev, uint64_t seed,
                           uint64_t counter, int64_t* out_slots_dev, float* out_rows_dev, rmc_stream_t s);

/* SumTree.update applied in batch order, duplicates: last wins (replay_memory.py:97-98,
 * sum_tree.py:15-32).  Bit-exact given identical float32 priorities. */
int32_t rmc_per_update(rmc_replay_t* r, const int64_t* nodes_dev, const float* pri_dev, int64_t batch,
                       rmc_stream_t s);
/* ReplayMemoryPrioritized.update_batch_priorities (replay_memory.py:94-98):
 * p = min(|td| + eps, pmax)^alpha in float32 (<= 1 ulp vs numpy), then rmc_per_update.
 * If out_pri_dev != NULL the float32 priorities are also written there. */
int32_t rmc_per_update_from_td(rmc_replay_t* r, const int64_t* nodes_dev, const float* abs_td_dev,
                               int64_t batch, float eps, float alpha, float pmax, float* out_pri_dev,
                               rmc_stream_t s);

/* ---------------------------------------------------------------- learner ------------- */
/* Network.__init__ / Deep(Dueling)QNetwork.__init__ (dqn/network.py:12-19,50-65,77-88) +
 * the agent wiring (dqn/agent.py:275-320).  Weights start at zero: load them with
 * rmc_learner_set_params (the Python mirror initialises with torch's nn.Linear init). */
int32_t rmc_learner_create(rmc_learner_t** out, const rmc_net_spec_t* spec, const rmc_hyper_t* hyper,
                           int64_t max_batch, int32_t device);
/* Same handle type for the hybrid CNN + MLP network: every rmc_learner_* entry point below (set/get_params,
 * step, output, loss, q_values, heads, act) dispatches on the network kind.  Exact fp32 path only; no
 * tensor-core mode, ensemble launch or sharded step for this network yet. */
int32_t rmc_learner_create_hybrid(rmc_learner_t** out, const rmc_hybrid_spec_t* spec, const rmc_hyper_t* hyper,
                                  int64_t max_batch, int32_t device);
int32_t rmc_learner_destroy(rmc_learner_t* l);
int64_t rmc_learner_param_count(const rmc_learner_t* l);

/* Parameters in torch state_dict() order, flattened (checkpoint order of network.py:27-31):
 * net.0.weight[H1,D] net.0.bias[H1] net.2.weight[H2,H1] net.2.bias[H2]
 * then fc_val.weight[1,H2] fc_val.bias[1] fc_adv.weight[A,H2] fc_adv.bias[A]  (dueling)
 * or   fc_out.weight[A,H2] fc_out.bias[A]. */
int32_t rmc_learner_set_params(rmc_learner_t* l, int32_t kind, const float* src, int64_t n, int32_t src_is_host,
                               rmc_stream_t s);
int32_t rmc_learner_get_params(rmc_learner_t* l, int32_t kind, float* dst, int64_t n, int32_t dst_is_host,
                               rmc_stream_t s);
int32_t rmc_learner_set_hyper(rmc_learner_t* l, const rmc_hyper_t* hyper);

/* {Simple,Double,PerDouble}Agent.learn (dqn/agent.py:166-185,204-226,245-272) and
 * Agent.update_target_network (:101-110) as ONE launch; `phases` selects which
 * parts run (split variants for parity tests).  Outputs stay on the device; see
 * rmc_learner_read_*. */
int32_t rmc_learner_step(rmc_learner_t* l, rmc_replay_t* r, const rmc_step_args_t* a, rmc_stream_t s);
/* Agent.store_transitions (dqn/agent.py:80-84) of the n (<= 8) rows of this env step followed by the learner step,
 * issued from one host call (train.py:91-101 calls them back to back): == rmc_replay_push_host + rmc_learner_step. */
int32_t rmc_learner_step_push(rmc_learner_t* l, rmc_replay_t* r, const rmc_step_args_t* a, const float* obs_host,
                              const int64_t* act_host, const float* rew_host, const float* done_host,
                              const float* next_obs_host, int64_t n, rmc_stream_t s);

/* Learner-step products of the last step (device pointers, valid until the next step):
 * name in {"nodes"(i64), "is_w","q_sa","y","abs_td","huber","pri","loss"(1), "q_next_tgt"(B*A),
 *          "q_next_on"(B*A), "q"(B*A), "rows"(B*row_floats)} */
int32_t rmc_learner_output(rmc_learner_t* l, const char* name, void** dev_ptr, int64_t* n_elems);
/* loss of the last step -> host (synchronises).  RMC_ERR_STATE if a launch of this learner tripped the in-kernel
 * watchdog (see rmc_learner_status). */
int32_t rmc_learner_loss_sync(rmc_learner_t* l, float* out_host, rmc_stream_t s);
/* Health of the fused step (does not synchronise): the kernel's agent barrier and hand-off words need every CTA of the
 * launch co-resident.  The default launch (programmatic dependent launch, RMC_LAUNCH unset) establishes that by
 * construction -- grid <= resident capacity, fused-step launches of one device serialised across streams -- and
 * RMC_LAUNCH=coop asks the driver to guarantee it.  Should a wait inside the kernel nevertheless exceed 2 s (a second
 * process running this kernel on the same GPU through MPS), the kernel gives up instead of hanging the GPU and this call
 * (like rmc_learner_loss_sync) returns RMC_ERR_STATE; *failed_epoch = launch epoch of the failed step, 0 = healthy. */
int32_t rmc_learner_status(rmc_learner_t* l, uint32_t* failed_epoch);

/* Network.forward (network.py:59-63,90-96): Q values of `which` net (RMC_ONLINE/RMC_TARGET). */
int32_t rmc_learner_q_values(rmc_learner_t* l, int32_t which, const float* obs_dev, int64_t n, float* q_out_dev,
                             rmc_stream_t s);
/* DuelingDeepQNetwork.value / .advantages (network.py:98-108): raw head outputs [n][NH]; dueling: column 0 is
 * the value stream, columns 1..A the advantages; plain head: NH = A (same as Q). */
int32_t rmc_learner_heads(rmc_learner_t* l, int32_t which, const float* obs_dev, int64_t n, float* heads_out_dev,
                          rmc_stream_t s);
/* Network.actions (network.py:67-74, 110-117): greedy actions, dueling -> argmax of RAW
 * advantages, plain -> argmax Q; first maximum wins. */
int32_t rmc_learner_act(rmc_learner_t* l, const float* obs_dev, int64_t n, int64_t* actions_dev, rmc_stream_t s);
/* Tensor-core mode of the two calls above for the dense batched-act config (tcgen05.mma, bf16 operands, fp32
 * accumulation in tensor memory).  NOT the parity path: Q within 1e-2 max-norm-relative of the fp32 kernel,
 * greedy actions equal except near-ties (stated looser bound of the north star).  obs_dim <= 16. */
int32_t rmc_learner_act_tc(rmc_learner_t* l, const float* obs_dev, int64_t n, int64_t* actions_dev, rmc_stream_t s);
int32_t rmc_learner_heads_tc(rmc_learner_t* l, const float* obs_dev, int64_t n, float* heads_out_dev, rmc_stream_t s);
/* same with host buffers (synchronises): what Agent.choose_actions calls once per environment step (dqn/agent.py:92-99).
 * n <= 32 states of the macro MLP (the n_env of train.py): one launch, no copies -- the states travel in the kernel-argument
 * buffer and the actions come back through mapped pinned host memory followed by the launch's epoch word.  Larger n:
 * H2D + kernel + D2H + stream synchronisation. */
int32_t rmc_learner_act_host_sync(rmc_learner_t* l, const float* obs_host, int64_t n, int64_t* actions_host,
                                  rmc_stream_t s);
/* Agent.choose_actions (dqn/agent.py:92-99) in one call for vectorised envs: greedy act, then row i explores with
 * probability epsilon (uniform action) from Philox(seed, counter, i) on the device.  Same distribution as the reference,
 * not the same random stream (the Python mirror's default keeps the reference's host RNG stream). */
int32_t rmc_learner_act_eps_host_sync(rmc_learner_t* l, const float* obs_host, int64_t n, int64_t* actions_host,
                                      float epsilon, uint64_t seed, uint64_t counter, rmc_stream_t s);

/* Diagnostics (not a reference interface): per-CTA phase timestamps (%globaltimer, ns) of the last
 * rmc_learner_step: out_host[cta*32 + k], k = 0 start, 1 sampled, 2 target weights landed,
 * 3 target pass done, 4 online weights landed, 5 row phase done, 6 past the barrier, 7 done, 8..19 finer stamps
 * (profiles/tools/phase_timeline.py names them). */
int32_t rmc_learner_debug_timing(rmc_learner_t* l, int32_t enable);
/* Diagnostics: kernel-span recorder for the multi-kernel pipelines (tensor-core step, grid-wide sampler / tree write-back,
 * peer exchange).  While enabled every such kernel notes its first CTA's start and last CTA's end (%globaltimer, ns); the
 * read call returns {start, end} for 64 kernel slots (unused: {~0, 0}), resets them, and names the slots (comma separated).
 * Shows the real overlap inside a graph-launched, two-stream step, which neither ncu nor CUDA events can. */
int32_t rmc_debug_spans(int32_t device, int32_t enable);
int32_t rmc_debug_spans_read_sync(int32_t device, uint64_t* out128_host, char* names_out, int32_t names_cap);
/* [min CTA start, max CTA end] (%globaltimer ns) of the last 64 launches, slot = epoch % 64 (launch-gap diagnostic) */
int32_t rmc_learner_debug_gaps_sync(rmc_learner_t* l, uint64_t* out128_host, rmc_stream_t s);
int32_t rmc_learner_debug_read_sync(rmc_learner_t* l, uint64_t* out_host, int32_t max_ctas, int32_t* n_ctas,
                                    rmc_stream_t s);

/* ---------------------------------------------------------------- ensembles (C4) ------- */
/* N independent agents (own replay, weights, Adam state, RNG stream) stepped by one launch
 * (grid.y = agent); no communication.  All members must share spec/hyper/batch. */
int32_t rmc_group_create(rmc_group_t** out, rmc_learner_t* const* learners, rmc_replay_t* const* replays,
                         int32_t n_agents);
int32_t rmc_group_destroy(rmc_group_t* g);
/* u_dev / idx_dev (if given) hold n_agents * batch entries, agent-major.  precision = RMC_PREC_BF16_TC: the members'
 * tensor-core steps (rmc_learner_step in that mode) run side by side on internal streams, forked from / joined into `s`;
 * meant for ensembles of large-batch members -- at B = 256 per member the fused fp32 launch is faster. */
int32_t rmc_group_step(rmc_group_t* g, const rmc_step_args_t* a, rmc_stream_t s);
/* Agent.store_transitions of EVERY member for one env step (dqn/agent.py:70-73 -> dqn/replay_memory.py:49-57): member i
 * stores the n host rows obs_host[i], act_host[i], ... (n <= 8 each; new rows enter with the member's max priority).  One
 * launch for the whole ensemble when the packed rows fit the kernel-argument buffer, else one small push per member. */
int32_t rmc_group_push_host(rmc_group_t* g, const float* const* obs_host, const int64_t* const* act_host,
                            const float* const* rew_host, const float* const* done_host,
                            const float* const* next_obs_host, int64_t n, rmc_stream_t s);

/* ---------------------------------------------------------------- sharded large batch (C5) --- */
/* The reference has no multi-device path; this is the data-parallel form of
 * PerDoubleAgent.learn (dqn/agent.py:245-272) for one logical agent replicated on `world` GPUs of one
 * node: rank r draws the strata [lo, hi) = shard_range(global_batch, r, world) of the GLOBAL stratified
 * sample (replay_memory.py:72-80 with seg = total / global_batch), computes its gradient slice scaled by
 * 1 / global_batch, and the exchange + Adam run as kernels over NVLink peer memory (no library collective):
 * each rank publishes its gradient blob into its own exchange buffer, flags every peer, and one kernel
 * sums the ranks' blobs in rank order straight from peer memory and applies Adam (+ Polyak) -- identical
 * bits on every replica.  PER: (leaf, |td|) slices travel the same way and every replica applies the full
 * write-back in global batch order.
 *   rmc_comm_create  : allocate this rank's exchange buffer (2 parity slots + flags)
 *   rmc_comm_export  : its cudaIpcMemHandle_t (64 bytes) to all-gather across the ranks' processes
 *                      (and/or the raw device pointer for ranks living in one process)
 *   rmc_comm_connect : map the peers' buffers (handles64 = world x 64 bytes, rank-major; or same-process
 *                      device pointers)
 *   rmc_learner_step_sharded : the whole step; `a` carries batch = hi - lo, shard_offset = lo,
 *                      global_batch, phases (FORWARD|BACKWARD|ADAM required), precision.  stages = 3: the
 *                      whole step (one GPU per rank); 1 = local gradients + publish, 2 = reduce + Adam +
 *                      write-back -- only for ranks EMULATED on one GPU, where a rank's waiting reduce kernel
 *                      would keep the other rank's cooperative step kernel from becoming resident
 *   rmc_comm_status_sync : 0, or the epoch of an exchange that timed out (a peer never published within
 *                      RMC_COMM_TIMEOUT_MS, default 10 s).  The decision is taken once per kernel for all its blocks:
 *                      a timed-out exchange applies NOTHING (no Adam, no Polyak, no write-back) on this rank. */
int32_t rmc_comm_create(rmc_comm_t** out, rmc_learner_t* l, int32_t rank, int32_t world, int64_t global_batch_max);
int32_t rmc_comm_export(rmc_comm_t* c, void* handle64_out, void** local_ptr_out);
int32_t rmc_comm_connect(rmc_comm_t* c, const void* handles64, void* const* same_process_ptrs);
int32_t rmc_comm_destroy(rmc_comm_t* c);
int32_t rmc_comm_status_sync(rmc_comm_t* c, uint32_t* timed_out_epoch, rmc_stream_t s);
int32_t rmc_learner_step_sharded(rmc_learner_t* l, rmc_replay_t* r, rmc_comm_t* c, const rmc_step_args_t* a,
                                 int32_t stages, rmc_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* RMC_B200_H */
