"""-m gpu: ensemble launch (C4) and the sharded large-batch step (C5) on one GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import parity_utils as PU
from tests import recipes as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo,B", [("PerDuelingDoubleDQNAgent", 32), ("DuelingDoubleDQNAgent", 32), ("PerDuelingDoubleDQNAgent", 256)])
def test_ensemble_launch_equals_individual_steps(algo, B):
    """N agents stepped by ONE launch (grid.y = agent) == the same agents stepped one by one."""
    from multimodal_drl_rmc_b200.parallel import AgentEnsemble
    n = 3
    solo = [PU.make_pair(algo, 14, B, 2000, 2000, seed=20 + k)[1] for k in range(n)]
    team = [PU.make_pair(algo, 14, B, 2000, 2000, seed=20 + k)[1] for k in range(n)]
    ens = AgentEnsemble(team)
    rng = np.random.default_rng(0)
    per = algo.startswith("Per")
    for step in range(3):
        inj = rng.random((n, B)) if per else np.stack([rng.permutation(2000)[:B] for _ in range(n)])
        for k, a in enumerate(solo):
            a.step = step
            a.learn(u=inj[k], fuse_target_update=True) if per else a.learn(indices=inj[k], fuse_target_update=True)
            a.update_target_network()
        for a in team:
            a.step = step
        ens.learn(u=inj) if per else ens.learn(indices=inj)
    # The single-agent launch (148 CTAs) forms the small tensors' gradients from per-tile partial sums and the ensemble launch
    # (18 CTAs per agent) sums them row by row: two fixed summation orders of the same fp32 terms, so the weights agree to
    # rounding, not bit for bit (each form is deterministic; the oracle comparison is test_ensemble_launch_matches_oracle_per_agent).
    for a, b in zip(solo, team):
        wa, wb = PU.flat_sd(a.online_network), PU.flat_sd(b.online_network)
        assert np.max(np.abs(wa - wb)) <= 3e-4 and np.mean(np.abs(wa - wb)) <= 1e-7        # ill-conditioned Adam elements move by <= lr per step
        ta, tb = PU.flat_sd(a.target_network), PU.flat_sd(b.target_network)
        assert R.max_rel(ta, tb) < 1e-5
        if per:
            sa, sb = a.replay_memory_buffer._ring.stats(), b.replay_memory_buffer._ring.stats()
            assert abs(sa.total_priority - sb.total_priority) <= 1e-5 * abs(sb.total_priority)
            assert sa.max_priority == sb.max_priority and sa.size == sb.size


def test_ensemble_push_of_all_members_in_one_launch_equals_member_pushes():
    """store_transitions of every member, delivered by AgentEnsemble in ONE launch (rmc_group_push_host, block = member):
    rings, cursors, trees and extremes bit-identical to the same rows pushed member by member (dqn/agent.py:70-73,
    dqn/replay_memory.py:49-57: new rows enter with the max priority), also across the ring's wrap-around."""
    from multimodal_drl_rmc_b200.parallel import AgentEnsemble
    n, B, cap = 3, 32, 600
    solo = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap - 2, seed=80 + k)[1] for k in range(n)]
    team = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap - 2, seed=80 + k)[1] for k in range(n)]
    ens = AgentEnsemble(team)
    rng = np.random.default_rng(9)
    for it in range(5):                      # 5 env steps of 1 row each: the cursor wraps after the second
        for k in range(n):
            o, o2 = rng.random((1, 14), dtype=np.float32), rng.random((1, 14), dtype=np.float32)
            a, r, d = [int(rng.integers(0, 8))], [float(rng.random())], [bool(rng.integers(0, 2))]
            for ag in (solo[k], team[k]):
                ag.store_transitions(o, a, r, d, o2, None)
        for ag in solo:
            ag.replay_memory_buffer._ring.flush()
        ens._deliver_pending()
        for a, b in zip(solo, team):
            ra, rb = a.replay_memory_buffer._ring, b.replay_memory_buffer._ring
            sa, sb = ra.stats(), rb.stats()
            assert (sa.size, sa.data_pointer, sa.total_priority, sa.max_priority, sa.min_priority) == \
                   (sb.size, sb.data_pointer, sb.total_priority, sb.max_priority, sb.min_priority)
            np.testing.assert_array_equal(a.replay_memory_buffer.replay_buffer.tree, b.replay_memory_buffer.replay_buffer.tree)
            np.testing.assert_array_equal(ra.read_rows(0, cap), rb.read_rows(0, cap))
    # and inside the normal flow: store, then one ensemble step (the rows are delivered by learn())
    for k in range(n):
        team[k].store_transitions(rng.random((1, 14), dtype=np.float32), [1], [0.5], [False], rng.random((1, 14), dtype=np.float32), None)
        team[k].step = 3
    ens.learn()
    assert all(t.replay_memory_buffer._ring._pending == 0 for t in team)
    assert all(np.isfinite(t.last_loss()) for t in team)


@pytest.mark.parametrize("B,cap", [(64, 1500), (250, 3001)])
def test_ensemble_launch_matches_oracle_per_agent(B, cap):
    """The one-launch ensemble step against the ORACLE, member by member (not only against the single-agent CUDA path):
    indices and trees bit-exact, Q / loss / gradients / weights at 1e-5 (dqn/agent.py:245-272 per agent).  B = 250 with three
    members: 49 CTAs per agent own 63 four-row tiles (ragged last one) -- several tiles per CTA, the in-kernel sampler with two
    descents per warp (per_descend_cached2) on a tree whose leaves lie on two depths."""
    from multimodal_drl_rmc_b200 import _lib
    from multimodal_drl_rmc_b200.parallel import AgentEnsemble
    n = 3
    pairs = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap, seed=60 + k) for k in range(n)]
    ens = AgentEnsemble([p[1] for p in pairs])
    rng = np.random.default_rng(4)
    for step in range(2):
        u = rng.random((n, B))
        traces = []
        for k, (orc, ag) in enumerate(pairs):
            orc.step = ag.step = 30 + step
            tr = {}
            orc.learn(u=u[k], trace=tr)
            orc.sync_target()
            traces.append(tr)
        ens.learn(u=u)
        for k, (orc, ag) in enumerate(pairs):
            tr = traces[k]
            np.testing.assert_array_equal(PU.gpu_out(ag, "nodes", torch.int64), tr["nodes"])
            assert R.max_rel(PU.gpu_out(ag, "q_sa"), tr["q_sa"].reshape(-1)) < 1e-5
            assert R.max_rel(PU.gpu_out(ag, "y"), tr["y"].reshape(-1)) < 1e-5
            assert abs(ag.last_loss() - tr["loss"]) <= 1e-5 * abs(tr["loss"])
            g_gpu = ag._lh.get_params(_lib.GRADS).cpu().numpy()
            g_ref = np.concatenate([tr["grads"][kk].ravel() for kk, _ in orc.online.named_parameters()])
            sizes = PU.tensor_sizes(orc.online)
            pt = PU.per_tensor_max_rel(g_gpu, g_ref, sizes)
            pt.pop("fc_val.bias", None)     # one-element cancelling sum: measured against sum |g_i| in run_parity_case
            assert max(pt.values()) < 1e-5, pt
            well = np.abs(g_ref) >= 1e-6
            w_gpu, w_ref = PU.flat_sd(ag.online_network), PU.flat_sd(orc.online)
            assert R.max_rel(np.where(well, w_gpu, w_ref), w_ref) < 1e-5 or step > 0
            assert np.max(np.abs(w_gpu - w_ref)) <= 1e-4 * (step + 1)
            assert R.max_rel(PU.flat_sd(ag.target_network), PU.flat_sd(orc.target)) < 1e-4
            # trees: equal once the 1-ulp |td| -> p differences are removed (same procedure as run_parity_case)
            p_ref = np.power(np.minimum(tr["abs_td"].reshape(-1) + np.float32(1e-4), np.float32(1.0)), np.float32(0.6)).astype(np.float32)
            n_t, p_t = torch.as_tensor(tr["nodes"], device=ag.device), torch.as_tensor(p_ref, device=ag.device)
            _lib.check(_lib.lib().rmc_per_update(ag.replay_memory_buffer._ring.handle, n_t.data_ptr(), p_t.data_ptr(), B, _lib.stream_ptr(ag.device.index)))
            np.testing.assert_array_equal(ag.replay_memory_buffer.replay_buffer.tree, orc.replay.tree.tree)


def test_ensemble_tensor_core_mode_equals_member_steps():
    """Tensor-core mode of the ensemble launch: every member's tcgen05 step, side by side on internal streams; with the same
    injected uniforms each member ends bit-identical to the same agent stepped alone in tensor-core mode."""
    from multimodal_drl_rmc_b200.parallel import AgentEnsemble
    n, B, cap = 3, 1024, 4000
    solo = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap, seed=70 + k)[1] for k in range(n)]
    team = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap, seed=70 + k)[1] for k in range(n)]
    for a in solo + team:
        a.learn_precision = "bf16"
    ens = AgentEnsemble(team)
    rng = np.random.default_rng(9)
    for step in range(3):
        u = rng.random((n, B))
        for k, a in enumerate(solo):
            a.step = step
            a.learn(u=u[k])
            a.update_target_network()
        for a in team:
            a.step = step
        ens.learn(u=u)
    torch.cuda.synchronize()
    for a, b in zip(solo, team):
        np.testing.assert_array_equal(PU.flat_sd(a.online_network), PU.flat_sd(b.online_network))
        np.testing.assert_array_equal(PU.flat_sd(a.target_network), PU.flat_sd(b.target_network))
        np.testing.assert_array_equal(a.replay_memory_buffer.replay_buffer.tree, b.replay_memory_buffer.replay_buffer.tree)
        assert a.last_loss() == b.last_loss()


def test_ensemble_rejects_members_with_different_hyper_parameters():
    from multimodal_drl_rmc_b200 import macro_config
    from multimodal_drl_rmc_b200.parallel import AgentEnsemble
    import tempfile
    tmp = tempfile.mkdtemp()
    a = PU.make_pair("DuelingDoubleDQNAgent", 14, 32, 200, 200, seed=1)[1]
    b = macro_config.make_agent("DuelingDoubleDQNAgent", 14, 32, 200, save_dir=tmp + "/", log_dir=tmp + "/", lr=3e-4)
    obs, act, rew, done, nxt = PU.synthetic_transitions(200, 14, 5)
    b.store_transitions(obs, act.tolist(), rew.tolist(), done.astype(bool).tolist(), nxt, None)
    with pytest.raises(ValueError):
        AgentEnsemble([a, b])


def test_sharded_step_emulated_on_one_gpu_equals_full_batch():
    """Two 'ranks' emulated one after the other on one GPU (replicas with identical state): the union of their
    shards is the full batch, the summed gradient blobs equal the full-batch gradients, and Adam from the summed
    gradients gives the full-batch weights."""
    from multimodal_drl_rmc_b200 import _lib
    from multimodal_drl_rmc_b200.parallel import shard_range
    B, W = 512, 2
    full = PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, 3000, 3000, seed=31)[1]
    reps = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, 3000, 3000, seed=31)[1] for _ in range(W)]
    u = np.random.default_rng(2).random(B)
    full.step = 9
    full.learn(u=u)
    g_full = full._lh.get_params(_lib.GRADS).cpu().numpy()
    nodes_full = PU.gpu_out(full, "nodes", torch.int64)
    loss_full = full.last_loss()
    lib = _lib.lib()
    blobs, nodes, losses = [], [], 0.0
    for r, ag in enumerate(reps):
        lo, hi = shard_range(B, r, W)
        a = _lib.StepArgs()
        a.batch, a.global_batch, a.shard_offset = hi - lo, B, lo
        a.phases = _lib.PH_SAMPLE | _lib.PH_FORWARD | _lib.PH_BACKWARD
        a.per_beta = ag._beta(9)
        ut = torch.as_tensor(u[lo:hi].copy(), device=ag.device)
        a.u_dev = ut.data_ptr()
        _lib.check(lib.rmc_learner_step(ag._lh.handle, ag.replay_memory_buffer._ring.handle, C.byref(a), _lib.stream_ptr()))
        blobs.append(ag._lh.output("grads_blob").clone())
        nodes.append(PU.gpu_out(ag, "nodes", torch.int64)[: hi - lo])
        losses += ag.last_loss()
    np.testing.assert_array_equal(np.concatenate(nodes), nodes_full)
    assert abs(losses - loss_full) / abs(loss_full) < 1e-5
    summed = blobs[0] + blobs[1]
    ag = reps[0]
    b = _lib.StepArgs()
    b.batch, b.adam_t, b.phases = B // W, 1, _lib.PH_ADAM
    b.grads_in_dev = summed.data_ptr()
    _lib.check(lib.rmc_learner_step(ag._lh.handle, ag.replay_memory_buffer._ring.handle, C.byref(b), _lib.stream_ptr()))
    ag._lh.version[_lib.ONLINE] += 1
    # gradients: compare in torch order through a scratch learner blob
    ag._lh_tmp = summed
    w_shard, w_full = PU.flat_sd(ag.online_network), PU.flat_sd(full.online_network)
    well = np.abs(g_full) >= 1e-6
    assert R.max_rel(np.where(well, w_shard, w_full), w_full) < 1e-5
    assert np.max(np.abs(w_shard - w_full)) <= 1e-4


def test_large_batch_per_step_uses_grid_wide_tree_path():
    """B = 8192 > 4096: several row tiles per CTA and the multi-kernel tree write-back; vs the oracle."""
    res = PU.run_parity_case("PerDuelingDoubleDQNAgent", 14, 8192, 20000, 20000, 1, seed=13)
    print(res)
    assert res["nodes_equal"] and res["tree_equal"]
    assert res["max_rel_q"] < 1e-5 and res["max_rel_loss"] < 1e-5 and res["max_rel_grads"] < 1e-5
    assert res["max_rel_weights"] < 1e-5 and res["max_pri_ulp"] <= 1.0


@pytest.mark.parametrize("B,precision,steps", [(512, "fp32", 3), (8192, "fp32", 2), (8192, "bf16", 2), (515, "fp32", 2)])
def test_peer_memory_exchange_two_ranks_on_one_gpu(B, precision, steps):
    """ShardedLearner(exchange="peer"): the gradient exchange + Adam as kernels over peer memory (here: two ranks emulated
    in one process on two streams of one GPU, buffers wired by plain device pointers instead of CUDA IPC handles).
    Replicas must end bit-identical (weights, target, tree); against the single-GPU full-batch step the weights agree to
    1e-5 (the rank-order gradient sum differs from the full-batch summation order) and sampled indices are bit-exact."""
    from multimodal_drl_rmc_b200 import _lib
    from multimodal_drl_rmc_b200.parallel import ShardedLearner
    W, cap = 2, 20000
    orc, full = PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap, seed=41)
    reps = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap, seed=41)[1] for _ in range(W)]
    for ag in reps + [full]:
        ag.learn_precision = precision
    members = [ShardedLearner(ag, exchange="peer", rank=r, world=W) for r, ag in enumerate(reps)]
    ShardedLearner.connect_same_process(members)
    streams = [torch.cuda.Stream() for _ in range(W)]
    rng = np.random.default_rng(3)
    torch.cuda.synchronize()
    for s in range(steps):
        u = rng.random(B)
        for ag in reps + [full]:
            ag.step = 50 + s
        full.learn(u=u, fuse_target_update=True)
        full.update_target_network()
        nodes_full = PU.gpu_out(full, "nodes", torch.int64)
        g_full = full._lh.get_params(_lib.GRADS).cpu().numpy()
        # one GPU hosts both ranks here: local gradients + publish of every rank first, then the waiting reduce kernels
        for stage in (1, 2):
            for m, st in zip(members, streams):
                with torch.cuda.stream(st):
                    m.learn(u=u, fuse_target_update=True, stages=stage)
        torch.cuda.synchronize()
        assert all(m.exchange_status() == 0 for m in members)
        got = np.concatenate([PU.gpu_out(m.agent, "nodes", torch.int64)[: m.hi - m.lo] for m in members])
        if s == 0:      # later steps start from weights/trees that differ from the full-batch run in the last bits
            np.testing.assert_array_equal(got, nodes_full)
        w = [PU.flat_sd(ag.online_network) for ag in reps]
        np.testing.assert_array_equal(w[0], w[1])
        np.testing.assert_array_equal(PU.flat_sd(reps[0].target_network), PU.flat_sd(reps[1].target_network))
        t = [ag.replay_memory_buffer.replay_buffer.tree for ag in reps]
        np.testing.assert_array_equal(t[0], t[1])
        if s == 0:
            assert abs(reps[0].last_loss() - full.last_loss()) / abs(full.last_loss()) < (1e-5 if precision == "fp32" else 1e-3)
        if precision == "fp32" and s == 0 and B <= 1024:      # ... and against the ORACLE's full-batch step (dqn/agent.py:245-272)
            tr = {}
            orc.step = 50
            orc.learn(u=u, trace=tr)
            orc.sync_target()
            np.testing.assert_array_equal(got, tr["nodes"])
            assert abs(reps[0].last_loss() - tr["loss"]) <= 1e-5 * abs(tr["loss"])
            g_ref = np.concatenate([tr["grads"][k].ravel() for k, _ in orc.online.named_parameters()])
            g_rep = reps[0]._lh.get_params(_lib.GRADS).cpu().numpy()
            pt = PU.per_tensor_max_rel(g_rep, g_ref, PU.tensor_sizes(orc.online))
            pt.pop("fc_val.bias", None)
            assert max(pt.values()) < 1e-5, pt
            well_o = np.abs(g_ref) >= 1e-6
            w_ref = PU.flat_sd(orc.online)
            assert R.max_rel(np.where(well_o, w[0], w_ref), w_ref) < 1e-5
            assert R.max_rel(PU.flat_sd(reps[0].target_network), PU.flat_sd(orc.target)) < 1e-4
        if precision == "fp32" and s == 0:
            w_full = PU.flat_sd(full.online_network)
            well = np.abs(g_full) >= 1e-6
            assert R.max_rel(np.where(well, w[0], w_full), w_full) < 1e-5
            t_full = full.replay_memory_buffer.replay_buffer.tree
            sa, sb = reps[0].replay_memory_buffer._ring.stats(), full.replay_memory_buffer._ring.stats()
            if B <= 1024:
                # shards and full batch run the same instantiation of the step kernel (one 4-row tile per CTA): same Q bits,
                # same |td|, same priorities -> the trees must be bit-identical
                np.testing.assert_array_equal(t[0], t_full)
                assert (sa.max_priority, sa.min_priority, sa.total_priority) == (sb.max_priority, sb.min_priority, sb.total_priority)
            else:
                # B = 8192: the full batch takes the batch-stationary row phase (rmc_rows_ws.cuh, >= 2 16-row tiles per CTA), the
                # 4096-row shards the 4-row tiles: Q differs in the last bits (K-sum order), so p = (|td| + eps)^alpha differs by
                # alpha * 1e-6 / eps ~ 1e-4 relative for the smallest |td|.  (Each path against the ORACLE: priorities <= 1 ulp,
                # tests/test_gpu_headline_sizes.py and test_large_batch_per_step_uses_grid_wide_tree_path.)
                np.testing.assert_allclose(t[0], t_full, rtol=5e-4, atol=0)
                assert abs(sa.total_priority - sb.total_priority) <= 5e-4 * sb.total_priority


# ------------------------------------------------------------------------------------------------------------------
# Real CUDA-IPC exchange between two PROCESSES on two GPUs (skipped on a one-GPU box): ShardedLearner(exchange="peer")
# maps the peers' exchange buffers through cudaIpcMemHandle_t all-gathered with torch.distributed (parallel.py
# _connect_ipc); replicas must stay bit-identical and equal the NCCL form's result at 1e-5.
_IPC_WORKER = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RMC_REPO"])
from tests import parity_utils as PU
from multimodal_drl_rmc_b200 import _lib
from multimodal_drl_rmc_b200.parallel import ShardedLearner, sharded_act
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
B, cap = 4096, 20000
out = {}
for precision in ("fp32", "bf16"):
    res = {}
    for exchange in ("peer", "nccl"):
        ag = PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap, seed=41, gpu=str(rank))[1]      # identical replicas on every rank
        ag.learn_precision = precision
        sl = ShardedLearner(ag, exchange=exchange)
        rng = np.random.default_rng(3)
        for s in range(3):
            ag.step = 50 + s
            sl.learn(u=rng.random(B))
        torch.cuda.synchronize()
        w = PU.flat_sd(ag.online_network)
        t = ag.replay_memory_buffer.replay_buffer.tree
        res[exchange] = (w, t, sl.exchange_status())
        gathered = [None] * world
        dist.all_gather_object(gathered, (w.tobytes(), t.tobytes()))
        res[exchange + "_identical"] = all(g == gathered[0] for g in gathered)
        del sl, ag
    out[precision] = dict(peer_identical=res["peer_identical"], nccl_identical=res["nccl_identical"], status=res["peer"][2],
                          peer_vs_nccl=float(np.max(np.abs(res["peer"][0] - res["nccl"][0]))),
                          trees_equal=bool(np.array_equal(res["peer"][1], res["nccl"][1])))
# C3 row split: every rank evaluates its slice, actions all-gathered
ag = PU.make_pair("DuelingDoubleDQNAgent", 14, 32, 200, 200, seed=2, gpu=str(rank))[1]
obs = np.random.default_rng(0).random((10001, 14), dtype=np.float32)
split = sharded_act(ag.online_network, obs)
out["act_split_equal"] = split == ag.online_network.actions(obs)
if rank == 0:
    print("IPC_RESULT " + json.dumps(out))
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (real CUDA IPC mapping between two processes)")
def test_peer_exchange_over_real_ipc_two_processes(tmp_path):
    import json
    import os
    import subprocess
    import sys
    script = tmp_path / "ipc_worker.py"
    script.write_text(_IPC_WORKER)
    env = dict(os.environ, RMC_REPO=R.GOLDEN_DIR.rsplit("/tests/", 1)[0])
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + "\n".join(l for l in res.stderr.splitlines() if not l.startswith("DEVICE"))[-4000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("IPC_RESULT ")][-1]
    out = json.loads(line[len("IPC_RESULT "):])
    print(out)
    for precision in ("fp32", "bf16"):
        o = out[precision]
        assert o["status"] == 0 and o["peer_identical"] and o["nccl_identical"], o
        assert o["trees_equal"] or precision == "bf16", o
        assert o["peer_vs_nccl"] <= (3e-4 if precision == "fp32" else 1e-3), o      # at most ~lr per step on ill-conditioned Adam elements
    assert out["act_split_equal"]


def test_sharded_act_row_split_single_rank():
    """BASELINE configs[2] across GPUs = rows split over the ranks; here the slicing logic with emulated
    ranks on one GPU: the concatenated slices equal the unsplit call."""
    from multimodal_drl_rmc_b200.parallel import sharded_act
    ag = PU.make_pair("DuelingDoubleDQNAgent", 14, 32, 200, 200, seed=2)[1]
    obs = np.random.default_rng(0).random((4099, 14), dtype=np.float32)
    whole = ag.online_network.actions(obs)
    for W in (1, 3, 8):
        parts, ranges = zip(*[sharded_act(ag.online_network, obs, gather=False, rank=r, world=W) for r in range(W)])
        assert [a for p in parts for a in p] == whole
        assert ranges[0][0] == 0 and ranges[-1][1] == 4099 and all(ranges[i][1] == ranges[i + 1][0] for i in range(W - 1))
