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
    for a, b in zip(solo, team):
        np.testing.assert_array_equal(PU.flat_sd(a.online_network), PU.flat_sd(b.online_network))
        np.testing.assert_array_equal(PU.flat_sd(a.target_network), PU.flat_sd(b.target_network))
        if per:
            np.testing.assert_array_equal(a.replay_memory_buffer.replay_buffer.tree, b.replay_memory_buffer.replay_buffer.tree)
            sa, sb = a.replay_memory_buffer._ring.stats(), b.replay_memory_buffer._ring.stats()
            assert (sa.max_priority, sa.min_priority, sa.total_priority) == (sb.max_priority, sb.min_priority, sb.total_priority)


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
    full = PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, cap, seed=41)[1]
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
        if precision == "fp32" and s == 0:
            w_full = PU.flat_sd(full.online_network)
            well = np.abs(g_full) >= 1e-6
            assert R.max_rel(np.where(well, w[0], w_full), w_full) < 1e-5
            np.testing.assert_array_equal(t[0], full.replay_memory_buffer.replay_buffer.tree)
            sa, sb = reps[0].replay_memory_buffer._ring.stats(), full.replay_memory_buffer._ring.stats()
            assert (sa.max_priority, sa.min_priority, sa.total_priority) == (sb.max_priority, sb.min_priority, sb.total_priority)
