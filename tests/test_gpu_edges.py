"""-m gpu: edge cases and error behaviour of the drop-in path (empty / tiny / ragged inputs, misuse)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import dqn_oracle as O
from tests import parity_utils as PU

pytestmark = pytest.mark.gpu


def _agent(algo, B, cap):
    import tempfile
    from multimodal_drl_rmc_b200 import macro_config
    tmp = tempfile.mkdtemp(prefix="rmc_edge_")
    return macro_config.make_agent(algo, 14, B, cap, save_dir=tmp + "/", log_dir=tmp + "/")


def test_learn_on_empty_replay_raises():
    a = _agent("PerDuelingDoubleDQNAgent", 32, 100)
    with pytest.raises(RuntimeError, match="empty"):
        a.learn()


def test_uniform_batch_larger_than_population_raises_like_random_sample():
    from multimodal_drl_rmc_b200._lib import RmcError
    a = _agent("DuelingDoubleDQNAgent", 32, 100)
    obs, act, rew, done, nxt = O.synthetic_transitions(10, 14, 1)
    a.store_transitions(obs, act.tolist(), rew.tolist(), done.astype(bool).tolist(), nxt, None)
    with pytest.raises(RmcError, match="larger than population"):   # reference: random.sample raises ValueError
        a.learn()


def test_per_with_fewer_transitions_than_batch_samples_duplicates_and_stays_exact():
    """PER samples with replacement: B = 64 from 5 stored transitions is legal in the reference."""
    res = PU.run_parity_case("PerDuelingDoubleDQNAgent", 14, 64, 50, 5, 2, seed=3)
    assert res["nodes_equal"] and res["tree_equal"]
    assert res["max_rel_q"] < 1e-5 and res["max_rel_grads"] < 1e-5


def test_capacity_one_and_batch_one():
    res = PU.run_parity_case("PerDuelingDoubleDQNAgent", 14, 1, 1, 3, 2, seed=4)
    assert res["nodes_equal"] and res["tree_equal"] and res["max_rel_q"] < 1e-5
    res = PU.run_parity_case("DQNAgent", 8, 1, 4, 9, 2, seed=4)
    assert res["max_rel_q"] < 1e-5 and res["max_rel_grads"] < 1e-5


def test_episode_bookkeeping_of_store_transitions():
    a = _agent("DuelingDoubleDQNAgent", 4, 100)
    obs, act, rew, done, nxt = O.synthetic_transitions(3, 14, 1)
    infos = [{"r": 1.0, "l": 5}, {"r": 2.0, "l": 6}, {"r": 3.0, "l": 7}]
    a.store_transitions(obs, act.tolist(), rew.tolist(), [False, True, True], nxt, infos)
    assert a.episode_count == 2 and [e["r"] for e in a.ep_info_buffer] == [2.0, 3.0]
    a.store_transitions(obs, act.tolist(), rew.tolist(), [True, True, True], nxt, None)   # falsy infos: no bookkeeping
    assert a.episode_count == 2
    assert len(a.replay_memory_buffer.replay_buffer) == 6


def test_choose_actions_epsilon_greedy_uses_python_random_like_the_reference():
    import random
    orc, a = PU.make_pair("DuelingDoubleDQNAgent", 14, 32, 200, 200, seed=6)
    x = np.random.default_rng(0).random((16, 14), dtype=np.float32)
    a.step = orc.step = 10**7               # epsilon = epsilon_min = 0.01
    random.seed(5)
    got = a.choose_actions(x)
    random.seed(5)
    ref = orc.greedy_actions(x)
    eps = float(np.exp(np.interp(10**7, [0, 2e6], [np.log(1.0), np.log(0.01)])))
    for i in range(len(ref)):
        if random.random() <= eps:
            ref[i] = random.randint(0, 7)
    assert got == ref
    assert abs(a.epsilon() - 0.01) < 1e-12


def test_checkpoint_roundtrip_through_the_agent(tmp_path):
    orc, a = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 32, 300, 300, seed=8)
    for s in range(3):
        a.step = s
        a.learn()
        a.update_target_network()
    a.step, a.save_frequency, a.resume_step = 10, 10, 0
    a.save_path = str(tmp_path / "ck" / "model.pack")
    a.save_model()
    b = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 32, 300, 300, seed=9)[1]
    b.save_path, b.load = a.save_path, True
    b.load_model()
    assert b.step == 10 and b.resume_step == 10
    np.testing.assert_array_equal(PU.flat_sd(a.online_network), PU.flat_sd(b.online_network))
    np.testing.assert_array_equal(PU.flat_sd(b.online_network), PU.flat_sd(b.target_network))   # target <- online on load
    x = np.random.default_rng(1).random((64, 14), dtype=np.float32)
    assert a.online_network.actions(x) == b.online_network.actions(x)


def test_sidecar_state_gives_bit_identical_resume(tmp_path):
    """with the Adam/target side-car a resumed learner continues exactly (the reference restarts Adam)."""
    _, a = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 64, 600, 600, seed=12)
    _, b = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 64, 600, 600, seed=12)
    rng = np.random.default_rng(0)
    us = [rng.random(64) for _ in range(6)]
    for s in range(3):
        for ag in (a, b):
            ag.step = s
            ag.learn(u=us[s], fuse_target_update=True)
    side = str(tmp_path / "state.npz")
    a.save_learner_state(side)
    # c: fresh learner object on b's replay state is not possible (replay is not in the side-car), so resume INTO b's
    # replay by scrambling b's learner state first, then restoring it from a's side-car
    z = torch.zeros(b._lh.n_params)
    for kind in (0, 1, 2, 3):
        b._lh.set_params(kind, z)
    b._adam_t = 0
    b.load_learner_state(side)
    for s in range(3, 6):
        for ag in (a, b):
            ag.step = s
            ag.learn(u=us[s], fuse_target_update=True)
    np.testing.assert_array_equal(PU.flat_sd(a.online_network), PU.flat_sd(b.online_network))
    np.testing.assert_array_equal(PU.flat_sd(a.target_network), PU.flat_sd(b.target_network))


def test_diagnostics_match_oracle_trace():
    orc, a = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 64, 800, 800, seed=15)
    u = np.random.default_rng(0).random(64)
    orc.step = a.step = 123
    tr = {}
    orc.learn(u=u, trace=tr)
    a.learn(u=u)
    d = a.diagnostics()
    assert abs(d["loss"] - tr["loss"]) <= 1e-5 * abs(tr["loss"])
    assert abs(d["abs_td_mean"] - float(tr["abs_td"].mean())) <= 1e-5
    assert abs(d["abs_td_max"] - float(tr["abs_td"].max())) <= 1e-5 * float(tr["abs_td"].max())
    assert abs(d["q_mean"] - float(tr["q_sa"].mean())) <= 1e-5
    assert d["replay_size"] == 800 and d["beta"] == float(orc.replay.beta(123))
    a.log_diagnostics()


def test_sidecar_with_replay_resumes_a_fresh_process_bit_identically(tmp_path):
    """SURVEY 8f-3 with the replay snapshot: a FRESH agent (empty replay, new weights) that loads the side-car continues
    exactly like the agent that wrote it -- device-side sampling (no injected uniforms), ring wrapped, PER tree included."""
    B, cap = 64, 700
    _, a = PU.make_pair("PerDuelingDoubleDQNAgent", 14, B, cap, 1000, seed=12)       # 1000 > cap: the ring has wrapped
    for s in range(3):
        a.step = s
        a.learn()
        a.update_target_network()
    side = str(tmp_path / "state")                       # no suffix: both calls must agree on the file name
    a.save_learner_state(side)
    b = _agent("PerDuelingDoubleDQNAgent", B, cap)
    b.load_learner_state(side)
    sa, sb = a.replay_memory_buffer._ring.stats(), b.replay_memory_buffer._ring.stats()
    assert (sa.size, sa.data_pointer, sa.total_priority, sa.max_priority, sa.min_priority) == (sb.size, sb.data_pointer, sb.total_priority, sb.max_priority, sb.min_priority)
    np.testing.assert_array_equal(a.replay_memory_buffer.replay_buffer.tree, b.replay_memory_buffer.replay_buffer.tree)
    obs, act, rew, done, nxt = O.synthetic_transitions(40, 14, 77)
    for s in range(3, 7):
        for ag in (a, b):
            ag.step = s
            ag.store_transitions(obs[s:s + 1], [int(act[s])], [float(rew[s])], [bool(done[s])], nxt[s:s + 1], None)
            ag.learn()
            ag.update_target_network()
        np.testing.assert_array_equal(PU.gpu_out(a, "nodes", torch.int64), PU.gpu_out(b, "nodes", torch.int64))
        assert a.last_loss() == b.last_loss()
    np.testing.assert_array_equal(PU.flat_sd(a.online_network), PU.flat_sd(b.online_network))
    np.testing.assert_array_equal(PU.flat_sd(a.target_network), PU.flat_sd(b.target_network))
    np.testing.assert_array_equal(a.replay_memory_buffer.replay_buffer.tree, b.replay_memory_buffer.replay_buffer.tree)
    # a side-car of another agent flavour is refused before anything is touched
    c = _agent("DuelingDoubleDQNAgent", B, cap)
    with pytest.raises(ValueError):
        c.load_learner_state(side)


def test_sidecar_restores_host_rng_streams(tmp_path):
    import random
    _, a = PU.make_pair("DuelingDoubleDQNAgent", 14, 32, 300, 300, seed=5)
    random.seed(11)
    np.random.seed(11)
    a.save_learner_state(str(tmp_path / "s.npz"))
    want = (random.random(), float(np.random.random_sample()))
    random.seed(99)
    np.random.seed(99)
    a.load_learner_state(str(tmp_path / "s.npz"))
    assert (random.random(), float(np.random.random_sample())) == want


def test_foreign_tree_indices_are_rejected_not_written():
    """index lists from outside must not corrupt device memory.  The Python mirror raises (the reference would
    raise IndexError); the C entry skips and counts entries outside the leaf range."""
    from multimodal_drl_rmc_b200 import _lib
    _, a = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 32, 100, 100, seed=6)
    mem = a.replay_memory_buffer
    tree0 = mem.replay_buffer.tree
    with pytest.raises(IndexError):
        mem.update_batch_priorities([99, 5, 250], np.array([0.5, 0.5, 0.5], np.float32))
    nodes = torch.as_tensor([99 + 3, -7, 10 ** 9, 99 + 3, 98], dtype=torch.int64, device=a.device)
    pri = torch.as_tensor([0.25, 0.5, 0.5, 0.75, 0.5], dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().rmc_per_update(mem._ring.handle, nodes.data_ptr(), pri.data_ptr(), 5, _lib.stream_ptr(a.device.index)))
    tree1 = mem.replay_buffer.tree
    assert tree1[99 + 3] == 0.75 and mem._ring.stats().rejected_nodes == 3
    changed = np.nonzero(tree0 != tree1)[0]
    path = []
    n = 99 + 3
    while True:
        path.append(n)
        if n == 0:
            break
        n = (n - 1) // 2
    assert set(changed.tolist()) <= set(path)
    assert tree1[0] == tree1[99:].sum()


def test_step_health_is_reported_and_default_launch_is_pdl():
    import os
    from multimodal_drl_rmc_b200 import _lib
    _, a = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 64, 500, 500, seed=7)
    for s in range(20):
        a.step = s
        a.learn()
        a.update_target_network()
    bad = C.c_uint32(123)
    _lib.check(_lib.lib().rmc_learner_status(a._lh.handle, C.byref(bad)))
    assert bad.value == 0 and np.isfinite(a.last_loss())
    assert os.environ.get("RMC_LAUNCH") in (None, "", "pdl"), "the test-suite runs the library's default launch mode"


def test_two_streams_share_the_device_safely():
    """Fused-step launches of one device on different streams are serialised by the library (an event edge at each stream
    switch), so two agents stepped from two torch streams cannot interleave half-scheduled grids; results equal the
    single-stream run."""
    pairs = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, 64, 500, 500, seed=30 + k)[1] for k in range(2)]
    refs = [PU.make_pair("PerDuelingDoubleDQNAgent", 14, 64, 500, 500, seed=30 + k)[1] for k in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    rng = np.random.default_rng(0)
    for s in range(30):
        us = [rng.random(64), rng.random(64)]
        for k in range(2):
            refs[k].step = pairs[k].step = s
            refs[k].learn(u=us[k])
            refs[k].update_target_network()
            with torch.cuda.stream(streams[k]):
                pairs[k].learn(u=us[k])
                pairs[k].update_target_network()
    torch.cuda.synchronize()
    for k in range(2):
        np.testing.assert_array_equal(PU.flat_sd(refs[k].online_network), PU.flat_sd(pairs[k].online_network))


def test_overlapped_actor_loop_against_synthetic_vec_env():
    """train.py's loop with n_env = 4 against the SubprocVecEnv stand-in; the overlapped form (learner step
    launched between step_async and step_wait) takes the same number of learner steps and fills the replay identically."""
    import tempfile
    from multimodal_drl_rmc_b200 import actor_loop, macro_config
    from multimodal_drl_rmc_b200.synthetic import SyntheticVecEnv
    runs = {}
    for name, loop in (("strict", actor_loop.strict_loop), ("overlapped", actor_loop.overlapped_loop)):
        tmp = tempfile.mkdtemp(prefix="rmc_loop_")
        torch.manual_seed(0)
        ag = macro_config.make_agent("PerDuelingDoubleDQNAgent", 14, 32, 5000, save_dir=tmp + "/", log_dir=tmp + "/", n_env=4, min_mem=200,
                                     log_freq=50, save_freq=10 ** 9)
        ag.exploration = "device"
        env = SyntheticVecEnv(4, 14, step_seconds=0.0002, episode_len=25, seed=1)
        assert actor_loop.init_replay_memory_buffer(ag, env, lambda: 3) == 200
        calls0 = ag._learn_calls
        assert loop(ag, env, 120) == 480
        torch.cuda.synchronize()
        st = ag.replay_memory_buffer._ring.stats()
        runs[name] = (ag._learn_calls - calls0, int(st.size), ag.episode_count, ag.last_loss())
        env.close()
    assert runs["strict"][:2] == runs["overlapped"][:2] == (120, 200 + 480)
    assert runs["strict"][2] == runs["overlapped"][2] > 0          # Monitor-style infos reached ep_info_buffer
    assert np.isfinite(runs["strict"][3]) and np.isfinite(runs["overlapped"][3])


def test_torch_op_learner_step_matches_agent_learn():
    from multimodal_drl_rmc_b200 import _lib, ops  # noqa: F401
    _, a = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 64, 500, 500, seed=8)
    _, b = PU.make_pair("PerDuelingDoubleDQNAgent", 14, 64, 500, 500, seed=8)
    u = np.random.default_rng(0).random(64)
    a.step = b.step = 3
    a.learn(u=u)
    ut = torch.as_tensor(u, device=b.device)
    loss = torch.ops.rmc_b200.learner_step(b._lh.handle.value, b.replay_memory_buffer._ring.handle.value, b.device.index, 64, _lib.PH_LEARN, b._beta(3), ut, None,
                                           0, 1, 1)
    b._lh.version[0] += 1
    assert float(loss) == a.last_loss()
    np.testing.assert_array_equal(PU.flat_sd(a.online_network), PU.flat_sd(b.online_network))
