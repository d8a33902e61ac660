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
