"""-m gpu: the repo-HEAD hybrid CNN + MLP network (env/dqn_config.py:66-193) through the same drop-in
Agent API / C ABI, against the oracle (whose restatement of the body is pinned bit-for-bit on a golden case produced by
the reference's own TwoStreamHybridNetwork class, tests/golden/learner_per_hybrid.npz).  Same bar as the macro MLP:
indices and tree bit-exact, Q / loss / gradients / post-Adam weights within 1e-5 (per-tensor max-norm relative)."""
import numpy as np
import pytest
import torch

from tests import parity_utils as PU
from tests import recipes as R

pytestmark = pytest.mark.gpu
TOL = 1e-5
D = 284


@pytest.mark.parametrize("algo,B,cap,fill,steps,soft,tf", [
    ("PerDuelingDoubleDQNAgent", 32, 300, 400, 5, True, 30000),       # steps 2.. replay the captured step graph (patched nodes)
    ("DuelingDoubleDQNAgent", 16, 128, 128, 2, True, 30000),
    ("DQNAgent", 24, 200, 150, 4, False, 2),                          # hard sync every 2nd step: two step shapes, two graphs
    ("PerDuelingDoubleDQNAgent", 130, 512, 512, 1, True, 30000),
])
def test_hybrid_learner_step_parity(algo, B, cap, fill, steps, soft, tf):
    res = PU.run_parity_case(algo, D, B, cap, fill, steps, seed=5, soft=soft, target_freq=tf, activation="elu", body="hybrid")
    print(res)
    assert res["nodes_equal"] and res["tree_equal"]
    assert res["max_pri_ulp"] <= 1.0
    assert res["max_rel_isw"] < 1e-6
    assert res["max_rel_q"] < TOL
    assert res["max_rel_loss"] < TOL
    assert res["max_rel_grads"] < TOL, res["worst_grad"]
    assert res["max_rel_weights"] < TOL, res["worst_w"]
    assert res["max_abs_weights_all"] <= 1e-4 * steps
    assert res["max_rel_target"] < 10 * TOL


def test_hybrid_act_q_and_checkpoint_roundtrip(tmp_path):
    orc, agent = PU.make_pair("DuelingDoubleDQNAgent", D, 8, 64, 64, seed=2, activation="elu", body="hybrid")
    states = np.random.default_rng(1).random((700, D), dtype=np.float32)
    with torch.no_grad():
        q_ref = orc.online(torch.as_tensor(states)).numpy()
    q_gpu = agent.online_network(torch.as_tensor(states)).cpu().numpy()
    assert R.max_rel(q_gpu, q_ref) < TOL
    assert agent.online_network.actions(states) == orc.greedy_actions(states)
    adv = agent.online_network.advantages(states).cpu().numpy()
    assert adv.shape == (700, 8)
    # .pack round trip: keys of the reference's hybrid checkpoints
    path = str(tmp_path / "h.pack")
    agent.online_network.save(path, 7, 3, 1.5, 90.0)
    keys = list(agent.online_network.state_dict().keys())
    assert keys[:2] == ["net.cnn_stream.0.weight", "net.cnn_stream.0.bias"] and keys[-1] == "fc_adv.bias"
    before = PU.flat_sd(agent.online_network)
    assert before.size == 885481
    agent.online_network.load_state_dict({k: torch.zeros_like(v) for k, v in agent.online_network.state_dict().items()})
    assert agent.online_network.load(path)[0] == 7
    np.testing.assert_array_equal(PU.flat_sd(agent.online_network), before)


def test_hybrid_rejects_modes_not_built():
    _, agent = PU.make_pair("DuelingDoubleDQNAgent", D, 8, 64, 64, seed=3, activation="elu", body="hybrid")
    agent.learn_precision = "bf16"
    with pytest.raises(Exception):
        agent.learn()                       # recorded ...
        agent.update_target_network()       # ... and refused when it is launched


def test_hybrid_sidecar_resume_diagnostics_and_epsilon_greedy(tmp_path):
    """The learner-state side-car, the on-device diagnostics (8f-4) and choose_actions work unchanged for
    the hybrid network; the step is deterministic (two identically seeded agents stay bit-identical)."""
    import random
    _, a = PU.make_pair("PerDuelingDoubleDQNAgent", D, 16, 128, 128, seed=12, activation="elu", body="hybrid")
    _, b = PU.make_pair("PerDuelingDoubleDQNAgent", D, 16, 128, 128, seed=12, activation="elu", body="hybrid")
    rng = np.random.default_rng(0)
    us = [rng.random(16) for _ in range(4)]
    for s in range(2):
        for ag in (a, b):
            ag.step = s
            ag.learn(u=us[s], fuse_target_update=True)
    np.testing.assert_array_equal(PU.flat_sd(a.online_network), PU.flat_sd(b.online_network))
    side = str(tmp_path / "state.npz")
    a.save_learner_state(side)
    z = torch.zeros(b._lh.n_params)
    for kind in (0, 1, 2, 3):
        b._lh.set_params(kind, z)
    b._adam_t = 0
    b.load_learner_state(side)
    for s in range(2, 4):
        for ag in (a, b):
            ag.step = s
            ag.learn(u=us[s], fuse_target_update=True)
    np.testing.assert_array_equal(PU.flat_sd(a.online_network), PU.flat_sd(b.online_network))
    np.testing.assert_array_equal(PU.flat_sd(a.target_network), PU.flat_sd(b.target_network))
    np.testing.assert_array_equal(a.replay_memory_buffer.replay_buffer.tree, b.replay_memory_buffer.replay_buffer.tree)
    d = a.diagnostics()
    assert np.isfinite(d["loss"]) and d["replay_size"] == 128
    obs = np.random.default_rng(3).random((4, D), dtype=np.float32)
    a.step = 10 ** 9                      # epsilon at its floor: mostly greedy
    random.seed(5)
    acts = a.choose_actions(obs)
    assert len(acts) == 4 and all(0 <= x < 8 for x in acts)
