"""-m gpu: batched greedy act / Q values vs the reference's outputs on the shipped trained checkpoint."""
import os

import numpy as np
import pytest
import torch

from oracle import dqn_oracle as O
from tests import parity_utils as PU
from tests import recipes as R

pytestmark = pytest.mark.gpu


def _net():
    from multimodal_drl_rmc_b200 import Networks
    from multimodal_drl_rmc_b200.macro_config import ObsSpace, network_config
    net = Networks.DuelingDeepQNetwork(torch.device("cuda:0"), 1e-4, network_config, ObsSpace(14), 8)
    meta = net.load(os.path.join(R.GOLDEN_DIR, "macro_with_lane.pack"))
    return net, meta


def test_act_matches_reference_on_trained_checkpoint():
    g = R.load_golden("act_macro_with_lane.npz")
    net, meta = _net()
    assert meta[0] == int(g["meta"][0]) and meta[1] == int(g["meta"][1])
    assert net.actions(g["states"]) == g["actions"].tolist()
    assert net.actions(g["states"].tolist()) == g["actions"].tolist()          # list-of-lists input, like evaluate.py:28
    assert net.actions([g["states"][3].tolist()])[0] == int(g["actions"][3])    # n = 1
    q = net(torch.as_tensor(g["states"])).cpu().numpy()
    assert R.max_rel(q, g["q"]) < 1e-5
    dev_actions = net.actions(torch.as_tensor(g["states"], device="cuda:0"))
    assert dev_actions == g["actions"].tolist()


def test_act_65536_states_vs_oracle():
    net, _ = _net()
    orc = O.OracleQNet(14, 8, dueling=True)
    orc.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    states = np.random.default_rng(0).random((65536, 14), dtype=np.float32)
    ref = np.asarray(orc.greedy(states))
    got = np.asarray(net.actions(states))
    # fp32 summation-order differences can flip only exact near-ties; SURVEY 7.2 measured 0 flips in 65,536
    assert (ref != got).mean() <= 1e-4
    with torch.no_grad():
        adv = orc.fc_adv(orc.net(torch.as_tensor(states))).numpy()
    top2 = np.sort(adv, axis=1)[:, -2:]
    flips = np.nonzero(ref != got)[0]
    assert all((top2[i, 1] - top2[i, 0]) < 1e-5 for i in flips)


@pytest.mark.parametrize("dueling,D,n", [(True, 14, 2048 + 5), (False, 8, 1024), (True, 8, 4096 + 63)])
def test_q_values_through_the_64_row_kernel(dueling, D, n):
    """n >= 1024 takes k_mlp_infer64 (8 x 8 register tiles, K of layer 2 split over the two halves of the CTA):
    Q values within 1e-5 of the oracle, greedy actions equal except exact near-ties; ragged last tile."""
    from multimodal_drl_rmc_b200 import Networks
    from multimodal_drl_rmc_b200.macro_config import ObsSpace, network_config
    torch.manual_seed(5)
    cls = Networks.DuelingDeepQNetwork if dueling else Networks.DeepQNetwork
    net = cls(torch.device("cuda:0"), 1e-4, network_config, ObsSpace(D), 8)
    orc = O.OracleQNet(D, 8, dueling=dueling)
    orc.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    x = np.random.default_rng(2).random((n, D), dtype=np.float32)
    with torch.no_grad():
        qref = orc(torch.as_tensor(x)).numpy()
    q = net(torch.as_tensor(x)).cpu().numpy()
    assert q.shape == qref.shape
    assert R.max_rel(q, qref) < 1e-5
    ref, got = np.asarray(orc.greedy(x)), np.asarray(net.actions(x))
    assert (ref != got).mean() <= 1e-3


@pytest.mark.parametrize("dueling,D", [(True, 14), (False, 8)])
def test_per_env_step_act_path_equals_the_batched_kernels(dueling, D):
    """n <= 32 host states take k_act_tiny (states in the kernel-argument buffer, actions through mapped host memory):
    same greedy actions as the oracle and as the general copy path, for every n around the 8-row tile and the 32-row limit."""
    from multimodal_drl_rmc_b200 import Networks
    from multimodal_drl_rmc_b200.macro_config import ObsSpace, network_config
    torch.manual_seed(7)
    cls = Networks.DuelingDeepQNetwork if dueling else Networks.DeepQNetwork
    net = cls(torch.device("cuda:0"), 1e-4, network_config, ObsSpace(D), 8)
    orc = O.OracleQNet(D, 8, dueling=dueling)
    orc.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    x = np.random.default_rng(4).random((200, D), dtype=np.float32)
    ref = orc.greedy(x)
    assert net.actions(x) == ref                                   # n = 200: copy path, 8-row kernel
    for n in (1, 2, 7, 8, 9, 16, 31, 32, 33):
        for off in (0, 50):
            assert net.actions(x[off:off + n]) == ref[off:off + n], (n, off)
    for _ in range(300):                                           # epoch hand-shake over many back-to-back calls
        assert net.actions(x[:3]) == ref[:3]


def test_plain_head_act_and_q():
    from multimodal_drl_rmc_b200 import Networks
    from multimodal_drl_rmc_b200.macro_config import ObsSpace, network_config
    torch.manual_seed(3)
    net = Networks.DeepQNetwork(torch.device("cuda:0"), 1e-4, network_config, ObsSpace(8), 8)
    orc = O.OracleQNet(8, 8, dueling=False)
    orc.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    x = np.random.default_rng(1).random((1000, 8), dtype=np.float32)
    with torch.no_grad():
        qref = orc(torch.as_tensor(x)).numpy()
    assert R.max_rel(net(torch.as_tensor(x)).cpu().numpy(), qref) < 1e-5
    assert net.actions(x) == orc.greedy(x)


def test_value_and_advantages_streams_and_torch_ops():
    g = R.load_golden("act_macro_with_lane.npz")
    net, _ = _net()
    x = torch.as_tensor(g["states"], device="cuda:0")
    adv = net.advantages(x).cpu().numpy()
    val = net.value(x).cpu().numpy()
    assert R.max_rel(adv, g["adv"]) < 1e-5
    q = val + (adv - adv.mean(axis=1, keepdims=True))
    assert R.max_rel(q, g["q"]) < 1e-5
    # the same entry points through torch.ops
    from multimodal_drl_rmc_b200 import ops  # noqa: F401  (registers torch.ops.rmc_b200.*)
    h = net._standalone_handle().handle.value
    a = torch.ops.rmc_b200.act(h, x)
    assert a.tolist() == g["actions"].tolist()
    qq = torch.ops.rmc_b200.q_values(h, 0, x, 8)
    assert R.max_rel(qq.cpu().numpy(), g["q"]) < 1e-5


@pytest.mark.parametrize("activation", ["relu", "elu"])
def test_tensor_core_act_mode_within_stated_bound(activation):
    """tcgen05 / bf16-operand mode: Q within 1e-2 max-norm-relative of the exact fp32 kernel, greedy actions
    equal except near-ties (the north star's 'stated looser bound' for tensor-core modes).  ELU = the repo-HEAD
    activation (env/dqn_config.py:175) on the same trained weights."""
    from multimodal_drl_rmc_b200 import _lib
    if activation == "relu":
        net, _ = _net()
    else:
        from multimodal_drl_rmc_b200 import Networks
        from multimodal_drl_rmc_b200.macro_config import ObsSpace, network_config_elu
        net = Networks.DuelingDeepQNetwork(torch.device("cuda:0"), 1e-4, network_config_elu, ObsSpace(14), 8)
        net.load(os.path.join(R.GOLDEN_DIR, "macro_with_lane.pack"))
    lh = net._standalone_handle()
    n = 65536 + 77           # ragged last tile
    states = np.random.default_rng(3).random((n, 14), dtype=np.float32)
    x = torch.as_tensor(states, device="cuda:0")
    exact = torch.empty(n, 9, device="cuda:0")
    tc = torch.empty(n, 9, device="cuda:0")
    _lib.check(_lib.lib().rmc_learner_heads(lh.handle, 0, x.data_ptr(), n, exact.data_ptr(), _lib.stream_ptr()))
    _lib.check(_lib.lib().rmc_learner_heads_tc(lh.handle, x.data_ptr(), n, tc.data_ptr(), _lib.stream_ptr()))
    e, t = exact.cpu().numpy(), tc.cpu().numpy()
    err = R.max_rel(t, e)
    print("tensor-core heads max-norm rel err", err)
    assert err < 1e-2
    a_exact = np.asarray(net.actions(x))
    a_tc = np.asarray(net.actions(x, precision="bf16"))
    flips = np.nonzero(a_exact != a_tc)[0]
    print("action disagreement", len(flips) / n)
    assert len(flips) / n < 0.02
    top2 = np.sort(e[:, 1:], axis=1)[:, -2:]
    gap = top2[flips, 1] - top2[flips, 0]
    assert np.all(gap < 0.05 * np.abs(e[:, 1:]).max())


@pytest.mark.parametrize("D", [16, 15, 9])
def test_tensor_core_act_bias_paths(D):
    """The biases of the hidden layers ride on the MMAs (csrc/rmc_tc.cuh: b0 in column 15 of the W0 image when obs_dim <= 15, b2 as
    one more K step); obs_dim = 16 has no spare column and adds b0 in the epilogue.  Randomly initialised nets (non-zero
    biases), biases scaled up so that a dropped or doubled bias would be far outside the bound."""
    from multimodal_drl_rmc_b200 import Networks, _lib
    from multimodal_drl_rmc_b200.macro_config import ObsSpace, network_config
    torch.manual_seed(5)
    net = Networks.DuelingDeepQNetwork(torch.device("cuda:0"), 1e-4, network_config, ObsSpace(D), 8)
    sd = net.state_dict()
    for k in sd:
        if k.endswith("bias"):
            sd[k] = sd[k] * 8.0
    net.load_state_dict(sd)
    lh = net._standalone_handle()
    n = 4096 + 5
    x = torch.as_tensor(np.random.default_rng(D).random((n, D), dtype=np.float32), device="cuda:0")
    exact = torch.empty(n, 9, device="cuda:0")
    tc = torch.empty(n, 9, device="cuda:0")
    _lib.check(_lib.lib().rmc_learner_heads(lh.handle, 0, x.data_ptr(), n, exact.data_ptr(), _lib.stream_ptr()))
    _lib.check(_lib.lib().rmc_learner_heads_tc(lh.handle, x.data_ptr(), n, tc.data_ptr(), _lib.stream_ptr()))
    err = R.max_rel(tc.cpu().numpy(), exact.cpu().numpy())
    print("obs_dim", D, "tensor-core heads max-norm rel err", err)
    assert err < 1e-2


def test_device_epsilon_greedy_distribution_and_reproducibility():
    """Agent.exploration = "device": greedy act + Philox epsilon-greedy in one call."""
    _, a = PU.make_pair("DuelingDoubleDQNAgent", 14, 32, 64, 64, seed=3)
    obs = np.random.default_rng(0).random((20000, 14), dtype=np.float32)
    greedy = np.asarray(a.online_network.actions(obs))
    a.exploration = "device"
    a.epsilon_exp_decay = False
    a.epsilon_start, a.epsilon_min, a.epsilon_decay = 0.3, 0.3, 1.0          # epsilon() == 0.3
    got = np.asarray(a.choose_actions(obs))
    changed = got != greedy
    # a row explores with probability 0.3 and then hits its own greedy action 1/8 of the time
    assert abs(changed.mean() - 0.3 * 7 / 8) < 0.015
    assert got.min() >= 0 and got.max() < 8
    a.epsilon_start = a.epsilon_min = 1.0                                       # every row explores: uniform over the 8 actions
    hist = np.bincount(np.asarray(a.choose_actions(obs)), minlength=8) / obs.shape[0]
    assert np.all(np.abs(hist - 0.125) < 0.012)
    a.epsilon_start = a.epsilon_min = 0.0
    assert a.choose_actions(obs) == greedy.tolist()
    _, b = PU.make_pair("DuelingDoubleDQNAgent", 14, 32, 64, 64, seed=3)
    b.exploration, b.epsilon_exp_decay = "device", False
    b.epsilon_start, b.epsilon_min, b.epsilon_decay = 0.3, 0.3, 1.0
    np.testing.assert_array_equal(np.asarray(b.choose_actions(obs)), got)       # same seed and call count -> same draw
