"""The oracle (oracle/dqn_oracle.py) must reproduce the fixtures that the UNMODIFIED
reference produced (tests/golden/make_golden.py).  Bit-exact when the CPU/torch/numpy
fingerprint equals the one the goldens were generated on; 2e-6 max-norm otherwise
(integer outputs always exact where they do not depend on float rounding)."""
import os

import numpy as np
import pytest
import torch

from oracle import dqn_oracle as O
from tests import recipes as R

META = R.golden_meta()
SAME_CPU = META["cpu"] == R.cpu_fingerprint()
STRIDE = META["sample_stride"]


@pytest.mark.parametrize("name", sorted(META["cases"]))
def test_learner_case_matches_reference_golden(name):
    c = META["cases"][name]
    g = R.load_golden("learner_%s.npz" % name)
    lrn = R.build_oracle_case(c, META)
    per = c["algo"].startswith("Per")
    if SAME_CPU:
        assert R.sha(R.flat_params(lrn.online)) == str(g["init_online_sha"])
        assert R.sha(R.flat_params(lrn.target)) == str(g["init_target_sha"])
    for s in range(c["steps"]):
        lrn.step = R.step_number(s)
        inj = g["step%d_inject" % s]
        tr = {}
        if per:
            lrn.learn(u=inj, trace=tr)
        else:
            lrn.learn(indices=[int(i) for i in inj], trace=tr)
        lrn.sync_target()
        if per:
            if SAME_CPU or s == 0:
                np.testing.assert_array_equal(tr["nodes"], g["step%d_nodes" % s])
            np.testing.assert_allclose(tr["is_w"], g["step%d_is_w" % s], rtol=1e-12 if SAME_CPU else 1e-5)
            np.testing.assert_allclose(tr["abs_td"].reshape(-1), g["step%d_abs_td" % s], rtol=2e-6, atol=1e-7)
            t = lrn.replay.tree
            stats = np.array([t.total, t.max_leaf, t.min_leaf, t.size, t.data_pointer])
            np.testing.assert_allclose(stats, g["step%d_tree_stats" % s], rtol=1e-6)
            if SAME_CPU:
                assert R.sha(t.tree) == str(g["step%d_tree_sha" % s])
                np.testing.assert_array_equal(tr["abs_td"].reshape(-1), g["step%d_abs_td" % s])
        if s == 0:
            grads = np.concatenate([tr["grads"][k].ravel() for k, _ in lrn.online.named_parameters()])
            assert R.max_rel(grads[::STRIDE], g["step0_grads_sample"]) < 2e-6
            if SAME_CPU:
                assert R.sha(grads) == str(g["step0_grads_sha"])
    fo, ft = R.flat_params(lrn.online), R.flat_params(lrn.target)
    assert R.max_rel(fo[::STRIDE], g["final_online_sample"]) < 2e-6
    assert R.max_rel(ft[::STRIDE], g["final_target_sample"]) < 2e-6
    if SAME_CPU:
        assert R.sha(fo) == str(g["final_online_sha"])
        assert R.sha(ft) == str(g["final_target_sha"])
        if per and "final_tree" in g.files:
            np.testing.assert_array_equal(lrn.replay.tree.tree, g["final_tree"])
    probe = np.random.default_rng(5).random((64, c["D"]), dtype=np.float32)
    np.testing.assert_array_equal(np.asarray(lrn.greedy_actions(probe)), g["probe_actions"])


@pytest.mark.parametrize("cap", [1, 2, 3, 7, 64, 69, 1000])
def test_sumtree_stream_matches_reference_golden(cap):
    g = R.load_golden("sumtree_cap%d.npz" % cap)
    t = O.OracleSumTree(cap)
    for kind, leaf, val, expect in g["ops"]:
        kind = int(kind)
        if kind == 0:
            t.push(float(val), ("row", t.data_pointer))
        elif kind == 1:
            t.assign(int(leaf), float(val))
        else:
            node, _p, _row = t.descend(float(val))
            assert node == int(expect)
    np.testing.assert_array_equal(t.tree, g["tree"])
    np.testing.assert_array_equal(np.array([t.total, t.max_leaf, t.min_leaf, t.size, t.data_pointer]), g["stats"])


def test_numpy_restatement_agrees_with_autograd_oracle():
    """Appendix-A formulas (what the CUDA kernels implement) vs the torch-autograd oracle."""
    c = META["cases"]["per_d14"]
    lrn = R.build_oracle_case(c, META)
    lrn.step = 17
    on = {k: v.detach().numpy().copy() for k, v in lrn.online.state_dict().items()}
    tg = {k: v.detach().numpy().copy() for k, v in lrn.target.state_dict().items()}
    tr = {}
    u = np.random.default_rng(3).random(c["B"])
    lrn.learn(u=u, trace=tr)
    r = O.numpy_td_and_grads(on, tg, tr["obs"], tr["act"], tr["rew"], tr["done"], tr["nxt"],
                             tr["is_w"].astype(np.float32), 0.99)
    assert R.max_rel(r["q"], tr["q"]) < 1e-5
    assert R.max_rel(r["y"], tr["y"].reshape(-1)) < 1e-5
    assert abs(r["loss"] - tr["loss"]) <= 1e-5 * abs(tr["loss"])
    for k, gref in tr["grads"].items():
        assert R.max_rel(r["grads"][k], gref) < 1e-5, k
    # Adam restatement on identical (p, g, m, v, t): <= 2 ulp
    p0 = on["net.2.weight"]
    g0 = tr["grads"]["net.2.weight"]
    p1, _, _ = O.numpy_adam(p0, g0, np.zeros_like(p0), np.zeros_like(p0), 1)
    ref = lrn.online.state_dict()["net.2.weight"].numpy()
    assert np.max(np.abs(p1 - ref)) <= 2 * np.spacing(np.float32(np.max(np.abs(ref))))


def test_act_fixture_consistent_with_checkpoint():
    import msgpack
    import sys
    g = R.load_golden("act_macro_with_lane.npz")
    with open(os.path.join(R.GOLDEN_DIR, "macro_with_lane.pack"), "rb") as fh:
        raw = msgpack.loads(fh.read(), raw=False, strict_map_key=False)
    net = O.OracleQNet(14, 8, dueling=True)
    sd = {}
    for k, v in raw["parameters"].items():
        k = k.decode() if isinstance(k, bytes) else k
        if isinstance(v, np.ndarray):   # the reference (if imported earlier in this process) monkey-patches msgpack
            arr = v
        else:
            arr = np.frombuffer(v[b"data"], dtype=np.dtype(v[b"type"])).reshape(v[b"shape"])
        sd[k] = torch.as_tensor(arr.copy())
    net.load_state_dict(sd)
    assert raw["step"] == int(g["meta"][0])
    np.testing.assert_array_equal(np.asarray(net.greedy(g["states"])), g["actions"])
    with torch.no_grad():
        assert R.max_rel(net(torch.as_tensor(g["states"])).numpy(), g["q"]) < 2e-6


def test_numpy_adam_bit_matches_torch():
    """The rounding sequence the CUDA Adam implements == torch.optim.Adam on CPU, bit for bit
    (<= 0.01 % of elements may differ by 1 ulp: float64 emulation of fmaf double-rounds)."""
    rng = np.random.default_rng(0)
    n = 100000
    p0 = (rng.normal(size=n) * 0.1).astype(np.float32)
    P = torch.nn.Parameter(torch.tensor(p0.copy()))
    opt = torch.optim.Adam([P], lr=1e-4)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for t in range(1, 5):
        g = (rng.normal(size=n) * 10.0 ** rng.uniform(-9, -1, size=n)).astype(np.float32)
        P.grad = torch.tensor(g.copy())
        opt.step()
        p, m, v = O.numpy_adam(p, g, m, v, t)
        assert np.array_equal(m, opt.state[P]["exp_avg"].numpy())
        assert np.array_equal(v, opt.state[P]["exp_avg_sq"].numpy())
        ref = P.detach().numpy()
        bad = p != ref
        assert bad.mean() < 1e-4
        assert np.all(np.abs(p[bad] - ref[bad]) <= 2 * np.spacing(np.abs(ref[bad])))
        p = ref.copy()
