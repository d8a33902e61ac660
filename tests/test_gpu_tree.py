"""-m gpu: device sum tree / replay ring vs the oracle and the reference-generated golden streams."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import dqn_oracle as O
from tests import recipes as R

pytestmark = pytest.mark.gpu


def _row(i, D=4):
    rng = np.random.default_rng(i)
    return (rng.random(D, dtype=np.float32), int(i % 8), float(np.float32(i) * 0.5), bool(i % 3 == 0), rng.random(D, dtype=np.float32))


@pytest.mark.parametrize("cap", [1, 2, 3, 7, 64, 69, 1000])
def test_sumtree_golden_stream_through_dropin(cap):
    """The op stream recorded from the reference SumTree (tests/golden/sumtree_cap*.npz), replayed
    through the drop-in SumTree.add/update/get_leaf: leaves, tree array and stats bit-exact."""
    from multimodal_drl_rmc_b200 import SumTree
    g = R.load_golden("sumtree_cap%d.npz" % cap)
    t = SumTree(cap)
    ops = g["ops"] if cap < 1000 else g["ops"][:1500]
    ref = O.OracleSumTree(cap)
    for n, (kind, leaf, val, expect) in enumerate(ops):
        kind = int(kind)
        if kind == 0:
            t.add(float(val), _row(n))
            ref.push(float(val), None)
        elif kind == 1:
            t.update(int(leaf), float(val))
            ref.assign(int(leaf), float(val))
        else:
            node, p, _row_ = t.get_leaf(float(val))
            assert node == int(expect)
            assert p == ref.tree[node]
    np.testing.assert_array_equal(t.tree, ref.tree)
    assert (t.total_priority, t.max_priority, t.min_priority, t.size, t.data_pointer) == \
        (ref.total, ref.max_leaf, ref.min_leaf, ref.size, ref.data_pointer)
    if cap < 1000:
        np.testing.assert_array_equal(t.tree, g["tree"])


def _filled(cap, n, D=14, seed=0):
    from multimodal_drl_rmc_b200 import ReplayMemoryPrioritized
    mem = ReplayMemoryPrioritized(cap, 64, 2e6)
    obs, act, rew, done, nxt = O.synthetic_transitions(n, D, seed)
    list(mem.store_transitions(obs, act.tolist(), rew.tolist(), done.astype(bool).tolist(), nxt))
    ref = O.OraclePrioritizedReplay(cap, 64, 2e6)
    for i in range(n):
        list(ref.store([obs[i]], [int(act[i])], [float(rew[i])], [bool(done[i])], [nxt[i]]))
    return mem, ref, (obs, act, rew, done, nxt)


@pytest.mark.parametrize("cap,n", [(1000, 1000), (1000, 2600), (5000, 4999), (70000, 70000), (100000, 150000)])
def test_bulk_and_incremental_push_equal_oracle(cap, n):
    """store_transitions through the bulk path (n > 4096 rows -> bottom-up rebuild) and the incremental
    path must both give the oracle's tree, size and data_pointer; rows land in ring order."""
    mem, ref, data = _filled(cap, n)
    t = mem.replay_buffer
    np.testing.assert_array_equal(t.tree, ref.tree.tree)
    assert (t.size, t.data_pointer) == (ref.tree.size, ref.tree.data_pointer)
    assert (t.total_priority, t.max_priority, t.min_priority) == (ref.tree.total, ref.tree.max_leaf, ref.tree.min_leaf)
    # ring content: slot k holds the last transition written there
    obs = data[0]
    rows = mem._ring.read_rows(0, min(cap, 50))
    for k in range(min(cap, 50, n)):
        last = k + ((n - 1 - k) // cap) * cap
        np.testing.assert_array_equal(rows[k, :14], obs[last])


@pytest.mark.parametrize("B", [64, 256, 5000, 20000])
def test_per_sample_and_writeback_match_oracle(B):
    cap = 20000
    mem, ref, _ = _filled(cap, cap, seed=3)
    mem.batch_size = ref.batch = B
    rng = np.random.default_rng(B)
    # non-degenerate priorities on both sides
    pri = np.power(np.minimum(np.abs(rng.normal(size=cap)).astype(np.float32) + np.float32(1e-4), np.float32(1.0)), np.float32(0.6)).astype(np.float32)
    for i in range(cap):
        ref.tree.assign(i + cap - 1, pri[i])
    from multimodal_drl_rmc_b200 import _lib
    dev = torch.device("cuda", mem._ring.device_index)
    _lib.check(_lib.lib().rmc_replay_set_priorities(mem._ring.handle, torch.as_tensor(pri, device=dev).data_ptr(), cap, _lib.stream_ptr()))
    np.testing.assert_array_equal(mem.replay_buffer.tree, ref.tree.tree)
    for it in range(3):
        u = rng.random(B)
        w_ref, nodes_ref, rows_ref = ref.sample(5000 * it, u=u)
        w, nodes, rows = mem.sample_transitions(5000 * it, u=u)
        assert nodes == [int(x) for x in nodes_ref], "tree indices must be bit-exact for injected uniforms"
        np.testing.assert_allclose(np.asarray(w), np.asarray(w_ref, np.float32).astype(np.float64), rtol=1e-6)
        np.testing.assert_array_equal(rows[7][0], rows_ref[7][0])
        assert rows[7][1:4] == tuple(rows_ref[7][1:4])
        # write-back from |td| with duplicates (stratified sampling repeats leaves when B ~ size)
        abs_td = np.abs(rng.normal(size=(B, 1))).astype(np.float32)
        p_ref = np.power(np.minimum(abs_td + np.float32(1e-4), np.float32(1.0)), np.float32(0.6)).reshape(-1)
        nodes_t = torch.as_tensor(np.asarray(nodes, np.int64), device=dev)
        td_t = torch.as_tensor(abs_td.reshape(-1), device=dev)
        p_out = torch.empty(B, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().rmc_per_update_from_td(mem._ring.handle, nodes_t.data_ptr(), td_t.data_ptr(), B, 1e-4, 0.6, 1.0,
                                                     p_out.data_ptr(), _lib.stream_ptr()))
        p_gpu = p_out.cpu().numpy()
        ulp = np.abs(p_gpu.astype(np.float64) - p_ref.astype(np.float64)) / np.spacing(p_ref)
        assert ulp.max() <= 1.0
        # bit-exact half of the contract: same float32 priorities on both sides -> identical trees
        ref.write_back_priorities(nodes_ref, [np.float32(x) for x in p_gpu])
        np.testing.assert_array_equal(mem.replay_buffer.tree, ref.tree.tree)
        st = mem._ring.stats()
        assert (st.total_priority, st.max_priority, st.min_priority) == (ref.tree.total, ref.tree.max_leaf, ref.tree.min_leaf)


def test_uniform_replay_sampling_and_fifo():
    from multimodal_drl_rmc_b200 import ReplayMemoryNaive
    mem = ReplayMemoryNaive(100, 16)
    obs, act, rew, done, nxt = O.synthetic_transitions(250, 8, 1)
    ends = list(mem.store_transitions(obs, act.tolist(), rew.tolist(), done.astype(bool).tolist(), nxt))
    assert ends == [i for i in range(250) if done[i]]
    assert len(mem.replay_buffer) == 100
    np.testing.assert_array_equal(mem.replay_buffer[0][0], obs[150])     # oldest surviving transition
    np.testing.assert_array_equal(mem.replay_buffer[99][4], nxt[249])
    tr = mem.sample_transitions(indices=list(range(16)))
    for k in range(16):
        np.testing.assert_array_equal(tr[k][0], obs[150 + k])
        assert tr[k][1] == int(act[150 + k]) and tr[k][2] == float(rew[150 + k]) and tr[k][3] == bool(done[150 + k])
    tr2 = mem.sample_transitions()
    assert len(tr2) == 16


def test_store_transitions_is_lazy_like_the_reference():
    from multimodal_drl_rmc_b200 import ReplayMemoryNaive
    mem = ReplayMemoryNaive(10, 2)
    obs, act, rew, done, nxt = O.synthetic_transitions(3, 8, 1)
    gen = mem.store_transitions(obs, act.tolist(), rew.tolist(), [False, False, False], nxt)
    assert len(mem.replay_buffer) == 0          # nothing happens until the generator is driven
    list(gen)
    assert len(mem.replay_buffer) == 3
