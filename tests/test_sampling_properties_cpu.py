"""CPU property tests of the stratified sum-tree sampler (dqn/replay_memory.py:69-92, dqn/utils/sum_tree.py:42-61) that
the CUDA write-back relies on: the team write-back elects the writer of a repeatedly drawn leaf by adjacency
(csrc/rmc_tree.cuh, tree_update_team), which is valid iff equal leaves are adjacent in the batch -- they are, because the drawn leaves move left to right
through the tree as the sample index grows."""
import numpy as np
import pytest

from oracle import dqn_oracle as O


def _filled(cap, fill, seed, spiky):
    rng = np.random.default_rng(seed)
    mem = O.OraclePrioritizedReplay(cap, 1, 1000)
    for k in range(fill):
        mem.tree.push(1.0, (k,))
    first_leaf = cap - 1
    pri = np.power(np.minimum(np.abs(rng.standard_normal(min(fill, cap))).astype(np.float32) + np.float32(1e-4), np.float32(1.0)),
                   np.float32(0.6)).astype(np.float32)
    if spiky:   # a few heavy leaves: many strata land on the same leaf (duplicates in one batch)
        pri[:] = np.float32(1e-3)
        pri[rng.integers(0, len(pri), 5)] = np.float32(1.0)
    for k, p in enumerate(pri):
        mem.tree.assign(first_leaf + k, p)
    return mem


def _inorder_rank(n_nodes):
    rank, stack, k, cur = {}, [], 0, 0
    while stack or cur < n_nodes:
        while cur < n_nodes:
            stack.append(cur)
            cur = 2 * cur + 1
        cur = stack.pop()
        rank[cur] = k
        k += 1
        cur = 2 * cur + 2
    return rank


@pytest.mark.parametrize("cap,fill,batch,spiky", [(1000, 1300, 64, False), (37, 37, 32, False), (333, 200, 30, False),
                                                  (4096, 4096, 590, True), (512, 100, 256, True)])
def test_stratified_draws_are_sorted_and_duplicates_adjacent(cap, fill, batch, spiky):
    mem = _filled(cap, fill, 5, spiky)
    mem.batch = batch
    rng = np.random.default_rng(9)
    for trial in range(4):
        u = rng.random(batch)
        if trial == 1:
            u[:] = 0.0            # every stratum at its lower edge
        if trial == 2:
            u[:] = np.nextafter(1.0, 0.0)   # ... and at its upper edge
        _, nodes, _ = mem.sample(0, u=u)
        nodes = np.asarray(nodes)
        # left-to-right (in-order) position of every drawn leaf: non-decreasing in the sample index.  (The heap INDEX is
        # not monotone for capacities that are not powers of two: the deepest level's leaves come first.)
        rank = _inorder_rank(2 * cap - 1)
        assert np.all(np.diff([rank[int(n)] for n in nodes]) >= 0), "drawn leaves must move left to right with the sample index"
        # therefore 'last sample in batch order wins' == 'sample i writes iff leaf[i+1] != leaf[i]'
        last_by_scan = {int(n): i for i, n in enumerate(nodes)}
        writers = [i for i in range(batch) if i + 1 == batch or nodes[i + 1] != nodes[i]]
        assert sorted(last_by_scan.values()) == writers
        if spiky:
            assert len(writers) < batch, "the case is meant to contain duplicates"


def test_stratum_value_never_exceeds_its_upper_edge():
    # v = lo + (hi - lo) * u <= hi for u < 1: hi - lo is exact (Sterbenz) and rounding is monotone
    rng = np.random.default_rng(3)
    for total in (1.0, 3.7e-3, 12345.678, float(np.float32(0.1)) * 999983):
        for batch in (30, 256, 65536):
            seg = total / batch
            i = rng.integers(0, batch, 2000).astype(np.float64)
            lo, hi = seg * i, seg * (i + 1)
            for u in (np.nextafter(1.0, 0.0), 0.5, rng.random(2000)):
                v = lo + (hi - lo) * u
                assert np.all(v <= hi) and np.all(v >= lo)
