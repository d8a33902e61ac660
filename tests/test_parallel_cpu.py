"""not-gpu: host logic of the multi-GPU forms, world_size 2 over gloo.
The sharding rule (global stratified segments, 1/B_global loss scale, gradient all-reduce, replicated
write-back) is checked against the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_drl_rmc_b200.parallel import shard_range
from oracle import dqn_oracle as O


def test_shard_ranges_tile_the_batch():
    for B in (1, 7, 256, 65536, 65537):
        for W in (1, 2, 3, 4, 8):
            spans = [shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(seed=3, B=64, cap=400):
    torch.set_num_threads(1)
    torch.manual_seed(seed)
    lrn = O.OracleLearner("PerDuelingDoubleDQNAgent", 14, 8, B, cap)
    obs, act, rew, done, nxt = O.synthetic_transitions(cap, 14, 99)
    for i in range(cap):
        lrn.store([obs[i]], [int(act[i])], [float(rew[i])], [bool(done[i])], [nxt[i]])
    pri = np.power(np.minimum(np.abs(np.random.default_rng(1).normal(size=cap)).astype(np.float32) + np.float32(1e-4), 1.0), np.float32(0.6))
    for i in range(cap):
        lrn.replay.tree.assign(i + cap - 1, np.float32(pri[i]))
    return lrn


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 64
    lrn = _build(B=B)
    u = np.random.default_rng(5).random(B)
    lo, hi = shard_range(B, rank, world)
    # every replica draws the GLOBAL batch's strata [lo, hi) (global segment length), then scales by 1/B_global
    t = lrn.replay.tree
    seg = t.total / B
    beta = lrn.replay.beta(17)
    max_w = pow(t.size * (t.min_leaf / t.total), -beta)
    nodes, w, rows = [], [], []
    for i in range(lo, hi):
        v = seg * i + (seg * (i + 1) - seg * i) * u[i]
        n, p, row = t.descend(v)
        nodes.append(n)
        w.append(pow(t.size * (p / t.total), -beta) / max_w)
        rows.append(row)
    obs, act, rew, done, nxt = lrn._tensorize(rows)
    on = {k: v.detach().numpy().copy() for k, v in lrn.online.state_dict().items()}
    tg = {k: v.detach().numpy().copy() for k, v in lrn.target.state_dict().items()}
    r = O.numpy_td_and_grads(on, tg, obs.numpy(), act.numpy(), rew.numpy(), done.numpy(), nxt.numpy(), np.asarray(w, np.float32), 0.99)
    scale = (hi - lo) / B     # numpy_td_and_grads normalises by the local batch: rescale to 1/B_global
    flat = torch.as_tensor(np.concatenate([r["grads"][k].ravel() for k in on]) * np.float32(scale))
    loss = torch.tensor([float(r["loss"]) * scale])
    dist.all_reduce(flat)
    dist.all_reduce(loss)
    gathered = [None] * world
    dist.all_gather_object(gathered, (nodes, r["abs_td"].tolist()))
    if rank == 0:
        ret["grads"] = flat.numpy()
        ret["loss"] = float(loss)
        ret["nodes"] = sum((g[0] for g in gathered), [])
        ret["abs_td"] = sum((g[1] for g in gathered), [])
    dist.destroy_process_group()


def test_sharded_step_equals_single_process_oracle():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    lrn = _build(B=64)
    lrn.step = 17
    tr = {}
    lrn.learn(u=np.random.default_rng(5).random(64), trace=tr)
    ref = np.concatenate([tr["grads"][k].ravel() for k, _ in lrn.online.named_parameters()])
    assert ret["nodes"] == tr["nodes"].tolist(), "union of the shards' strata must be the single-process batch"
    assert np.max(np.abs(ret["grads"] - ref)) / np.max(np.abs(ref)) < 1e-5
    assert abs(ret["loss"] - tr["loss"]) / abs(tr["loss"]) < 1e-5
    np.testing.assert_allclose(np.asarray(ret["abs_td"], np.float32), tr["abs_td"].reshape(-1), rtol=1e-5, atol=1e-7)
