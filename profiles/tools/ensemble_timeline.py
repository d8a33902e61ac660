#!/usr/bin/env python
"""Per-CTA phase timeline of the 8-agent ensemble launch (k_learner_step, grid.y = agent)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from multimodal_drl_rmc_b200 import _lib  # noqa: E402
from multimodal_drl_rmc_b200.parallel import AgentEnsemble  # noqa: E402

wl = dict(bench.WORKLOADS["per256"], size=200_000, cap=200_000)
members = [bench.build_gpu_agent(wl, 0, seed=50 + k)[0] for k in range(8)]
lib = _lib.lib()
for m in members:            # before the group is created: the group keeps a copy of every member's context
    lib.rmc_learner_debug_timing(m._lh.handle, 1)
ens = AgentEnsemble(members)


def step():
    for m in members:
        m.step += 1
    ens.learn()


for _ in range(20):
    step()
names = ["start", "sampled", "tgt_w_landed", "tgt_pass", "onl_w_landed", "rows_done", "past_barrier", "done",
         "s8", "s9", "s10", "s11", "s12", "s13"]
for k in (0, 7):
    acc = []
    for it in range(6):
        step()
        buf = np.zeros(1024 * 32, np.uint64)
        torch.cuda.synchronize()
        # raw read of the whole debug buffer of member k
        n = C.c_int32()
        members[k]._lh  # keep alive
        _lib.check(lib.rmc_learner_debug_read_sync(members[k]._lh.handle, buf.ctypes.data, 1024, C.byref(n), _lib.stream_ptr()))
        t = buf.reshape(1024, 32)[:, :14].astype(np.int64)
        rows = t[(t[:, 0] > 0)]
        acc.append(rows)
    a = acc[-1]
    t0 = a[:, 0].min()
    a = np.where(a > 0, a - t0, -1)
    print("agent", k, "CTAs with stamps:", a.shape[0])
    for j, nm in enumerate(names):
        col = a[:, j][a[:, j] >= 0]
        if len(col):
            print("  %-16s min %8d  median %8d  max %8d" % (nm, col.min(), np.median(col), col.max()))
