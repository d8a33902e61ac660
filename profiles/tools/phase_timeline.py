#!/usr/bin/env python
"""Per-CTA phase timeline of k_learner_step from its in-kernel %globaltimer stamps (diagnostic).
usage (GPU box): python profiles/tools/phase_timeline.py [per256|default32] > gpurun_out/timeline.txt"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from multimodal_drl_rmc_b200 import _lib  # noqa: E402

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "per256"])
if len(sys.argv) > 2:
    wl["B"] = int(sys.argv[2])
agent, _ = bench.build_gpu_agent(wl, 0, 0)
lib = _lib.lib()
for _ in range(20):
    agent.step += 1
    agent.learn()
    agent.update_target_network()
lib.rmc_learner_debug_timing(agent._lh.handle, 1)
names = ["start", "sampled", "tgt_w_landed", "tgt_pass", "onl_w_landed", "rows_done", "past_barrier", "done",
         "s8:top_synced|pri_done", "s9:descent_start|stamped", "s10:descent_end|applied", "s11:pow_done|extremes", "s12:row_stored|fenced", "s13:online_fwd_done",
         "s14:td_done", "s15:dz2_done", "s16:unit_start", "s17:unit_staged", "s18:unit_summed", "s19:unit_adam_done"]
NS = len(names)
acc = []
for it in range(10):
    agent.step += 1
    agent.learn()
    agent.update_target_network()
    buf = np.zeros(1024 * 32, np.uint64)
    n = C.c_int32()
    _lib.check(lib.rmc_learner_debug_read_sync(agent._lh.handle, buf.ctypes.data, 1024, C.byref(n), _lib.stream_ptr()))
    t = buf[: n.value * 32].reshape(n.value, 32)[:, :NS].astype(np.int64)
    t0 = t[:, 0][t[:, 0] > 0].min()
    acc.append(np.where(t > 0, t - t0, -1))
a = np.stack(acc[2:])
print("grid", a.shape[1], "CTAs; ns since the first CTA started (median over", a.shape[0], "launches)")
med = np.median(a, axis=0)
print("%-28s %10s %10s %10s" % ("stamp", "min", "median", "max"))
for k, nm in enumerate(names):
    col = med[:, k][med[:, k] >= 0]
    if len(col):
        print("%-28s %10.0f %10.0f %10.0f" % (nm, col.min(), np.median(col), col.max()))
print("per-CTA rows (first 8, last 2):")
for c in list(range(min(8, med.shape[0]))) + [64, 65, 100, 136, 137] + list(range(max(8, med.shape[0] - 2), med.shape[0])):
    print(c, " ".join("%7.0f" % x for x in med[c]))
order = np.argsort(-med[:, 7])
print("slowest CTAs (by done):")
for c in order[:6]:
    print(c, " ".join("%7.0f" % x for x in med[c]))
if med.shape[0] == 148 and wl.get("B") == 256:
    print("streamed phase B (B = 256): row CTAs 0-63 and target CTAs 64-127 own one W2 unit each, 128-139 reduce W0/b0/heads, 140-147 tree team")
    for nm, lo, hi in (("row CTAs", 0, 64), ("target CTAs", 64, 128), ("reduce CTAs", 128, 140), ("tree team", 140, 148)):
        d = med[lo:hi, 7]
        print("  %-12s done: median %7.0f  max %7.0f" % (nm, np.median(d), d.max()))
if False:
    # phase-B duration (barrier exit -> done) per kind of work: CTAs 0-7 own the W0 units (16x32), 8-135 the W2 units
    # (16x16), 136-139 the head units (32x16), 140-147 are the priority write-back team
    print("phase B: ns from barrier exit (team: from the |td| flags) to done, by kind of unit (median / max over CTAs)")
    for nm, lo, hi in (("W0 16x32", 0, 8), ("W2 16x16", 8, 136), ("heads 32x16", 136, 140), ("tree team", 140, 148)):
        d = med[lo:hi, 7] - med[lo:hi, 6]
        print("  %-12s %7.0f %7.0f   done at %7.0f (max)" % (nm, np.median(d), d.max(), med[lo:hi, 7].max()))
