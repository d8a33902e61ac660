#!/usr/bin/env python
"""Back-to-back learner steps: in-kernel [first CTA start, last CTA end] per launch -> kernel span and the idle gap
between consecutive launches (diagnostic)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from multimodal_drl_rmc_b200 import _lib  # noqa: E402

agent, _ = bench.build_gpu_agent(bench.WORKLOADS["per256"], 0, 0)
lib = _lib.lib()
for _ in range(50):
    agent.step += 1
    agent.learn(fuse_target_update=True)
torch.cuda.synchronize()
lib.rmc_learner_debug_timing(agent._lh.handle, 1)
for _ in range(48):
    agent.step += 1
    agent.learn(fuse_target_update=True)
buf = np.zeros(128, np.uint64)
_lib.check(lib.rmc_learner_debug_gaps_sync(agent._lh.handle, buf.ctypes.data, _lib.stream_ptr()))
se = buf.reshape(64, 2).astype(np.int64)
se = se[(se[:, 1] > 0) & (se[:, 0] < 2**62)]
se = se[np.argsort(se[:, 0])]
span = se[:, 1] - se[:, 0]
gap = se[1:, 0] - se[:-1, 1]
period = se[1:, 0] - se[:-1, 0]
print("launches", len(se))
print("kernel span  (first CTA start -> last CTA end) ns: median %d  min %d  max %d" % (np.median(span), span.min(), span.max()))
print("idle gap     (last CTA end -> next first CTA start) ns: median %d  min %d  max %d" % (np.median(gap), gap.min(), gap.max()))
print("period ns: median %d" % np.median(period))
