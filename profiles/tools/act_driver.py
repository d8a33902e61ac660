#!/usr/bin/env python
"""Runs the exact fp32 batched act (k_mlp_infer64) on 65,536 states a few times (for ncu captures / timing).
usage (GPU box): python profiles/tools/act_driver.py [n_calls]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from multimodal_drl_rmc_b200 import _lib  # noqa: E402

wl = dict(bench.WORKLOADS["per256"])
wl["size"] = 4096
wl["cap"] = 4096
agent, _ = bench.build_gpu_agent(wl, 0, 0)
lib = _lib.lib()
states = torch.as_tensor(np.random.default_rng(0).random((65536, bench.D), dtype=np.float32), device=agent.device)
acts = torch.empty(65536, dtype=torch.int64, device=agent.device)
n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_calls + 1)]
ev[0].record()
for k in range(n_calls):
    _lib.check(lib.rmc_learner_act(agent._lh.handle, states.data_ptr(), 65536, acts.data_ptr(), _lib.stream_ptr()))
    ev[k + 1].record()
torch.cuda.synchronize()
print("k_mlp_infer64, 65,536 states: ms per call", ["%.4f" % ev[k].elapsed_time(ev[k + 1]) for k in range(n_calls)])
