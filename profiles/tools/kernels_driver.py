#!/usr/bin/env python
"""Runs each kernel of the path a few times on the BASELINE shapes (for ncu captures):
k_per_sample (65,536 stratified samples + row gather from the 1M-row ring), k_mlp_infer (fp32 act, 65,536),
k_mlp_infer_tc (tcgen05 act, 65,536), k_learner_step (PER B=256)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from multimodal_drl_rmc_b200 import _lib  # noqa: E402

agent, _ = bench.build_gpu_agent(bench.WORKLOADS["per256"], 0, 0)
lib = _lib.lib()
dev = agent.device
ring = agent.replay_memory_buffer._ring
B = 65536
nodes = torch.empty(B, dtype=torch.int64, device=dev)
w = torch.empty(B, dtype=torch.float32, device=dev)
rows = torch.empty(B, ring.row_floats, dtype=torch.float32, device=dev)
states = torch.as_tensor(np.random.default_rng(0).random((B, 14), dtype=np.float32), device=dev)
acts = torch.empty(B, dtype=torch.int64, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(name, fn, algo_bytes=None, flops=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    extra = ""
    if algo_bytes:
        extra += "  %.1f GB/s algorithmic" % (algo_bytes / (ms * 1e-3) / 1e9)
    if flops:
        extra += "  %.1f TFLOP/s" % (flops / (ms * 1e-3) / 1e12)
    print("%-16s %.4f ms%s" % (name, ms, extra))


timed("k_per_sample", lambda: _lib.check(lib.rmc_per_sample(ring.handle, B, 0.5, None, 7, 1, nodes.data_ptr(), w.data_ptr(), rows.data_ptr(), _lib.stream_ptr())),
      algo_bytes=B * (4 * 31 + 8 * 21))
timed("k_mlp_infer", lambda: _lib.check(lib.rmc_learner_act(agent._lh.handle, states.data_ptr(), B, acts.data_ptr(), _lib.stream_ptr())),
      flops=B * 2 * (14 * 256 + 256 * 128 + 128 * 8))
timed("k_mlp_infer_tc", lambda: _lib.check(lib.rmc_learner_act_tc(agent._lh.handle, states.data_ptr(), B, acts.data_ptr(), _lib.stream_ptr())),
      flops=B * 2 * (16 * 256 + 256 * 128 + 128 * 16))


def step():
    agent.step += 1
    agent.learn(fuse_target_update=True)


timed("k_learner_step", step, algo_bytes=256 * 628, flops=256 * 367872)
