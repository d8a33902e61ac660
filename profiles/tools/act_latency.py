#!/usr/bin/env python
"""Latency of the per-env-step greedy act through the public API (Agent.choose_actions with epsilon = 0, host RNG path
and the one-call device path) for n_env = 1, 4, 16 host state vectors.  usage (GPU box): python profiles/tools/act_latency.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

HYBRID = len(sys.argv) > 1 and sys.argv[1] == "hybrid"      # the repo-HEAD CNN + MLP network (state[284]) instead of the macro MLP
if HYBRID:
    import tempfile
    from multimodal_drl_rmc_b200 import macro_config
    tmp = tempfile.mkdtemp(prefix="rmc_act_")
    bench.D = macro_config.HYBRID_OBS_DIM
    agent = macro_config.make_agent("DuelingDoubleDQNAgent", bench.D, 32, 4096, save_dir=tmp + "/", log_dir=tmp + "/", activation="hybrid")
else:
    wl = dict(bench.WORKLOADS["per256"])
    wl["size"] = wl["cap"] = 4096
    agent, _ = bench.build_gpu_agent(wl, 0, 0)
agent.epsilon_start = agent.epsilon_min = 0.0
rng = np.random.default_rng(0)
for mode in ("host", "device"):
    agent.exploration = mode
    for n_env in (1, 4, 16):
        x = rng.random((n_env, bench.D), dtype=np.float32)
        for _ in range(50):
            agent.choose_actions(x)
        t0 = time.perf_counter()
        for _ in range(2000):
            agent.choose_actions(x)
        us = 1e6 * (time.perf_counter() - t0) / 2000
        print("choose_actions exploration=%-6s n_env=%2d: %.1f us per call (host states in, python list of actions out)" % (mode, n_env, us))
