#!/usr/bin/env python
"""Host-side time of each public-API call of one e2e step (diagnostic): store_transitions / learn /
update_target_network / last_loss, wall-clock per call, plus the same loop without the loss read-back."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

wl = bench.WORKLOADS["per256"]
agent, (obs, act, rew, done, nxt) = bench.build_gpu_agent(wl, 0, 0)
N = 3000
acc = np.zeros(4)
for k in range(N + 200):
    j = k % 4096
    t0 = time.perf_counter()
    agent.store_transitions(obs[j:j + 1], [int(act[j])], [float(rew[j])], [bool(done[j])], nxt[j:j + 1], None)
    t1 = time.perf_counter()
    agent.step += 1
    agent.learn()
    t2 = time.perf_counter()
    agent.update_target_network()
    t3 = time.perf_counter()
    agent.last_loss()
    t4 = time.perf_counter()
    if k >= 200:
        acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
print("per-call host time (us): store %.1f  learn %.1f  update_target %.1f  last_loss(sync) %.1f  total %.1f" % (*(1e6 * acc / N), 1e6 * acc.sum() / N))
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(N):
    j = k % 4096
    agent.store_transitions(obs[j:j + 1], [int(act[j])], [float(rew[j])], [bool(done[j])], nxt[j:j + 1], None)
    agent.step += 1
    agent.learn()
    agent.update_target_network()
torch.cuda.synchronize()
print("no per-step sync: %.1f us/step" % (1e6 * (time.perf_counter() - t0) / N))
t0 = time.perf_counter()
for k in range(N):
    agent.step += 1
    agent.learn()
    agent.update_target_network()
torch.cuda.synchronize()
print("learn only, no sync: %.1f us/step" % (1e6 * (time.perf_counter() - t0) / N))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for k in range(2000):
    j = k % 4096
    agent.store_transitions(obs[j:j + 1], [int(act[j])], [float(rew[j])], [bool(done[j])], nxt[j:j + 1], None)
    agent.step += 1
    agent.learn()
    agent.update_target_network()
    agent.last_loss()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
