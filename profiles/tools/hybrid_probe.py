#!/usr/bin/env python
"""Learner step of the repo-HEAD hybrid CNN + MLP network (env/dqn_config.py:66-193): us/step on the GPU path and, with
--cpu, the oracle port of the reference learner on the host cores.  usage: hybrid_probe.py [B] [algo] [--cpu]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from multimodal_drl_rmc_b200 import macro_config  # noqa: E402
from oracle.dqn_oracle import synthetic_transitions  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if len(args) > 0 else 32
algo = args[1] if len(args) > 1 else "DuelingDoubleDQNAgent"
D, N = macro_config.HYBRID_OBS_DIM, 20000
tmp = tempfile.mkdtemp(prefix="rmc_hyb_")
obs, act, rew, done, nxt = synthetic_transitions(N, D, 20251018)
if "--cpu" in sys.argv:
    from oracle.dqn_oracle import OracleLearner
    torch.set_num_threads(os.cpu_count())
    orc = OracleLearner(algo, D, 8, B, N, activation="elu", body="hybrid")
    for i in range(N):
        orc.store([obs[i]], [int(act[i])], [float(rew[i])], [bool(done[i])], [nxt[i]])
    for _ in range(3):
        orc.learn(); orc.sync_target()
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        orc.learn(); orc.sync_target()
    ms = 1e3 * (time.perf_counter() - t0) / K
    print("CPU oracle port (%d threads) hybrid %s B=%d: %.2f ms/step, %.0f transitions/s" % (torch.get_num_threads(), algo, B, ms, B / (ms * 1e-3)))
    sys.exit(0)
agent = macro_config.make_agent(algo, D, B, N, save_dir=tmp + "/", log_dir=tmp + "/", activation="hybrid")
agent.replay_memory_buffer._ring.push_host(obs, act, rew, done, nxt)
torch.cuda.synchronize()


def step():
    agent.step += 1
    agent.learn(fuse_target_update=True)


for _ in range(5):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 100
s.record()
for _ in range(K):
    step()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / K
print("GPU hybrid %s B=%d: %.1f us/step, %.0f transitions/s, loss %.5f" % (algo, B, 1e3 * ms, B / (ms * 1e-3), agent.last_loss()))
