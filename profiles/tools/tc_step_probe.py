#!/usr/bin/env python
"""A few B = 65,536 learner steps in the tensor-core (bf16 / tcgen05) mode: prints ms/step; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split of the step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
agent, _ = bench.build_gpu_agent(dict(bench.WORKLOADS["per256"], B=B), 0, 12)
agent.learn_precision = prec


def step():
    agent.step += 1
    agent.learn()
    agent.update_target_network()


for _ in range(3):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    step()
e.record()
torch.cuda.synchronize()
print("B=%d precision=%s: %.4f ms/step, loss %.6f" % (B, prec, s.elapsed_time(e) / reps, agent.last_loss()))
