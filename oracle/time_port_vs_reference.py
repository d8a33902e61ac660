"""TEST INFRASTRUCTURE ONLY -- build-container measurement behind ``cpu_baseline.port_vs_reference`` in bench.py.

bench.py's CPU arm times the oracle PORT (oracle/dqn_oracle.py), because the unmodified reference cannot travel to the GPU
box.  This script times both on the same cores and the same workload (BASELINE configs[1]: PER + double + dueling, B = 256,
cap = size = 1,000,000, replay filled directly like bench.build_cpu_learner does), so the ratio port / reference can be
stated next to every GPU-vs-CPU number.  Run where /root/reference exists:

    python oracle/time_port_vs_reference.py [steps=60] [threads=nproc]
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import refharness  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    wl = bench.WORKLOADS["per256"]
    # ---- the port, exactly as bench.py times it
    lrn = bench.build_cpu_learner(wl, threads)
    done, dt = bench.time_cpu_learner(lrn, steps, 5, budget_s=120.0)
    port_ms = 1e3 * dt / done
    del lrn
    # ---- the unmodified reference classes (dqn/agent.py:275-320), same fill
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    agent = refharness.make_reference_agent(wl["algo"], bench.D, wl["B"], wl["cap"], tempfile.mkdtemp())
    n = wl["size"]
    obs, act, rew, done_, nxt = bench.synthetic(n, 20251018)
    rows = list(zip(list(obs), act.tolist(), rew.tolist(), (done_ != 0).tolist(), list(nxt)))
    t = agent.replay_memory_buffer.replay_buffer                # dqn/utils/sum_tree.py SumTree
    cap = t.capacity
    t.data[:n] = np.array(rows + [None], dtype=object)[:n]
    t.size, t.data_pointer = n, n % cap
    t.tree[cap - 1:cap - 1 + n] = bench.seeded_priorities(n, 7)
    for L in range(int(np.floor(np.log2(cap - 1))), -1, -1):
        first, last = (1 << L) - 1, min((1 << (L + 1)) - 2, cap - 2)
        idx = np.arange(first, last + 1)
        t.tree[idx] = t.tree[2 * idx + 1] + t.tree[2 * idx + 2]
    leaves = t.tree[cap - 1:cap - 1 + n]
    t.max_priority_index, t.min_priority_index = int(np.argmax(leaves)) + cap - 1, int(np.argmin(leaves)) + cap - 1
    np.random.seed(1000)
    for s in range(5):
        agent.step = s
        agent.learn()
        agent.update_target_network()
    t0 = time.perf_counter()
    k = 0
    for s in range(steps):
        agent.step = 5 + s
        agent.learn()
        agent.update_target_network()
        k += 1
        if time.perf_counter() - t0 > 120.0:
            break
    ref_ms = 1e3 * (time.perf_counter() - t0) / k
    print("threads %d | oracle port %.2f ms/step (%d steps) | unmodified reference %.2f ms/step (%d steps) | port / reference = %.3f"
          % (threads, port_ms, done, ref_ms, k, port_ms / ref_ms))


if __name__ == "__main__":
    main()
