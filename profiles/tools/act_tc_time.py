"""tcgen05 act on 65,536 states: microseconds per call launched call by call, and GPU-bound inside a CUDA graph of 20 calls.
usage (GPU box): python profiles/tools/act_tc_time.py"""
import sys, os
sys.path.insert(0, '.')
import numpy as np, torch, ctypes as C
import bench
from multimodal_drl_rmc_b200 import _lib
agent, _ = bench.build_gpu_agent(dict(bench.WORKLOADS["per256"], size=4096, cap=4096), 0, 0)
lib = _lib.lib()
n = 65536
states = torch.as_tensor(np.random.default_rng(0).random((n, 14), dtype=np.float32), device=agent.device)
acts = torch.empty(n, dtype=torch.int64, device=agent.device)
def f():
    _lib.check(lib.rmc_learner_act_tc(agent._lh.handle, states.data_ptr(), n, acts.data_ptr(), _lib.stream_ptr()))
for _ in range(10): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for rep in range(5):
    e0.record()
    for _ in range(200): f()
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 200 * 1e3)
print("act_tc 65536 us/call (weights packed once):", ["%.2f" % t for t in ts])
# the same through a CUDA graph of 20 calls (takes the host's launch rate out of the measurement)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    f()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(20):
            f()
torch.cuda.synchronize()
ts = []
for rep in range(5):
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 200 * 1e3)
print("act_tc 65536 us/call inside a CUDA graph of 20 calls:", ["%.2f" % t for t in ts])
