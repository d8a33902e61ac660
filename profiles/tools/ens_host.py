"""Is the 8-agent ensemble launch host-bound?  Host issue time per AgentEnsemble.learn() against the drained device time
(measured: 9.7 us of host time per 73 us step -- no).   usage (GPU box): python profiles/tools/ens_host.py"""
import sys, time
sys.path.insert(0, '.')
import torch, bench
from multimodal_drl_rmc_b200.parallel import AgentEnsemble
wl = dict(bench.WORKLOADS["per256"])
members = [bench.build_gpu_agent(wl, 0, seed=50 + k)[0] for k in range(8)]
ens = AgentEnsemble(members)
def stepe():
    for m in members: m.step += 1
    ens.learn()
for _ in range(50): stepe()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): stepe()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue time per step %.1f us; drained after another %.1f us per step" % ((t1 - t0) / 300 * 1e6, (t2 - t1) / 300 * 1e6))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(300): stepe()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
