#!/bin/bash
# Step times of the multi-tile launches (8-agent ensembles, B = 1,024 / 8,192 / 65,536 exact fp32) and the phase stamps of the
# B = 65,536 step (csrc/rmc_rows_ws.cuh).   usage (GPU box): bash profiles/tools/ws_probe.sh > gpurun_out/ws_probe.txt 2>&1
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python - <<'PY' 2>&1 | grep -v DEVICE
import sys
sys.path.insert(0, '.')
import bench
from multimodal_drl_rmc_b200.parallel import AgentEnsemble
for name, wl in (("ensemble8_per256", dict(bench.WORKLOADS["per256"])), ("ensemble8_default32", dict(bench.WORKLOADS["default32"], size=100_000, cap=200_000))):
    members = [bench.build_gpu_agent(wl, 0, seed=50 + k)[0] for k in range(8)]
    ens = AgentEnsemble(members)
    def stepe():
        for m in members: m.step += 1
        ens.learn()
    print(name, "us/step", 1e3 * bench._time_steps(stepe, 300), file=sys.stderr)
    del ens, members
for B in (1024, 8192, 65536):
    a5, _ = bench.build_gpu_agent(dict(bench.WORKLOADS["per256"], B=B), 0, seed=12)
    def step5():
        a5.step += 1; a5.learn(); a5.update_target_network()
    print("B=%d fp32 ms/step" % B, bench._time_steps(step5, 20, 3), file=sys.stderr)
    del a5
PY
python profiles/tools/phase_timeline.py per256 65536 2>&1 | grep -v DEVICE | head -14
