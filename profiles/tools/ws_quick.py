"""Exact fp32 step time at B = 8,192 and 65,536 (k_learner_step<2>).   usage (GPU box): python profiles/tools/ws_quick.py"""
import sys
sys.path.insert(0, '.')
import bench
for B in (8192, 65536):
    a5, _ = bench.build_gpu_agent(dict(bench.WORKLOADS["per256"], B=B), 0, seed=12)
    def step5():
        a5.step += 1; a5.learn(); a5.update_target_network()
    print("B=%d fp32 ms/step" % B, bench._time_steps(step5, 30, 5))
    del a5
