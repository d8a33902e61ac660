import sys
sys.path.insert(0, '.')
import bench, torch
a5, _ = bench.build_gpu_agent(dict(bench.WORKLOADS["per256"], B=65536), 0, seed=12)
for _ in range(6):
    a5.step += 1; a5.learn(); a5.update_target_network()
torch.cuda.synchronize()
