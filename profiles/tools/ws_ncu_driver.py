"""Six exact fp32 learner steps at B = 65,536 (k_learner_step<2>, the batch-stationary row phase) for an ncu capture:
  ncu --set full --clock-control none --import-source on -k regex:k_learner_step -s 3 -c 1 -o gpurun_out/ws python profiles/tools/ws_ncu_driver.py"""
import sys
sys.path.insert(0, '.')
import bench, torch
a5, _ = bench.build_gpu_agent(dict(bench.WORKLOADS["per256"], B=65536), 0, seed=12)
for _ in range(6):
    a5.step += 1; a5.learn(); a5.update_target_network()
torch.cuda.synchronize()
