#!/usr/bin/env python
"""In-pipeline kernel timeline of one learner step from the kernel-span recorder (rmc_debug_spans): first-CTA start and
last-CTA end of every kernel of the step, relative to the step's first kernel, median over steps.  Works inside graph
launches and across the two streams of the tensor-core step.
usage: python profiles/tools/kernel_spans.py [B=65536] [precision=bf16] [reps=12]        (one GPU)
       torchrun ... profiles/tools/kernel_spans.py 65536 bf16 12 sharded                 (sharded step, rank 0 prints)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from multimodal_drl_rmc_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
sharded = len(sys.argv) > 4 and sys.argv[4] == "sharded"
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
sl = None
if sharded:
    import torch.distributed as dist
    from multimodal_drl_rmc_b200.parallel import ShardedLearner
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
agent, _ = bench.build_gpu_agent(dict(bench.WORKLOADS["per256"], B=B), local, 0 if sharded else 12)
agent.learn_precision = prec
if sharded:
    sl = ShardedLearner(agent, exchange="peer")


def step():
    agent.step += 1
    if sl is not None:
        sl.learn()
    else:
        agent.learn()
        agent.update_target_network()


lib = _lib.lib()
for _ in range(5):
    step()
torch.cuda.synchronize()
_lib.check(lib.rmc_debug_spans(local, 1))
names_buf = C.create_string_buffer(512)
acc = []
for _ in range(reps):
    if sharded:
        dist.barrier()
    step()
    buf = (C.c_uint64 * 128)()
    _lib.check(lib.rmc_debug_spans_read_sync(local, buf, names_buf, 512))
    acc.append(np.array(buf[:], dtype=np.float64).reshape(64, 2))
_lib.check(lib.rmc_debug_spans(local, 0))
names = names_buf.value.decode().split(",")
a = np.stack(acc[2:])
if rank == 0:
    used = [k for k in range(len(names)) if np.all(a[:, k, 1] > 0)]
    t0 = np.min(a[:, used, 0], axis=1)
    print("B=%d precision=%s %s: kernel spans of one step, us since the step's first kernel (median over %d steps)" % (B, prec, "sharded x%s" % os.environ.get("WORLD_SIZE") if sharded else "single GPU", a.shape[0]))
    print("%-14s %9s %9s %9s" % ("kernel", "start", "end", "duration"))
    rows = sorted(((np.median(a[:, k, 0] - t0) * 1e-3, np.median(a[:, k, 1] - t0) * 1e-3, names[k]) for k in used))
    for s, e, nm in rows:
        print("%-14s %9.1f %9.1f %9.1f" % (nm, s, e, e - s))
    print("step span %.1f us" % max(e for _, e, _ in rows))
if sharded:
    dist.destroy_process_group()
