#!/usr/bin/env python
"""Stage clocks of the tcgen05 forward kernel (CTA 0, pipeline 0, first 4 tiles): where a tile's time goes.
Run with RMC_TC_FWD_DBG=1."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["RMC_TC_FWD_DBG"] = "1"
import bench  # noqa: E402
from multimodal_drl_rmc_b200 import _lib  # noqa: E402

agent, _ = bench.build_gpu_agent(dict(bench.WORKLOADS["per256"], size=4096, cap=4096), 0, 0)
lib = _lib.lib()
agent.learn()                                   # sets last_grid (the debug reader's CTA count)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
states = torch.as_tensor(np.random.default_rng(0).random((n, 14), dtype=np.float32), device=agent.device)
acts = torch.empty(n, dtype=torch.int64, device=agent.device)
for _ in range(3):
    _lib.check(lib.rmc_learner_act_tc(agent._lh.handle, states.data_ptr(), n, acts.data_ptr(), _lib.stream_ptr()))
torch.cuda.synchronize()
buf = (C.c_uint64 * 128)()          # 4 CTA records of 32 slots (kDbgSlots); slots 32..34: kernel entry / weights landed / exit
got = C.c_int32(0)
_lib.check(lib.rmc_learner_debug_read_sync(agent._lh.handle, buf, 4, C.byref(got), _lib.stream_ptr()))
v = np.array(buf[:32], dtype=np.int64).reshape(4, 8)
names = ["X packed", "MMA1 done", "epi1 done", "MMA2 done", "epi2 done", "MMA3 done", "epi3 done", "tile done"]
for t in range(4):
    if v[t, 0] == 0:
        continue
    d = np.diff(v[t])
    print("tile %d: " % t + "  ".join("%s +%d" % (names[k + 1], d[k]) for k in range(7)) + "   total %d cycles" % (v[t, 7] - v[t, 0]),
          ("| gap to next tile start %d" % (v[t + 1, 0] - v[t, 7])) if t < 3 and v[t + 1, 0] else "")

k = np.array(buf[32:35], dtype=np.int64)
print("CTA 0: entry -> weights landed %d cycles, -> first tile's X packed %d, entry -> exit %d cycles" % (k[1] - k[0], v[0, 0] - k[0], k[2] - k[0]))
