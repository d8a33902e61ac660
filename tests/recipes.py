"""Shared recipe for the learner parity cases: how a case from tests/golden/golden_meta.json
is replayed through the oracle (and, in the gpu tests, through the CUDA path)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import torch

from oracle.dqn_oracle import OracleLearner, synthetic_transitions

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_meta() -> dict:
    with open(os.path.join(GOLDEN_DIR, "golden_meta.json")) as fh:
        return json.load(fh)


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def flat_params(net) -> np.ndarray:
    return np.concatenate([v.detach().cpu().numpy().ravel() for v in net.state_dict().values()])


def cpu_fingerprint() -> str:
    model = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return "%s|torch%s|numpy%s" % (model, torch.__version__, np.__version__)


def perturb_target(target_net, seed: int) -> None:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for pt in target_net.parameters():
            pt.add_(torch.randn(pt.shape, generator=g) * 0.01)


def build_oracle_case(c: dict, meta: dict) -> OracleLearner:
    """Same construction as tests/golden/make_golden.py::run_case, but with the oracle port."""
    torch.set_num_threads(1)
    torch.manual_seed(meta["weight_seed"])
    lrn = OracleLearner(c["algo"], c["D"], 8, c["B"], c["cap"], soft=c["soft"],
                        target_freq=c.get("target_freq", 30000), activation=c.get("activation", "relu"),
                        body=c.get("body", "macro"))
    perturb_target(lrn.target, meta["target_noise_seed"])
    obs, act, rew, done, nxt = synthetic_transitions(c["fill"], c["D"], meta["data_seed"])
    for i in range(c["fill"]):
        lrn.store([obs[i]], [int(act[i])], [float(rew[i])], [bool(done[i])], [nxt[i]])
    return lrn


def step_number(s: int) -> int:
    return 1000 * s + 17


def max_rel(a, b) -> float:
    """Per-tensor max-norm relative error."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b))) / den
