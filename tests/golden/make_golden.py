"""Generate the committed golden fixtures by running the UNMODIFIED reference learner
(/root/reference/dqn, imported through oracle/refharness.py) on seeded synthetic
transitions.  Runs only in the build container (the reference does not travel).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz, *.json

What is recorded per case (see CASES): the configuration and seeds, the injected sampling
randomness, and for every learner step the reference's sampled tree nodes, IS weights,
|td|, per-sample Huber terms / loss; the gradients of the first step; strided samples and
sha256 digests of the final online / target weights and of the final sum tree.  The test
``tests/test_oracle_golden.py`` replays the same recipe through ``oracle/dqn_oracle.py``.
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refharness  # noqa: E402
from oracle.dqn_oracle import synthetic_transitions  # noqa: E402

CASES = {
    # name: algo, D, B, cap, n_fill, steps, soft, target_freq
    "per_d14": dict(algo="PerDuelingDoubleDQNAgent", D=14, B=64, cap=1000, fill=1300, steps=4, soft=True),
    "per_d8_small": dict(algo="PerDuelingDoubleDQNAgent", D=8, B=32, cap=37, fill=37, steps=3, soft=True),
    "dueling_d14": dict(algo="DuelingDoubleDQNAgent", D=14, B=32, cap=512, fill=700, steps=3, soft=True),
    "double_d8_hard": dict(algo="DoubleDQNAgent", D=8, B=32, cap=256, fill=200, steps=4, soft=False,
                           target_freq=2),
    "dqn_d14": dict(algo="DQNAgent", D=14, B=16, cap=128, fill=128, steps=2, soft=True),
    "per_hybrid": dict(algo="PerDuelingDoubleDQNAgent", D=284, B=16, cap=150, fill=150, steps=2, soft=True, activation="elu", body="hybrid"),
    "per_d14_elu": dict(algo="PerDuelingDoubleDQNAgent", D=14, B=64, cap=600, fill=600, steps=3, soft=True, activation="elu"),
}
WEIGHT_SEED = 0
DATA_SEED = 20251018
TARGET_NOISE_SEED = 7
SAMPLE_STRIDE = 8


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def flat_params(net) -> np.ndarray:
    return np.concatenate([v.detach().numpy().ravel() for v in net.state_dict().values()])


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


def perturb_target(agent) -> None:
    """target <- online + N(0, 0.01^2) so that the two nets differ."""
    g = torch.Generator().manual_seed(TARGET_NOISE_SEED)
    with torch.no_grad():
        for pt in agent.target_network.parameters():
            pt.add_(torch.randn(pt.shape, generator=g) * 0.01)


def run_case(name: str, c: dict, tmp: str) -> dict:
    torch.set_num_threads(1)
    torch.manual_seed(WEIGHT_SEED)
    agent = refharness.make_reference_agent(c["algo"], c["D"], c["B"], c["cap"], tmp,
                                            soft=c["soft"], target_freq=c.get("target_freq", 30000),
                                            activation=c.get("activation", "relu"), body=c.get("body", "macro"))
    perturb_target(agent)
    per = c["algo"].startswith("Per")
    out = {"init_online_sha": sha(flat_params(agent.online_network)),
           "init_target_sha": sha(flat_params(agent.target_network))}

    obs, act, rew, done, nxt = synthetic_transitions(c["fill"], c["D"], DATA_SEED)
    for i in range(c["fill"]):
        agent.store_transitions([obs[i]], [int(act[i])], [float(rew[i])], [bool(done[i])], [nxt[i]], None)

    rec = {}
    mem = agent.replay_memory_buffer
    orig_sample = mem.sample_transitions
    if per:
        def sample_spy(step):
            r = orig_sample(step)
            rec["is_w"], rec["nodes"] = np.asarray(r[0], np.float64), np.asarray(r[1], np.int64)
            return r
        mem.sample_transitions = sample_spy
        orig_wb = mem.update_batch_priorities

        def wb_spy(nodes, abs_td):
            rec["abs_td"] = np.array(abs_td, copy=True)
            return orig_wb(nodes, abs_td)
        mem.update_batch_priorities = wb_spy
    agent.online_network.loss.register_forward_hook(
        lambda _m, _i, o: rec.__setitem__("huber", o.detach().numpy().copy()))

    steps = []
    for s in range(c["steps"]):
        agent.step = 1000 * s + 17  # exercises beta interpolation / hard-update modulus
        rec.clear()
        if per:
            np.random.seed(1000 + s)
            u = np.random.random_sample(c["B"])
            np.random.seed(1000 + s)
            inj = u
        else:
            random.seed(1000 + s)
            idx = random.sample(range(len(mem.replay_buffer)), c["B"])
            random.seed(1000 + s)
            inj = np.asarray(idx, np.int64)
        agent.learn()
        agent.update_target_network()
        st = {"inject": inj, "huber": rec["huber"].reshape(-1)}
        if per:
            st.update(nodes=rec["nodes"], is_w=rec["is_w"], abs_td=rec["abs_td"].reshape(-1),
                      tree_sha=sha(mem.replay_buffer.tree),
                      tree_stats=np.array([mem.replay_buffer.total_priority, mem.replay_buffer.max_priority,
                                           mem.replay_buffer.min_priority, mem.replay_buffer.size,
                                           mem.replay_buffer.data_pointer], np.float64))
        if s == 0:
            grads = np.concatenate([p.grad.detach().numpy().ravel() for p in agent.online_network.parameters()])
            st["grads_sample"] = grads[::SAMPLE_STRIDE].copy()
            st["grads_sha"] = sha(grads)
        steps.append(st)

    fo, ft = flat_params(agent.online_network), flat_params(agent.target_network)
    out.update(final_online_sample=fo[::SAMPLE_STRIDE].copy(), final_online_sha=sha(fo),
               final_target_sample=ft[::SAMPLE_STRIDE].copy(), final_target_sha=sha(ft))
    if per:
        out["final_tree"] = mem.replay_buffer.tree.copy() if c["cap"] <= 1000 else None
    probe = np.random.default_rng(5).random((64, c["D"]), dtype=np.float32)
    out["probe_actions"] = np.asarray(agent.online_network.actions(probe), np.int64)
    out["steps"] = steps
    return out


def sumtree_case(cap: int, n_ops: int, seed: int) -> dict:
    """Random push/assign/descend stream on the reference SumTree (dqn/utils/sum_tree.py)."""
    dqn = refharness.import_reference()
    from dqn.utils import SumTree
    rng = np.random.default_rng(seed)
    t = SumTree(cap)
    ops = []
    for _ in range(n_ops):
        kind = rng.integers(0, 3) if t.size > 0 else 0
        if kind == 0:
            p = float(np.float32(min(abs(rng.normal()) + 1e-4, 1.0)) ** np.float32(0.6))
            p = float(np.float32(p))
            t.add(p, ("row", int(t.data_pointer)))
            ops.append((0, 0, p, -1))
        elif kind == 1:
            leaf = int(rng.integers(0, t.size)) + cap - 1
            p = float(np.float32(np.float32(min(abs(rng.normal()) + 1e-4, 1.0)) ** np.float32(0.6)))
            t.update(leaf, p)
            ops.append((1, leaf, p, -1))
        else:
            v = float(rng.random() * t.total_priority)
            leaf, _p, _d = t.get_leaf(v)
            ops.append((2, 0, v, int(leaf)))
    return dict(cap=cap, ops=np.asarray(ops, np.float64), tree=t.tree.copy(),
                stats=np.array([t.total_priority, t.max_priority, t.min_priority, t.size, t.data_pointer]))


def main() -> None:
    if not refharness.reference_available():
        raise SystemExit("reference tree missing; goldens can only be generated in the build container")
    tmp = tempfile.mkdtemp(prefix="rmc_golden_")
    meta = {"cpu": cpu_fingerprint(), "weight_seed": WEIGHT_SEED, "data_seed": DATA_SEED,
            "target_noise_seed": TARGET_NOISE_SEED, "sample_stride": SAMPLE_STRIDE, "cases": CASES}
    try:
        only = [a for a in sys.argv[1:] if a in CASES]       # `make_golden.py per_d14_elu`: add one learner case, keep the rest
        for name, c in CASES.items():
            if only and name not in only:
                continue
            res = run_case(name, c, tmp)
            flat = {}
            for k, v in res.items():
                if k == "steps":
                    for i, st in enumerate(v):
                        for kk, vv in st.items():
                            flat["step%d_%s" % (i, kk)] = np.asarray(vv)
                elif v is not None:
                    flat[k] = np.asarray(v)
            np.savez_compressed(os.path.join(HERE, "learner_%s.npz" % name), **flat)
            print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(flat.items())[:4]})
        if only:
            old = json.load(open(os.path.join(HERE, "golden_meta.json")))
            old["cases"] = CASES
            with open(os.path.join(HERE, "golden_meta.json"), "w") as fh:
                json.dump(old, fh, indent=1, sort_keys=True)
            return
        for cap, n_ops, seed in [(1, 20, 1), (2, 60, 2), (3, 80, 3), (7, 200, 4), (69, 1500, 5), (64, 800, 6),
                                 (1000, 6000, 7)]:
            r = sumtree_case(cap, n_ops, seed)
            np.savez_compressed(os.path.join(HERE, "sumtree_cap%d.npz" % cap), **r)
            print("wrote sumtree cap", cap)
        # trained macro checkpoint (data fixture, 152 KB) + the reference's greedy actions on it
        pack_src = os.path.join(refharness.REFERENCE_ROOT, "env/custom_env/macro with lane",
                                "DuelingDoubleDQNAgent_lr0.0001_model_2e6_1e6.pack")
        pack_dst = os.path.join(HERE, "macro_with_lane.pack")
        shutil.copyfile(pack_src, pack_dst)
        dqn = refharness.import_reference()
        net = dqn.Networks.DuelingDeepQNetwork(torch.device("cpu"), 1e-4, refharness.macro_network_config,
                                               refharness.ObsBox(14), 8)
        step, episodes, rew_mean, len_mean = net.load(pack_dst)
        states = np.random.default_rng(11).random((512, 14), dtype=np.float32)
        with torch.no_grad():
            x = torch.as_tensor(states)
            np.savez_compressed(os.path.join(HERE, "act_macro_with_lane.npz"), states=states,
                                actions=np.asarray(net.actions(states), np.int64),
                                adv=net.advantages(x).numpy(), q=net(x).numpy(),
                                meta=np.array([step, episodes, rew_mean, len_mean], np.float64))
        resave = os.path.join(tmp, "resave", "m.pack")
        net.save(resave, step, episodes, rew_mean, len_mean)
        meta["pack_resave_identical"] = open(resave, "rb").read() == open(pack_dst, "rb").read()
        meta["pack_sha"] = hashlib.sha256(open(pack_dst, "rb").read()).hexdigest()
        with open(os.path.join(HERE, "golden_meta.json"), "w") as fh:
            json.dump(meta, fh, indent=1, sort_keys=True)
        print("meta", meta["cpu"], "pack resave identical:", meta["pack_resave_identical"])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
