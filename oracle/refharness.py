"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loader for the *unmodified* reference learner (/root/reference/dqn) inside the
build container.  The reference cannot travel to the GPU box, so this module is
used only by ``tests/golden/make_golden.py`` (to generate committed fixtures)
and by CPU tests that are skipped when ``/root/reference`` is absent.

The reference's ``dqn/__init__.py:1-4`` imports ``gymnasium`` (dqn/env_wrap.py:1)
and ``colorama`` (dqn/agent.py:12), neither of which is installed here; both are
irrelevant to the learner hot path, so tiny stand-ins are placed in
``sys.modules`` before the import.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RMC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dqn"))


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:  # minimal stand-in for gymnasium.Env
            def reset(self, seed=None, options=None):
                return None

        class Wrapper:
            def __init__(self, env):
                self.env = env

        spaces = types.ModuleType("gymnasium.spaces")

        class Box:
            def __init__(self, low=0.0, high=1.0, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        class Discrete:
            def __init__(self, n):
                self.n = n

        class Dict(dict):
            pass

        class Tuple(tuple):
            pass

        spaces.Box, spaces.Discrete, spaces.Dict, spaces.Tuple = Box, Discrete, Dict, Tuple
        core = types.ModuleType("gymnasium.core")
        core.Wrapper = Wrapper
        gym.Env, gym.Wrapper, gym.spaces, gym.core = Env, Wrapper, spaces, core
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
        sys.modules["gymnasium.core"] = core
    if "colorama" not in sys.modules:
        col = types.ModuleType("colorama")

        class _Blank:
            def __getattr__(self, _name):
                return ""

        col.Fore, col.Style, col.Back = _Blank(), _Blank(), _Blank()
        col.init = lambda *a, **k: None
        sys.modules["colorama"] = col


def import_reference():
    """Return the reference's ``dqn`` package (Agents, Networks, ...) unmodified."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import dqn  # noqa: WPS433  (the reference package)

    return dqn


class ObsBox:
    """What the reference's ``nn_conf_func`` receives as ``input_dim`` (a Box)."""

    def __init__(self, dim: int):
        self.shape = (int(dim),)


def macro_network_config(input_dim_space):
    """Restatement of the macro-state ``network_config``
    (reference: ``env/custom_env/macro with lane/dqn_config.py:58-104``):
    Linear(D,256)-ReLU-Linear(256,128)-ReLU body, fc_out_dim 128, Adam, SmoothL1."""
    import torch.nn as nn
    import torch.optim as optim

    d = input_dim_space.shape[0]
    body = nn.Sequential(nn.Linear(d, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU())
    return body, 128, optim.Adam, nn.SmoothL1Loss


def macro_network_config_elu(input_dim_space):
    """The macro body with the repo-HEAD activation (reference: ``env/dqn_config.py:175`` ``ACTIVATION = nn.ELU()``)."""
    import torch.nn as nn
    import torch.optim as optim

    d = input_dim_space.shape[0]
    body = nn.Sequential(nn.Linear(d, 256), nn.ELU(), nn.Linear(256, 128), nn.ELU())
    return body, 128, optim.Adam, nn.SmoothL1Loss


def hybrid_network_config(input_dim_space):
    """The reference's OWN repo-HEAD network: ``TwoStreamHybridNetwork`` is executed from the class definition in
    ``env/dqn_config.py:66-143`` (the module itself imports the SUMO bindings and cannot be imported here, so only that
    class is compiled, from the reference tree, at fixture-generation time) with the parameters of
    ``network_config`` (:148-193): macro 14, grid (2,27,5), CNN (32,(3,3),(1,1)) (64,(3,3),(2,1)) (64,(3,3),(2,2)),
    dense [512, 256], ELU, Adam, SmoothL1."""
    import ast
    import torch
    import torch.nn as nn
    import torch.optim as optim

    path = os.path.join(REFERENCE_ROOT, "env", "dqn_config.py")
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TwoStreamHybridNetwork"][0]
    ns = {"nn": nn, "T": torch}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), path, "exec"), ns)
    net = ns["TwoStreamHybridNetwork"](macro_vec_len=14, micro_shape_chw=(2, 27, 5),
                                       cnn_params=[(32, (3, 3), (1, 1)), (64, (3, 3), (2, 1)), (64, (3, 3), (2, 2))],
                                       dense_params=[512, 256], activation_fn=nn.ELU())
    return net, net.fc_out_dim, optim.Adam, nn.SmoothL1Loss


def make_reference_agent(algo: str, obs_dim: int, batch: int, cap: int, tmpdir: str, *,
                         lr=1e-4, gamma=0.99, tau=1e-3, soft=True, target_freq=30000,
                         eps_decay=2e6, n_env=1, n_actions=8, activation="relu", body="macro"):
    """Construct a reference agent exactly as train.py:24-48 would (macro MLP)."""
    dqn = import_reference()
    cls = getattr(dqn.Agents, algo)
    return cls(n_env=n_env, lr=lr, gamma=gamma, epsilon_start=1.0, epsilon_min=0.01,
               epsilon_decay=eps_decay, epsilon_exp_decay=True,
               nn_conf_func=hybrid_network_config if body == "hybrid" else (macro_network_config_elu if activation == "elu" else macro_network_config),
               input_dim=ObsBox(obs_dim), output_dim=n_actions, batch_size=batch,
               min_buffer_size=batch, buffer_size=cap, update_target_frequency=target_freq,
               target_soft_update=soft, target_soft_update_tau=tau, save_frequency=10000,
               log_frequency=4500, save_dir=tmpdir + "/", log_dir=tmpdir + "/", load=False,
               algo=algo, gpu="0")
