"""Synthetic inputs of the learner path -- product-side generators used by ``bench.py`` and the
examples.  No SUMO, no network: transitions of the 1ramp_1x3 macro-state shape, and a stand-in for the reference's
vectorised environment (``dqn/utils/baselines_wrappers/subproc_vec_env.py``) so the trainer loop around the hot path can be
exercised and timed without the simulator.

(The test oracle keeps its own copy of the transition generator, ``oracle/dqn_oracle.py::synthetic_transitions``;
``tests/test_host_logic_cpu.py`` asserts both produce identical arrays.)
"""
from __future__ import annotations

import threading
import time

import numpy as np


def synthetic_transitions(n: int, obs_dim: int, seed: int = 20251018, n_actions: int = 8):
    """States U[0,1) float32 chained s'_t = s_{t+1} with the last feature on the 8-level action grid
    (``env/custom_env/rl_controller.py:299-319`` clips every feature to [0,1]; ``norm_last_action``), actions U{0..A-1},
    rewards clip(N(0.3, 1.5^2), -24, 3) (``rl_controller.py:391-425``), done on every 90th transition (3600 s / 40 s)."""
    rng = np.random.default_rng(seed)
    s = rng.random((n + 1, obs_dim), dtype=np.float32)
    s[:, -1] = (rng.integers(1, n_actions + 1, size=n + 1) / n_actions).astype(np.float32)
    a = rng.integers(0, n_actions, size=n).astype(np.int64)
    r = np.clip(rng.normal(0.3, 1.5, size=n), -24.0, 3.0).astype(np.float32)
    d = np.zeros(n, dtype=np.float32)
    d[89::90] = 1.0
    return s[:-1].copy(), a, r, d, s[1:].copy()


def seeded_priorities(n: int, seed: int) -> np.ndarray:
    """Non-degenerate leaf priorities: p = min(|N(0,1)| + 1e-4, 1)^0.6 in float32, what
    ``update_batch_priorities`` (dqn/replay_memory.py:94-98) would have left after |td| ~ |N(0,1)|."""
    rng = np.random.default_rng(seed)
    return np.power(np.minimum(np.abs(rng.normal(size=n)).astype(np.float32) + np.float32(1e-4), np.float32(1.0)),
                    np.float32(0.6)).astype(np.float32)


class SyntheticVecEnv:
    """Stand-in for ``SubprocVecEnv`` (``subproc_vec_env.py:39-112``): ``n_env`` environments stepped by a worker thread,
    with the reference's asynchronous interface -- ``reset()``, ``step_async(actions)``, ``step_wait()`` and ``step(actions)``
    = both.  Each step costs ``step_seconds`` of wall time on the CPU side (the SUMO step it stands in for takes
    milliseconds) and emits macro-state observations, the reward model of ``synthetic_transitions`` and episodes of
    ``episode_len`` steps with Monitor-style ``infos`` (``{'r': return, 'l': length}`` on the terminal step,
    ``dqn/utils/baselines_wrappers/monitor.py``)."""

    def __init__(self, n_env: int, obs_dim: int = 14, n_actions: int = 8, step_seconds: float = 0.0, episode_len: int = 90, seed: int = 0):
        self.num_envs, self.obs_dim, self.n_actions = int(n_env), int(obs_dim), int(n_actions)
        self.step_seconds, self.episode_len = float(step_seconds), int(episode_len)
        self._rng = np.random.default_rng(seed)
        self._t = np.zeros(self.num_envs, np.int64)
        self._ret = np.zeros(self.num_envs, np.float64)
        self._obs = self._draw_obs(np.zeros(self.num_envs, np.int64))
        self._pending = None
        self._result = None
        self._thread = None

    def _draw_obs(self, last_actions):
        o = self._rng.random((self.num_envs, self.obs_dim), dtype=np.float32)
        o[:, -1] = ((np.asarray(last_actions) + 1) / self.n_actions).astype(np.float32)
        return o

    def reset(self):
        self._t[:] = 0
        self._ret[:] = 0.0
        self._obs = self._draw_obs(np.zeros(self.num_envs, np.int64))
        return self._obs.copy()

    def _work(self, actions):
        if self.step_seconds > 0:
            time.sleep(self.step_seconds)
        rew = np.clip(self._rng.normal(0.3, 1.5, size=self.num_envs), -24.0, 3.0).astype(np.float32)
        self._t += 1
        self._ret += rew
        done = self._t >= self.episode_len
        infos = [dict() for _ in range(self.num_envs)]
        new_obs = self._draw_obs(actions)
        for e in np.nonzero(done)[0]:
            infos[e] = {'r': float(self._ret[e]), 'l': int(self._t[e])}
            self._t[e] = 0
            self._ret[e] = 0.0
        self._obs = new_obs
        self._result = (new_obs.copy(), rew, done.copy(), infos)

    def step_async(self, actions):
        if self._thread is not None:
            raise RuntimeError("step_async called while a step is pending")      # subproc_vec_env.py:71-75 asserts the same
        self._thread = threading.Thread(target=self._work, args=(list(actions),), daemon=True)
        self._thread.start()

    def step_wait(self):
        if self._thread is None:
            raise RuntimeError("step_wait without step_async")
        self._thread.join()
        self._thread = None
        return self._result

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
