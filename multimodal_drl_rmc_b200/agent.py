"""Drop-in agents: same constructor, attributes and methods as ``dqn/agent.py`` (Agent :18-158 and
the four concrete agents :275-320); ``learn()`` / ``update_target_network()`` / ``choose_actions()``
run as fused CUDA launches of librmc_b200 on a GPU-resident replay.

Randomness of the minibatch draw (``Agent.sampling``):
  "device" (default)  counter-based Philox / keyed permutation on the GPU, no host work per step
  "host"              consume the host RNG streams exactly like the reference does
                      (``np.random.random_sample(B)`` for PER, ``random.sample(range(n), B)`` for
                      uniform replay) and upload them: with equal seeds the sampled transitions
                      are bit-identical to the reference's.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import os
import random
import time
from collections import deque
from datetime import timedelta

import numpy as np
import torch as T

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .network import DeepQNetwork, DuelingDeepQNetwork, LearnerHandle
from .replay_memory import ReplayMemoryNaive, ReplayMemoryPrioritized


@functools.lru_cache(maxsize=64)
def _logs(start, end):
    return float(np.log(start)), float(np.log(end))


def _interp2(x, x1, y0, y1):
    """np.interp(x, [0, x1], [y0, y1]) for a scalar x in python floats: the IEEE operations of numpy's compiled interp
    (end points returned as they are, otherwise slope * (x - 0) + y0) without the array round trip."""
    x = float(x)
    x1 = float(x1)
    if x <= 0.0:
        return y0
    if x >= x1:
        return y1
    slope = (y1 - y0) / (x1 - 0.0)
    r = slope * (x - 0.0) + y0
    if r != r:                         # numpy's NaN fix-up (infinite end points, e.g. log(0)): the other end, then the common value
        r = slope * (x - x1) + y1
        if r != r and y0 == y1:
            r = y0
    return r


def epsilon_value(x, start, end, decay, exp_decay):
    """Exploration rate after x environment steps (dqn/agent.py:86-90): exp of a linear interpolation of the logs, or the
    linear interpolation itself; bit-identical to the reference's np.exp(np.interp(...)) (tests/test_host_logic_cpu.py),
    several microseconds cheaper per call -- it is evaluated on every environment step."""
    if exp_decay:
        l0, l1 = _logs(start, end)
        return np.exp(_interp2(x, decay, l0, l1))
    return np.float64(_interp2(x, decay, float(start), float(end)))


class Agent:
    # flavour switches, fixed by the concrete classes below (dqn/agent.py:275-320)
    _PER = False
    _DUELING = False
    _DOUBLE = False

    def __init__(self, n_env, lr, gamma, epsilon_start, epsilon_min, epsilon_decay, epsilon_exp_decay, nn_conf_func,
                 input_dim, output_dim, batch_size, min_buffer_size, buffer_size, update_target_frequency,
                 target_soft_update, target_soft_update_tau, save_frequency, log_frequency, save_dir, log_dir, load,
                 algo, gpu):
        self.n_env = n_env
        self.lr = lr
        self.gamma = gamma
        self.epsilon_start = epsilon_start
        self.epsilon_min = epsilon_min
        self.epsilon_decay = epsilon_decay
        self.epsilon_exp_decay = epsilon_exp_decay
        self.nn_conf_func = nn_conf_func
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.batch_size = batch_size
        self.min_buffer_size = min_buffer_size
        self.buffer_size = buffer_size
        self.update_target_frequency = update_target_frequency
        self.target_soft_update = target_soft_update
        self.target_soft_update_tau = target_soft_update_tau
        self.save_frequency = save_frequency
        self.log_frequency = log_frequency
        self.load = load

        self.step = 0
        self.resume_step = 0
        self.episode_count = 0
        self.ep_info_buffer = deque([], maxlen=50)

        path = algo + '_lr' + str(lr)
        self.save_path = save_dir + path + '_' + 'model.pack'
        self.summary_writer = self._make_writer(log_dir + path + '/')

        if not T.cuda.is_available():
            raise RuntimeError("multimodal_drl_rmc_b200 agents need a CUDA (sm_100a) device: there is no CPU fallback")
        self.device = T.device("cuda:" + str(gpu))
        print("DEVICE", "=", self.device, T.cuda.get_device_name(self.device))
        self.start_time = time.time()

        # ---- wiring of dqn/agent.py:275-320 -------------------------------------------------
        if self._PER:
            self.replay_memory_buffer = ReplayMemoryPrioritized(self.buffer_size, self.batch_size, self.epsilon_decay)
        else:
            self.replay_memory_buffer = ReplayMemoryNaive(self.buffer_size, self.batch_size)
        self.replay_memory_buffer._ring.device_index = self.device.index
        net_cls = DuelingDeepQNetwork if self._DUELING else DeepQNetwork
        reduction = 'none' if self._PER else 'mean'
        self.online_network = net_cls(self.device, self.lr, self.nn_conf_func, self.input_dim, self.output_dim, reduction=reduction)
        self.target_network = net_cls(self.device, self.lr, self.nn_conf_func, self.input_dim, self.output_dim, reduction=reduction)

        hyper = _lib.Hyper(float(lr), 0.9, 0.999, 1e-8, float(gamma), float(target_soft_update_tau) * n_env,
                           1e-4, 0.6, 1.0)
        self._lh = LearnerHandle(self.online_network._obs_dim, output_dim, self._DUELING, self._DOUBLE, self._PER,
                                 max(int(batch_size), 1), self.device.index, hyper,
                                 activation=self.online_network._activation, hybrid=self.online_network._hybrid)
        self._pending_step = None        # a learn() that has been recorded but not launched yet (see learn())
        self.online_network._bind(self._lh, _lib.ONLINE)
        self.target_network._bind(self._lh, _lib.TARGET)
        self.update_target_network(force=True)

        self.sampling = "device"
        self._learn_calls = 0
        self._adam_t = 0
        self._B = int(batch_size)
        self._dev_index = self.device.index
        self._args = _lib.StepArgs()
        self._args.batch = self._B
        self._args_ref = C.byref(self._args)
        self._step_fn = lib().rmc_learner_step
        self._step_push_fn = lib().rmc_learner_step_push
        self.replay_memory_buffer._ring.defer_small_pushes = True     # per-step rows ride with the next learn()
        # learn() is lazy (see learn / update_target_network): whoever looks at the learner or the replay first makes the
        # recorded step happen
        self._pending_step = None
        self._lh._flush_hook = self._flush_step
        self.replay_memory_buffer._ring._owner_flush = self._flush_step
        self._learn_phases = _lib.PH_LEARN if self._PER else (_lib.PH_LEARN & ~_lib.PH_PRIORITY)
        self.sampling_seed = 0x5EED

    @property
    def learn_precision(self):
        """"fp32" (default): the exact FFMA path, parity with the reference at 1e-5.  "bf16": the tcgen05/TMEM
        tensor-core path for dense batches (BASELINE configs[4]); stated looser bound, see csrc/rmc_tc_train.cuh."""
        return "bf16" if self._args.precision == _lib.PREC_BF16_TC else "fp32"

    @learn_precision.setter
    def learn_precision(self, v):
        if v not in ("fp32", "bf16"):
            raise ValueError("learn_precision must be 'fp32' or 'bf16'")
        self._args.precision = _lib.PREC_BF16_TC if v == "bf16" else _lib.PREC_FP32

    @property
    def sampling_seed(self):
        return self._args.seed

    @sampling_seed.setter
    def sampling_seed(self, v):
        self._args.seed = int(v)

    @staticmethod
    def _make_writer(path):
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(path)

    # ------------------------------------------------------------------ replay ---------------
    def store_transitions(self, obses, actions, rews, dones, new_obses, infos):
        for i in self.replay_memory_buffer.store_transitions(obses, actions, rews, dones, new_obses):
            if infos:
                self.ep_info_buffer.append({'r': infos[i]['r'], 'l': infos[i]['l']})
                self.episode_count += 1

    # ------------------------------------------------------------------ acting ---------------
    def epsilon(self):
        """dqn/agent.py:86-90, same floating-point operations (see epsilon_value)."""
        return epsilon_value(self.step * self.n_env, self.epsilon_start, self.epsilon_min, self.epsilon_decay, self.epsilon_exp_decay)

    def choose_actions(self, obses):
        """dqn/agent.py:92-99.  ``self.exploration = "host"`` (default) consumes Python's ``random`` exactly like the reference
        (identical stream); ``"device"`` does greedy act + epsilon-greedy in ONE library call with a Philox stream keyed by
        (sampling_seed, step) -- same distribution, for vectorised envs."""
        if getattr(self, "exploration", "host") == "device":
            net = self.online_network
            net._push()
            x = np.ascontiguousarray(np.asarray(obses, dtype=np.float32)).reshape(-1, net._obs_dim)
            out = np.empty(x.shape[0], np.int64)
            self._act_calls = getattr(self, "_act_calls", 0) + 1
            check(lib().rmc_learner_act_eps_host_sync(self._lh.handle, x.ctypes.data, x.shape[0], out.ctypes.data, float(self.epsilon()),
                                                      int(self.sampling_seed) ^ 0xAC7, self._act_calls, stream_ptr(self._dev_index)))
            return out.tolist()
        actions = self.online_network.actions(obses)
        eps = self.epsilon()          # the reference re-evaluates it per env; `step` does not move inside the call, so once is the same value
        for i in range(len(actions)):
            if random.random() <= eps:
                actions[i] = random.randint(0, self.output_dim - 1)
        return actions

    # ------------------------------------------------------------------ learning -------------
    def _step_args(self, phases, u=None, indices=None):
        a = self._args                      # one StepArgs per agent, refreshed in place (no per-step allocation)
        a.phases = int(phases)
        a.counter = self._learn_calls
        a.adam_t = self._adam_t
        a.u_dev = None
        a.idx_dev = None
        keep = None
        if phases & _lib.PH_SAMPLE:
            if self._PER:
                a.per_beta = self._beta(self.step * self.n_env)
                if u is None and self.sampling == "host":
                    u = np.random.random_sample(self._B)
                if u is not None:
                    keep = T.as_tensor(np.asarray(u, np.float64), device=self.device)
                    a.u_dev = keep.data_ptr()
            else:
                if indices is None and self.sampling == "host":
                    indices = random.sample(range(len(self.replay_memory_buffer.replay_buffer)), self._B)
                if indices is not None:
                    keep = T.as_tensor(np.asarray(indices, np.int64), device=self.device)
                    a.idx_dev = keep.data_ptr()
        return a, keep

    def _beta(self, x):
        """np.interp(x, [0, beta_inc], [0.4, 1.0]) (dqn/replay_memory.py:74) in plain python floats: same
        IEEE operations as numpy's compiled interp (slope * (x - x0) + y0), without the array round trip."""
        mem = self.replay_memory_buffer
        x = float(x)
        x1 = float(mem.beta_inc)
        if x <= 0.0:
            return mem.beta_start
        if x >= x1:
            return mem.beta_end
        return ((mem.beta_end - mem.beta_start) / (x1 - 0.0)) * (x - 0.0) + mem.beta_start

    def learn(self, u=None, indices=None, fuse_target_update=False):
        """dqn/agent.py:166-185 / 204-226 / 245-272.  ``u`` / ``indices`` inject the sampling randomness (tests).

        The reference's trainer calls ``learn(); update_target_network()`` back to back (train.py:99-101) and the fused
        kernel can do both in one launch (Polyak / hard sync on the post-Adam weights, same bits as two launches --
        tests/test_gpu_learner.py::test_fused_target_update_equals_separate_call).  So ``learn()`` is LAZY: it fixes
        everything that defines the step (batch, beta, sampling randomness, Adam step count, the env rows held back by
        ``store_transitions``) and records it; the ``update_target_network()`` that follows launches it together with the
        target update.  Anything else that looks at the learner or the replay first (``last_loss``, ``state_dict``,
        ``choose_actions``, ``store_transitions``, replay statistics, another ``learn``) launches the recorded step as it
        is, so the laziness is not observable.  ``fuse_target_update=True`` launches immediately with this step's target
        update inside (the following ``update_target_network()`` call is then skipped once)."""
        self._flush_step()
        ring = self.replay_memory_buffer._ring
        if ring._handle is None:
            raise RuntimeError("replay memory is empty (no transition stored yet)")
        if not self._PER and indices is None and self._B > ring.count:
            # random.sample raises ValueError here in the reference (dqn/replay_memory.py:39); raised now, not when the recorded step is launched
            raise _lib.RmcError("librmc_b200 error -3: rmc_learner_step: sample larger than population")
        self._learn_calls += 1
        self._adam_t += 1
        phases = self._learn_phases
        a, keep = self._step_args(phases, u, indices)
        self._keepalive = keep
        if fuse_target_update:
            self._launch_step(phases | self._target_phase())
            self._target_fused_for = self._learn_calls
        else:
            self._pending_step = phases

    def _launch_step(self, phases):
        """One host call: this env step's rows (held back by store_transitions) + the learner step [+ target update]."""
        a = self._args
        a.phases = int(phases)
        ring = self.replay_memory_buffer._ring
        n_new = ring.take_pending()
        if n_new:
            rc = self._step_push_fn(self._lh.handle, ring._handle, self._args_ref, *ring._small_ptrs, n_new, stream_ptr(self._dev_index))
        else:
            rc = self._step_fn(self._lh.handle, ring._handle, self._args_ref, stream_ptr(self._dev_index))
        if rc:
            check(rc)
        ver = self._lh.version
        ver[0] += 1
        if phases & (_lib.PH_POLYAK | _lib.PH_HARDSYNC):
            ver[1] += 1

    def _flush_step(self):
        """Launch a recorded learn() that no update_target_network() has picked up."""
        phases = self._pending_step
        if phases is not None:
            self._pending_step = None
            self._launch_step(phases)

    def _target_phase(self, force=False):
        if (not self.target_soft_update and self.step % (self.update_target_frequency // self.n_env) == 0) or force:
            return _lib.PH_HARDSYNC
        if self.target_soft_update:
            return _lib.PH_POLYAK
        return 0

    def update_target_network(self, force=False):
        """dqn/agent.py:101-110.  Directly after ``learn()`` (train.py:99-101) this is the call that launches the step,
        with the target update fused into it."""
        pend = self._pending_step
        if pend is not None and not force:
            self._pending_step = None
            self._launch_step(pend | self._target_phase())
            return
        self._flush_step()
        if not force and getattr(self, "_target_fused_for", None) == self._learn_calls:
            self._target_fused_for = None
            return
        phase = self._target_phase(force)
        if not phase:
            return
        self.online_network._push()
        self.target_network._push()
        a = _lib.StepArgs()
        a.batch, a.phases, a.adam_t = 1, int(phase), 1
        ring = self.replay_memory_buffer._ring
        rh = ring.handle if ring.handle is not None else self._dummy_ring().handle
        check(lib().rmc_learner_step(self._lh.handle, rh, C.byref(a), stream_ptr(self.device.index)))
        self._lh.version[_lib.TARGET] += 1

    def _dummy_ring(self):
        # target sync before the first transition arrives (constructor): a 1-slot ring of the right kind
        if getattr(self, "_dummy", None) is None:
            from .replay_memory import DeviceRing
            self._dummy = DeviceRing(1, prioritized=self._PER, device_index=self.device.index).ensure(self.online_network._obs_dim)
        return self._dummy

    def last_loss(self):
        """Loss of the last learn() (device -> host read; synchronises)."""
        self._flush_step()
        out = C.c_float()
        check(lib().rmc_learner_loss_sync(self._lh.handle, C.byref(out), stream_ptr(self._dev_index)))
        return out.value

    # ------------------------------------------------------------------ checkpoints / logs ---
    def load_model(self):
        if self.load and os.path.exists(self.save_path):
            print()
            print("Resume training from " + self.save_path + "...")
            self.resume_step, self.episode_count, rew_mean, len_mean = self.online_network.load(self.save_path)
            [self.ep_info_buffer.append({'r': rew_mean, 'l': len_mean}) for _ in range(np.min([self.episode_count, self.ep_info_buffer.maxlen]))]
            print("Step: ", self.resume_step * self.n_env, ", Episodes: ", self.episode_count, ", Avg Rew: ", rew_mean, ", Avg Ep Len: ", len_mean)
            self.update_target_network(force=True)
            self.step = self.resume_step

    def save_model(self):
        if self.step % self.save_frequency == 0 and self.step > self.resume_step:
            print()
            print("Saving model...")
            self.online_network.save(self.save_path, self.step, self.episode_count, self.info_mean('r'), self.info_mean('l'))
            print("OK!")

    def log(self):
        if self.step % self.log_frequency == 0 and self.step > self.resume_step:
            rew_mean, len_mean = self.info_mean('r'), self.info_mean('l')
            print()
            print('Step: ', self.step * self.n_env, ' (' + str(self.step) + 'x' + str(self.n_env) + ')')
            print('Avg Rew: ', rew_mean)
            print('Avg Ep Len: ', len_mean)
            print('Episodes: ', self.episode_count)
            print('---', str(timedelta(seconds=round((time.time() - self.start_time), 0))), '---')
            self.summary_writer.add_scalar('AvgRew', rew_mean, global_step=(self.step * self.n_env))
            self.summary_writer.add_scalar('AvgEpLen', len_mean, global_step=(self.step * self.n_env))
            self.summary_writer.add_scalar('Episodes', self.episode_count, global_step=(self.step * self.n_env))

    # ------------------------------------------------------------------ learner diagnostics ------
    def diagnostics(self):
        """Scalars of the last learn() step, reduced on the device and read back in one go (call it at log
        frequency, it synchronises): loss, mean / max |td|, mean Q(s,a), mean target, PER beta and tree extremes.
        The reference logs none of these (dqn/agent.py:130-143)."""
        lh = self._lh
        B = self._B
        td, q, y = lh.output("abs_td")[:B], lh.output("q_sa")[:B], lh.output("y")[:B]
        vals = T.stack([lh.output("loss")[0], td.mean(), td.max(), q.mean(), y.mean()]).tolist()
        out = {"loss": vals[0], "abs_td_mean": vals[1], "abs_td_max": vals[2], "q_mean": vals[3], "target_mean": vals[4]}
        if self._PER:
            st = self.replay_memory_buffer._ring.stats()
            out.update(beta=self._beta(self.step * self.n_env), total_priority=st.total_priority,
                       max_priority=st.max_priority, min_priority=st.min_priority, replay_size=int(st.size))
        return out

    def log_diagnostics(self):
        """Write diagnostics() to the agent's SummaryWriter under Learner/* (same global_step as Agent.log)."""
        for k, v in self.diagnostics().items():
            self.summary_writer.add_scalar("Learner/" + k, v, global_step=(self.step * self.n_env))

    # ------------------------------------------------------------------ exact-resume side-car ---
    @staticmethod
    def _sidecar_path(path):
        path = os.fspath(path)
        return path if path.endswith(".npz") else path + ".npz"      # np.savez appends the suffix; load must look for the same file

    def save_learner_state(self, path, replay=True, host_rng=True):
        """Side-car next to the (unchanged) ``.pack`` checkpoint.  The reference's resume keeps only the online weights
        and four counters (dqn/agent.py:112-121, dqn/network.py:27-47): Adam restarts from zero moments, the target net
        is re-copied and the replay is refilled with 100 k fresh env steps (train.py:63-81).  This file holds what is
        needed to continue as if the process had never stopped:

          always      online / target weights, Adam moments and step count, learner counters, sampling seed, episode statistics
          replay      the replay ring rows, cursor and size and -- for PER -- the leaf priorities (inner tree nodes and
                      max / min are rebuilt exactly on load); 132 MB for the 1M x D=14 default, read back in one D2H pass
          host_rng    the states of python's ``random`` and of ``np.random`` (the host exploration / sampling streams)

        A run resumed with ``load_learner_state`` produces bit-identical sampled indices, losses and weights
        (tests/test_gpu_edges.py::test_sidecar_resume_is_bit_identical)."""
        lh = self._lh
        ring = self.replay_memory_buffer._ring
        out = dict(format=np.int64(2), n_params=np.int64(lh.n_params), obs_dim=np.int64(self.online_network._obs_dim),
                   n_actions=np.int64(self.output_dim), flavour=np.array([self._PER, self._DUELING, self._DOUBLE], np.int64),
                   online=lh.get_params(_lib.ONLINE).cpu().numpy(), target=lh.get_params(_lib.TARGET).cpu().numpy(),
                   adam_m=lh.get_params(_lib.ADAM_M).cpu().numpy(), adam_v=lh.get_params(_lib.ADAM_V).cpu().numpy(),
                   adam_t=self._adam_t, learn_calls=self._learn_calls, step=self.step, seed=self.sampling_seed,
                   episode_count=self.episode_count,
                   ep_info=np.array([[e['r'], e['l']] for e in self.ep_info_buffer], np.float64).reshape(-1, 2))
        if replay and ring.handle is not None:
            st = ring.stats()
            size = int(st.size)
            out.update(replay_capacity=np.int64(st.capacity), replay_size=np.int64(size), replay_dp=np.int64(st.data_pointer),
                       replay_rows=ring.read_rows(0, size) if size else np.zeros((0, ring.row_floats), np.float32))
            if self._PER:
                leaves = np.empty(size, np.float64)
                if size:
                    check(lib().rmc_replay_read_tree_sync(ring.handle, leaves.ctypes.data, int(st.capacity) - 1, size, stream_ptr(self._dev_index)))
                pri = leaves.astype(np.float32)
                assert np.array_equal(pri.astype(np.float64), leaves), "leaf priorities are float32-exact by construction"
                out["replay_leaves"] = pri
        if host_rng:
            import pickle
            out["host_rng"] = np.frombuffer(pickle.dumps((random.getstate(), np.random.get_state())), np.uint8)
        np.savez(self._sidecar_path(path), **out)

    def load_learner_state(self, path):
        z = np.load(self._sidecar_path(path), allow_pickle=False)
        lh = self._lh
        if "n_params" in z.files:       # format 2: validate before touching the learner
            flavour = [int(v) for v in z["flavour"]]
            if (int(z["n_params"]) != lh.n_params or int(z["obs_dim"]) != self.online_network._obs_dim or int(z["n_actions"]) != self.output_dim
                    or flavour != [int(self._PER), int(self._DUELING), int(self._DOUBLE)]):
                raise ValueError("side-car %s was written by a different agent (params %d, obs_dim %d, actions %d, PER/dueling/double %s)"
                                 % (path, int(z["n_params"]), int(z["obs_dim"]), int(z["n_actions"]), flavour))
        self._flush_step()
        for kind, key in ((_lib.ONLINE, "online"), (_lib.TARGET, "target"), (_lib.ADAM_M, "adam_m"), (_lib.ADAM_V, "adam_v")):
            lh.set_params(kind, T.as_tensor(z[key]))
        lh.version[_lib.ONLINE] += 1
        lh.version[_lib.TARGET] += 1
        self.online_network._module_dirty = self.target_network._module_dirty = False
        self._adam_t, self._learn_calls, self.step = int(z["adam_t"]), int(z["learn_calls"]), int(z["step"])
        self.sampling_seed = int(z["seed"])
        if "episode_count" in z.files:
            self.episode_count = int(z["episode_count"])
            self.ep_info_buffer.clear()
            for r, l in z["ep_info"]:
                self.ep_info_buffer.append({'r': float(r), 'l': float(l)})
        if "replay_rows" in z.files:
            ring = self.replay_memory_buffer._ring
            if int(z["replay_capacity"]) != ring.capacity:
                raise ValueError("side-car replay capacity %d differs from this agent's %d" % (int(z["replay_capacity"]), ring.capacity))
            rows = np.ascontiguousarray(z["replay_rows"], np.float32)
            if ring._handle is not None:      # start from an empty ring of the same shape
                lib().rmc_replay_destroy(ring._handle)
                ring._handle, ring._pending = None, 0
            ring.ensure(self.online_network._obs_dim)
            if rows.shape[1] != ring.row_floats:
                raise ValueError("side-car replay rows have %d floats, this agent's have %d" % (rows.shape[1], ring.row_floats))
            pri = np.ascontiguousarray(z["replay_leaves"], np.float32) if "replay_leaves" in z.files else None
            check(lib().rmc_replay_load_host(ring._handle, rows.ctypes.data, None if pri is None else pri.ctypes.data, int(z["replay_size"]),
                                             int(z["replay_dp"]), stream_ptr(self._dev_index)))
            ring.count = int(z["replay_size"])
        if "host_rng" in z.files:
            import pickle
            py_state, np_state = pickle.loads(z["host_rng"].tobytes())
            random.setstate(py_state)
            np.random.set_state(np_state)

    def info_mean(self, i):
        i_mean = np.mean([e[i] for e in self.ep_info_buffer]) if len(self.ep_info_buffer) else float('nan')
        return i_mean if not math.isnan(i_mean) else 0.


class SimpleAgent(Agent):
    """dqn/agent.py:148-185: y = r + (1-d) gamma max_a Q_target(s')."""
    _DOUBLE = False


class DoubleAgent(Agent):
    """dqn/agent.py:188-226: a* = argmax_a Q_online(s'), y uses Q_target(s')[a*]."""
    _DOUBLE = True


class PerDoubleAgent(Agent):
    """dqn/agent.py:229-272: DoubleAgent + IS weights + |td| write-back before backward."""
    _DOUBLE = True
    _PER = True


class DQNAgent(SimpleAgent):
    pass


class DoubleDQNAgent(DoubleAgent):
    pass


class DuelingDoubleDQNAgent(DoubleAgent):
    _DUELING = True


class PerDuelingDoubleDQNAgent(PerDoubleAgent):
    _DUELING = True
