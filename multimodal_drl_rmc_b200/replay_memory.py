"""GPU-resident replay memories with the reference's class/method surface.

Mirrors ``dqn/replay_memory.py`` (ReplayMemoryNaive :24-39, ReplayMemoryPrioritized :43-98) and
``dqn/utils/sum_tree.py`` (SumTree).  Storage is a ring buffer (+ float64 sum tree) in HBM owned
by librmc_b200; the Python objects are thin views.  ``Agent.learn()`` never goes through the
list-returning ``sample_transitions`` -- that method exists for external callers and tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class DeviceRing:
    """Owner of one ``rmc_replay_t`` (created lazily: the reference constructors do not know the
    observation size, so the ring is allocated when the first transition arrives)."""

    def __init__(self, capacity: int, prioritized: bool, device_index=None):
        self.capacity = int(capacity)
        self.prioritized = bool(prioritized)
        self.device_index = device_index
        self._handle = None
        self._pending = 0           # rows of the current env step held back for the fused store + learn call
        self._owner_flush = None    # set by the owning Agent: launches a recorded (lazy) learn() before the replay is touched
        self.count = 0              # host mirror of the number of stored transitions (min(pushed, capacity))
        self.defer_small_pushes = False
        self.obs_dim = None
        self.row_floats = None

    # Every access to the handle from outside the fused path first delivers held-back rows, so deferral is invisible:
    # stats, sampling, tree reads, explicit pushes ... all see the replay exactly as the reference would.
    @property
    def handle(self):
        if self._owner_flush is not None:   # a recorded learn() samples the replay as it was when learn() was called
            self._owner_flush()
        if self._pending:
            self.flush()
        return self._handle

    @handle.setter
    def handle(self, h):
        self._handle = h

    def flush(self):
        n, self._pending = self._pending, 0
        if n:
            rc = self._push_fn(self._handle, *self._small_ptrs, n, stream_ptr(self.device_index))
            if rc:
                check(rc)

    def take_pending(self):
        """(n, pointers) of the held-back rows for rmc_learner_step_push; the caller delivers them."""
        n, self._pending = self._pending, 0
        return n

    def ensure(self, obs_dim: int):
        if self.handle is not None:
            if int(obs_dim) != self.obs_dim:
                raise ValueError("observation size changed: %d -> %d" % (self.obs_dim, obs_dim))
            return self
        torch = _lib.require_cuda()
        if self.device_index is None:
            self.device_index = torch.cuda.current_device()
        h = C.c_void_p()
        check(lib().rmc_replay_create(C.byref(h), self.capacity, int(obs_dim), int(self.prioritized),
                                      int(self.device_index)))
        self.handle, self.obs_dim = h, int(obs_dim)
        self.row_floats = lib().rmc_replay_row_floats(h)
        d = self.obs_dim
        self._small = (np.zeros((8, d), np.float32), np.zeros(8, np.int64), np.zeros(8, np.float32),
                       np.zeros(8, np.float32), np.zeros((8, d), np.float32))
        self._small_ptrs = tuple(a.ctypes.data for a in self._small)
        self._push_fn = lib().rmc_replay_push_host
        return self

    def require(self):
        if self.handle is None:
            raise RuntimeError("replay memory is empty (no transition stored yet)")
        return self.handle

    def __del__(self):
        try:
            if self._handle is not None:
                lib().rmc_replay_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # -- host-buffer push (what store_transitions uses) ---------------------------------
    def push_host(self, obses, actions, rews, dones, new_obses):
        n = len(actions)
        if n <= 8 and self.handle is not None:      # (the handle access launches a recorded learn() and delivers rows held back earlier)
            # per-env-step push: copy into preallocated scratch arrays whose addresses are cached
            b = self._small
            b[0][:n] = obses
            b[1][:n] = actions
            b[2][:n] = rews
            b[3][:n] = dones
            b[4][:n] = new_obses
            self.count = min(self.capacity, self.count + n)
            if self.defer_small_pushes:             # the Agent's next learn() carries them (one host call for store + learn)
                self._pending = n
                return
            rc = self._push_fn(self._handle, *self._small_ptrs, n, stream_ptr(self.device_index))
            if rc:
                check(rc)
            return
        obs = np.ascontiguousarray(np.asarray(obses, dtype=np.float32))
        if obs.ndim == 1:
            obs = obs.reshape(1, -1)
        n = obs.shape[0]
        obs = obs.reshape(n, -1)
        nxt = np.ascontiguousarray(np.asarray(new_obses, dtype=np.float32)).reshape(n, -1)
        act = np.ascontiguousarray(np.asarray(actions, dtype=np.int64)).reshape(n)
        rew = np.ascontiguousarray(np.asarray(rews, dtype=np.float32)).reshape(n)
        done = np.ascontiguousarray(np.asarray(dones, dtype=np.float32)).reshape(n)
        self.ensure(obs.shape[1])
        check(lib().rmc_replay_push_host(self.handle, obs.ctypes.data, act.ctypes.data, rew.ctypes.data,
                                         done.ctypes.data, nxt.ctypes.data, n, stream_ptr(self.device_index)))
        self.count = min(self.capacity, self.count + n)

    def push_device(self, obs, act, rew, done, nxt):
        """torch CUDA tensors: obs/nxt float32 [n,D], act int64 [n], rew/done float32 [n]."""
        n = obs.shape[0]
        self.ensure(obs.shape[1])
        check(lib().rmc_replay_push(self.handle, ptr(obs), ptr(act), ptr(rew), ptr(done), ptr(nxt), n, stream_ptr(self.device_index)))
        self.count = min(self.capacity, self.count + n)

    def stats(self) -> _lib.ReplayStats:
        st = _lib.ReplayStats()
        if self.handle is None:
            st.capacity = self.capacity
            return st
        check(lib().rmc_replay_stats_sync(self.handle, C.byref(st), stream_ptr(self.device_index)))
        return st

    def rows_to_transitions(self, rows: np.ndarray):
        """[n,row_floats] float32 rows -> list of (obs, action, rew, done, new_obs) tuples."""
        d = self.obs_dim
        acts = rows[:, 2 * d].copy().view(np.int32)
        return [(rows[i, :d].copy(), int(acts[i]), float(rows[i, 2 * d + 1]), bool(rows[i, 2 * d + 2] != 0.0),
                 rows[i, d:2 * d].copy()) for i in range(rows.shape[0])]

    def read_rows(self, first_slot: int, n: int) -> np.ndarray:
        out = np.empty((n, self.row_floats), np.float32)
        check(lib().rmc_replay_read_rows_sync(self.require(), out.ctypes.data, int(first_slot), int(n), stream_ptr(self.device_index)))
        return out


class _RingView:
    """What ``ReplayMemoryNaive.replay_buffer`` exposes (the reference has a deque there):
    ``len()``, ``maxlen`` and indexing by age (0 = oldest)."""

    def __init__(self, ring: DeviceRing):
        self._ring = ring
        self.maxlen = ring.capacity

    def __len__(self):
        return int(self._ring.stats().size)

    def __getitem__(self, pos):
        st = self._ring.stats()
        n = int(st.size)
        if pos < 0:
            pos += n
        if not 0 <= pos < n:
            raise IndexError(pos)
        slot = (st.data_pointer + pos) % st.capacity if n == st.capacity else pos
        return self._ring.rows_to_transitions(self._ring.read_rows(slot, 1))[0]


class ReplayMemory:
    def __init__(self, buffer_size, batch_size):
        self.batch_size = batch_size
        self.buffer_size = buffer_size

    def store_transitions(self, obses, actions, rews, dones, new_obses):
        raise NotImplementedError

    def sample_transitions(self, step):
        raise NotImplementedError


class ReplayMemoryNaive(ReplayMemory):
    """dqn/replay_memory.py:24-39: bounded FIFO, uniform sampling without replacement."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._ring = DeviceRing(self.buffer_size, prioritized=False)
        self.replay_buffer = _RingView(self._ring)
        self._draws = 0
        self.seed = 0x5EED

    def store_transitions(self, obses, actions, rews, dones, new_obses):
        """Generator (must be iterated, like the reference): appends, then yields the env index of
        every finished episode.  The device push happens on the first ``next()``."""
        self._ring.push_host(obses, actions, rews, dones, new_obses)
        for e, done in enumerate(dones):
            if done:
                yield e

    def sample_transitions(self, step=None, indices=None):
        torch = _lib.require_cuda()
        h = self._ring.require()
        B = int(self.batch_size)
        dev = torch.device("cuda", self._ring.device_index)
        slots = torch.empty(B, dtype=torch.int64, device=dev)
        rows = torch.empty(B, self._ring.row_floats, dtype=torch.float32, device=dev)
        idx = None if indices is None else torch.as_tensor(np.asarray(indices, np.int64), device=dev)
        self._draws += 1
        check(lib().rmc_uniform_sample(h, B, ptr(idx), self.seed, self._draws, ptr(slots), ptr(rows), stream_ptr(self._ring.device_index)))
        return self._ring.rows_to_transitions(rows.cpu().numpy())


class SumTree:
    """dqn/utils/sum_tree.py on the device: same heap layout and indices; ``tree`` / ``data`` are
    materialised on the host on access (tests, debugging)."""

    def __init__(self, capacity, _ring: DeviceRing = None):
        self.capacity = int(capacity)
        self._ring = _ring if _ring is not None else DeviceRing(self.capacity, prioritized=True)

    # -- reference attributes ------------------------------------------------------------
    @property
    def size(self):
        return int(self._ring.stats().size)

    @property
    def data_pointer(self):
        return int(self._ring.stats().data_pointer)

    @property
    def tree(self):
        n = 2 * self.capacity - 1
        out = np.zeros(n, np.float64)
        if self._ring.handle is not None:
            check(lib().rmc_replay_read_tree_sync(self._ring.handle, out.ctypes.data, 0, n, stream_ptr(self._ring.device_index)))
        return out

    @property
    def data(self):
        out = np.zeros(self.capacity, dtype=object)
        st = self._ring.stats()
        if st.size:
            rows = self._ring.read_rows(0, int(st.size))
            for i, t in enumerate(self._ring.rows_to_transitions(rows)):
                out[i] = t
        return out

    @property
    def total_priority(self):
        return self._ring.stats().total_priority

    @property
    def max_priority(self):
        return self._ring.stats().max_priority

    @property
    def min_priority(self):
        return self._ring.stats().min_priority

    # -- reference methods ---------------------------------------------------------------
    def add(self, priority, data):
        """sum_tree.py:34-40 -- ring write + leaf update with an explicit priority."""
        obs, action, rew, done, new_obs = data
        leaf = self.data_pointer + self.capacity - 1 if self._ring.handle is not None else self.capacity - 1
        self._ring.push_host([obs], [action], [rew], [done], [new_obs])
        self.update(leaf, priority)

    def update(self, tree_index, priority):
        """sum_tree.py:15-32."""
        torch = _lib.require_cuda()
        dev = torch.device("cuda", self._ring.device_index)
        nodes = torch.as_tensor([int(tree_index)], dtype=torch.int64, device=dev)
        pri = torch.as_tensor(np.asarray(priority, np.float32).reshape(1), device=dev)
        check(lib().rmc_per_update(self._ring.require(), ptr(nodes), ptr(pri), 1, stream_ptr(self._ring.device_index)))

    def get_leaf(self, v):
        """sum_tree.py:42-61 -> (leaf_index, priority, transition)."""
        torch = _lib.require_cuda()
        dev = torch.device("cuda", self._ring.device_index)
        vv = torch.as_tensor([float(v)], dtype=torch.float64, device=dev)
        node = torch.empty(1, dtype=torch.int64, device=dev)
        pri = torch.empty(1, dtype=torch.float64, device=dev)
        check(lib().rmc_tree_get_leaf(self._ring.require(), ptr(vv), 1, ptr(node), ptr(pri), stream_ptr(self._ring.device_index)))
        leaf = int(node.item())
        row = self._ring.read_rows(leaf - self.capacity + 1, 1)
        return leaf, float(pri.item()), self._ring.rows_to_transitions(row)[0]


class ReplayMemoryPrioritized(ReplayMemory):
    """dqn/replay_memory.py:43-98: proportional prioritisation, stratified sum-tree sampling."""

    def __init__(self, buffer_size, batch_size, eps_dec):
        super().__init__(buffer_size, batch_size)
        self._ring = DeviceRing(self.buffer_size, prioritized=True)
        self.replay_buffer = SumTree(self.buffer_size, _ring=self._ring)
        self.epsilon = 0.0001
        self.alpha = 0.6
        self.beta_start = 0.4
        self.beta_end = 1.
        self.beta_inc = eps_dec
        self.max_priority_high = 1.
        self._draws = 0
        self.seed = 0x5EED

    def beta(self, step):
        return float(np.interp(step, [0, self.beta_inc], [self.beta_start, self.beta_end]))

    def store_transitions(self, obses, actions, rews, dones, new_obses):
        # new leaves get max_priority (1.0 when it is 0), read once per call -- done by the push kernel
        self._ring.push_host(obses, actions, rews, dones, new_obses)
        for e, done in enumerate(dones):
            if done:
                yield e

    def sample_transitions(self, step, u=None):
        """-> (is_weights, tree_indices, transitions) as python lists.  ``u`` injects the uniforms;
        default: consumes ``np.random.random_sample(batch)`` so a seeded numpy RNG reproduces the
        reference's draw sequence (np.random.uniform(lo,hi) == lo+(hi-lo)*random_sample())."""
        torch = _lib.require_cuda()
        h = self._ring.require()
        B = int(self.batch_size)
        dev = torch.device("cuda", self._ring.device_index)
        if u is None:
            u = np.random.random_sample(B)
        u_t = torch.as_tensor(np.asarray(u, np.float64), device=dev)
        nodes = torch.empty(B, dtype=torch.int64, device=dev)
        w = torch.empty(B, dtype=torch.float32, device=dev)
        rows = torch.empty(B, self._ring.row_floats, dtype=torch.float32, device=dev)
        self._draws += 1
        check(lib().rmc_per_sample(h, B, self.beta(step), ptr(u_t), self.seed, self._draws, ptr(nodes), ptr(w),
                                   ptr(rows), stream_ptr(self._ring.device_index)))
        return (w.cpu().numpy().astype(np.float64).tolist(), nodes.cpu().tolist(),
                self._ring.rows_to_transitions(rows.cpu().numpy()))

    def update_batch_priorities(self, tree_indices, abs_td_errors_np):
        torch = _lib.require_cuda()
        dev = torch.device("cuda", self._ring.device_index)
        idx = np.asarray(tree_indices, np.int64).reshape(-1)
        cap = self._ring.capacity
        if idx.size and (idx.min() < cap - 1 or idx.max() > 2 * cap - 2):      # the reference would raise IndexError / corrupt inner nodes
            raise IndexError("tree_indices outside the leaf range [%d, %d]" % (cap - 1, 2 * cap - 2))
        nodes = torch.as_tensor(idx, device=dev)
        td = torch.as_tensor(np.asarray(abs_td_errors_np, np.float32).reshape(-1), device=dev)
        pri = torch.empty_like(td)
        check(lib().rmc_per_update_from_td(self._ring.require(), ptr(nodes), ptr(td), nodes.numel(), self.epsilon,
                                           self.alpha, self.max_priority_high, ptr(pri), stream_ptr(self._ring.device_index)))
