"""Multi-GPU forms of the learner path.  One process per GPU, torch.distributed for the
plumbing (NCCL on GPUs, gloo in the CPU tests of the host logic).

* ``AgentEnsemble``  -- N independent agents of ONE GPU stepped by a single launch (``rmc_group_*``); across
  GPUs ensembles need no communication at all (config C4).
* ``ShardedLearner`` -- large-batch learner (config C5): the minibatch is split across ranks, every rank keeps
  a full replica of replay, tree, weights and Adam state; per step one gradient all-reduce (+ an all-gather of
  (leaf, |td|) for PER so that every replica applies the identical write-back).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch as T

from . import _lib
from ._lib import check, lib, stream_ptr


def shard_range(batch: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the global stratified sample indices owned by ``rank``; slices tile [0, batch)."""
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class AgentEnsemble:
    """Agents sharing spec / hyper-parameters / batch size on one GPU, stepped together.  Each member keeps its own
    replay ring, tree, weights, Adam state and RNG stream (Philox counter = (step, agent index))."""

    def __init__(self, agents):
        self.agents = list(agents)
        n = len(self.agents)
        a0 = self.agents[0]
        for a in self.agents:       # the launch carries ONE set of scalars (agent 0's): members must agree on them
            a._flush_step()
            same = (type(a) is type(a0) and a.n_env == a0.n_env and a.batch_size == a0.batch_size and a.lr == a0.lr and a.gamma == a0.gamma
                    and a.target_soft_update == a0.target_soft_update and a.target_soft_update_tau == a0.target_soft_update_tau
                    and a.update_target_frequency == a0.update_target_frequency and a.epsilon_decay == a0.epsilon_decay)
            if not same:
                raise ValueError("AgentEnsemble members must share class, n_env, batch size, lr, gamma, target-update and epsilon-decay settings")
        lh = (C.c_void_p * n)(*[a._lh.handle for a in self.agents])
        rh = (C.c_void_p * n)(*[a.replay_memory_buffer._ring.require() for a in self.agents])
        g = C.c_void_p()
        check(lib().rmc_group_create(C.byref(g), lh, rh, n))
        self.handle = g
        self._args = _lib.StepArgs()
        self._args.batch = int(self.agents[0].batch_size)
        self._dev = self.agents[0].device.index

    def __del__(self):
        try:
            if self.handle is not None:
                lib().rmc_group_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _deliver_pending(self):
        """The rows every member's ``store_transitions`` held back for this env step: ONE launch for the whole ensemble
        (``rmc_group_push_host``) when all members hold the same number of rows, else one small push per member."""
        rings = [a.replay_memory_buffer._ring for a in self.agents]
        n = rings[0]._pending
        if n and all(r._pending == n and r._handle is not None for r in rings):
            k = len(rings)
            cols = [(C.c_void_p * k)(*[r._small_ptrs[j] for r in rings]) for j in range(5)]
            check(lib().rmc_group_push_host(self.handle, *cols, n, stream_ptr(self._dev)))
            for r in rings:
                r._pending = 0
            return
        for r in rings:
            r.flush()

    def learn(self, fuse_target_update=True, u=None, indices=None):
        """One learner step of every member (dqn/agent.py learn() + update_target_network()).
        ``u`` / ``indices``: optional injected sampling randomness, shape [n_agents, batch]."""
        a0 = self.agents[0]
        if any(a.step != a0.step for a in self.agents):
            raise ValueError("AgentEnsemble.learn: members are at different steps (beta and the hard-sync schedule come from one step value)")
        for a in self.agents:
            a._flush_step()                          # a lazily recorded single-agent learn()
        self._deliver_pending()                      # rows held back by the members' store_transitions of this env step
        for a in self.agents:
            a._learn_calls += 1
            a._adam_t += 1
        args = self._args
        args.phases = a0._learn_phases | (a0._target_phase() if fuse_target_update else 0)
        args.counter = a0._learn_calls
        args.adam_t = a0._adam_t
        args.seed = a0.sampling_seed
        if a0._PER:
            args.per_beta = a0._beta(a0.step * a0.n_env)
        args.precision = a0._args.precision      # "bf16": the members' tensor-core steps side by side (see rmc_group_step)
        args.u_dev = args.idx_dev = None
        keep = None
        if u is not None:
            keep = T.as_tensor(np.ascontiguousarray(np.asarray(u, np.float64)), device=a0.device)
            args.u_dev = keep.data_ptr()
        if indices is not None:
            keep = T.as_tensor(np.ascontiguousarray(np.asarray(indices, np.int64)), device=a0.device)
            args.idx_dev = keep.data_ptr()
        self._keep = keep
        check(lib().rmc_group_step(self.handle, C.byref(args), stream_ptr(self._dev)))
        for a in self.agents:
            a._lh.version[_lib.ONLINE] += 1
            if fuse_target_update:
                a._lh.version[_lib.TARGET] += 1
                a._target_fused_for = a._learn_calls


class ShardedLearner:
    """Data-parallel learner step for one logical agent replicated on every rank.

    Per step and rank r of W (global batch B, slice [lo, hi) = shard_range(B, r, W)):
      1. sample the slice's strata with the GLOBAL segment length total/B and the global uniforms u[lo:hi]
         (the union over ranks is exactly the single-GPU batch), forward, TD, dgrad, weight gradients scaled
         by 1/B -> local gradient blob;
      2. all-reduce(sum) of the gradient blob (P floats) and of the loss partial;
      3. PER: all-gather (leaf index, |td|) and apply the full write-back on every replica in global batch order;
      4. Adam (+ Polyak) from the reduced gradients, identical on every rank.
    """

    def __init__(self, agent, group=None, exchange="nccl", rank=None, world=None):
        """``exchange="nccl"``: torch.distributed collectives (all-reduce + all-gather) around two library calls.
        ``exchange="peer"``: the exchange runs inside the library as kernels over NVLink peer memory (gradient blobs
        read straight from the peers' buffers and reduced in rank order in the same kernel that applies Adam); the
        whole step is ONE C call with no collective and no host synchronisation.  The peers' buffers are mapped with
        CUDA IPC handles all-gathered once at construction (``connect_same_process`` for ranks emulated in one
        process)."""
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' or 'peer'")
        self.agent, self.group, self.exchange = agent, group, exchange
        if rank is None:
            import torch.distributed as dist
            self.dist = dist
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:                               # explicit placement (ranks emulated inside one process)
            self.dist = None
            self.rank, self.world = int(rank), int(world)
        self.B = int(agent.batch_size)
        self.lo, self.hi = shard_range(self.B, self.rank, self.world)
        self._args = _lib.StepArgs()
        self._comm = None
        self.status_every = 64      # peer exchange: poll the comm's error word every this many steps (a host-side sync)
        self._steps = 0
        if exchange == "peer":
            h = C.c_void_p()
            check(lib().rmc_comm_create(C.byref(h), agent._lh.handle, self.rank, self.world, self.B))
            self._comm = h
            if self.dist is not None:
                self._connect_ipc()

    def __del__(self):
        try:
            if self._comm is not None:
                lib().rmc_comm_destroy(self._comm)
                self._comm = None
        except Exception:
            pass

    def _export(self):
        buf = (C.c_ubyte * 64)()
        p = C.c_void_p()
        check(lib().rmc_comm_export(self._comm, buf, C.byref(p)))
        return bytes(buf), p.value

    def _connect_ipc(self):
        handle, _ = self._export()
        dev = self.agent.device
        mine = T.tensor(list(handle), dtype=T.uint8, device=dev)
        parts = [T.empty(64, dtype=T.uint8, device=dev) for _ in range(self.world)]
        self.dist.all_gather(parts, mine, group=self.group)
        blob = b"".join(bytes(p.cpu().tolist()) for p in parts)
        check(lib().rmc_comm_connect(self._comm, blob, None))
        self.dist.barrier(group=self.group)

    @staticmethod
    def connect_same_process(members):
        """Wire the exchange buffers of ``members`` (one ShardedLearner per emulated rank, all in this process)."""
        ptrs = [m._export()[1] for m in members]
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        for m in members:
            check(lib().rmc_comm_connect(m._comm, None, arr))
            # Emulated ranks share ONE device: a waiting exchange kernel of one rank and the fused step of another (a whole
            # SM per CTA) cannot be co-resident, so the second half of a split step (stages=2) is ordered after every
            # member's first half by events.  Real ranks own a GPU each and need no such edge.
            m._same_process = list(members)

    def _learn_peer(self, u, fuse_target_update, stages=3):
        ag = self.agent
        ag._flush_step()
        if stages == 2:       # second half of a split step (emulated ranks): same arguments as the first half
            for m in getattr(self, "_same_process", ()):
                ev = getattr(m, "_stage1_event", None)
                if ev is not None:
                    T.cuda.current_stream(ag.device).wait_event(ev)
            check(lib().rmc_learner_step_sharded(ag._lh.handle, ag.replay_memory_buffer._ring.require(), self._comm,
                                                 C.byref(self._args), 2, stream_ptr(ag.device.index)))
            return ag._lh.output("loss")
        ag._learn_calls += 1
        ag._adam_t += 1
        a = self._args
        a.batch, a.global_batch, a.shard_offset = self.hi - self.lo, self.B, self.lo
        a.phases = ag._learn_phases | (ag._target_phase() if fuse_target_update else 0)
        a.seed, a.counter, a.adam_t = ag.sampling_seed, ag._learn_calls, ag._adam_t
        a.grads_in_dev, a.u_dev, a.idx_dev = None, None, None
        a.precision = ag._args.precision
        keep = None
        if ag._PER:
            a.per_beta = ag._beta(ag.step * ag.n_env)
            if u is not None:
                keep = T.as_tensor(np.asarray(u, np.float64)[self.lo:self.hi].copy(), device=ag.device)
                a.u_dev = keep.data_ptr()
        rh = ag.replay_memory_buffer._ring.require()
        check(lib().rmc_learner_step_sharded(ag._lh.handle, rh, self._comm, C.byref(a), stages, stream_ptr(ag.device.index)))
        if stages == 1 and getattr(self, "_same_process", None):
            self._stage1_event = T.cuda.Event()
            self._stage1_event.record(T.cuda.current_stream(ag.device))
        ag._lh.version[_lib.ONLINE] += 1
        if fuse_target_update:
            ag._lh.version[_lib.TARGET] += 1
            ag._target_fused_for = ag._learn_calls
        self._keep = keep
        self._steps += 1
        if self.status_every and self._steps % self.status_every == 0 and stages == 3:
            bad = self.exchange_status()
            if bad:
                raise RuntimeError("ShardedLearner: the peer exchange of step epoch %d timed out (a peer never published); "
                                   "that step was skipped on this rank -- replicas may have diverged" % bad)
        return ag._lh.output("loss")

    def exchange_status(self):
        """0, or the epoch of an exchange that timed out because a peer never published (synchronises)."""
        if self._comm is None:
            return 0
        v = C.c_uint32(0)
        check(lib().rmc_comm_status_sync(self._comm, C.byref(v), stream_ptr(self.agent.device.index)))
        return int(v.value)

    def learn(self, u=None, fuse_target_update=True, stages=3):
        if self.exchange == "peer":
            return self._learn_peer(u, fuse_target_update, stages)
        ag, dist = self.agent, self.dist
        ag._flush_step()
        lh = ag._lh
        rh = ag.replay_memory_buffer._ring.require()
        ag._learn_calls += 1
        ag._adam_t += 1
        a = self._args
        n_local = self.hi - self.lo
        a.batch, a.global_batch, a.shard_offset = n_local, self.B, self.lo
        a.phases = _lib.PH_SAMPLE | _lib.PH_FORWARD | _lib.PH_BACKWARD
        a.seed, a.counter, a.adam_t = ag.sampling_seed, ag._learn_calls, ag._adam_t
        a.grads_in_dev = None
        a.precision = ag._args.precision      # "bf16": tensor-core forward/backward, fp32 all-reduce + Adam
        keep = None
        if ag._PER:
            a.per_beta = ag._beta(ag.step * ag.n_env)
            if u is not None:
                keep = T.as_tensor(np.asarray(u, np.float64)[self.lo:self.hi].copy(), device=ag.device)
                a.u_dev = keep.data_ptr()
            else:
                a.u_dev = None
        s = stream_ptr(ag.device.index)
        check(lib().rmc_learner_step(lh.handle, rh, C.byref(a), s))
        grads = lh.output("grads_blob")
        loss = lh.output("loss")
        dist.all_reduce(grads, group=self.group)
        dist.all_reduce(loss, group=self.group)
        if ag._PER:
            nodes = lh.output("nodes", T.int64)[:n_local]
            td = lh.output("abs_td")[:n_local]
            sizes = [shard_range(self.B, r, self.world) for r in range(self.world)]
            all_nodes = [T.empty(h - l, dtype=T.int64, device=ag.device) for l, h in sizes]
            all_td = [T.empty(h - l, dtype=T.float32, device=ag.device) for l, h in sizes]
            dist.all_gather(all_nodes, nodes.contiguous(), group=self.group)
            dist.all_gather(all_td, td.contiguous(), group=self.group)
            gn, gt = T.cat(all_nodes), T.cat(all_td)
            pri = T.empty_like(gt)
            mem = ag.replay_memory_buffer
            check(lib().rmc_per_update_from_td(rh, gn.data_ptr(), gt.data_ptr(), gn.numel(), mem.epsilon, mem.alpha,
                                               mem.max_priority_high, pri.data_ptr(), s))
        b = _lib.StepArgs()
        b.batch, b.adam_t = n_local, ag._adam_t
        b.phases = _lib.PH_ADAM | (ag._target_phase() if fuse_target_update else 0)
        b.grads_in_dev = grads.data_ptr()
        check(lib().rmc_learner_step(lh.handle, rh, C.byref(b), s))
        lh.version[_lib.ONLINE] += 1
        if fuse_target_update:
            lh.version[_lib.TARGET] += 1
            ag._target_fused_for = ag._learn_calls
        self._keep = (keep, grads, loss)
        return loss


def sharded_act(network, obses, group=None, gather=True, precision="fp32", rank=None, world=None):
    """Batched greedy ``Network.actions`` (dqn/network.py:67-74, 110-117) with the rows split across the ranks (BASELINE
    configs[2] on several GPUs): rows are independent, every rank holds a replica of the weights, rank r
    evaluates rows ``shard_range(n, r, W)`` with the batched act kernel and -- ``gather=True`` -- the int64 actions are
    all-gathered so every rank returns the full list in row order (the only communication: 8 bytes per state).  With
    ``gather=False`` the rank's own slice is returned together with its row range.

    ``obses``: [n, D] host array or CUDA tensor holding ALL rows (each rank reads only its slice)."""
    if rank is None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        dist = None
    n = len(obses)
    lo, hi = shard_range(n, rank, world)
    mine = network.actions(obses[lo:hi], precision=precision) if hi > lo else []
    if not gather:
        return mine, (lo, hi)
    if dist is None:
        raise ValueError("gather=True needs an initialised torch.distributed process group")
    dev = T.device(network.device)
    sizes = [shard_range(n, r, world) for r in range(world)]
    parts = [T.empty(h - l, dtype=T.int64, device=dev) for l, h in sizes]
    dist.all_gather(parts, T.as_tensor(mine, dtype=T.int64, device=dev), group=group)
    return T.cat(parts).tolist()
