"""torch custom-op registration of the hot-path entry points (``torch.ops.rmc_b200.*``).

The ops are thin: they take the opaque C-ABI handles as python ints plus torch tensors, and forward to
librmc_b200 on the tensors' device pointers and torch's current stream (no copies).  They exist so the path can be
driven from code that speaks ``torch.ops`` (dispatcher, profiler ranges, ``torch.library.opcheck``)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def _h(handle) -> int:
    return handle.value if isinstance(handle, C.c_void_p) else int(handle)


@torch.library.custom_op("rmc_b200::per_sample", mutates_args=())
def per_sample(replay: int, batch: int, beta: float, u: Tensor, row_floats: int) -> Tuple[Tensor, Tensor, Tensor]:
    """ReplayMemoryPrioritized.sample_transitions (dqn/replay_memory.py:69-92) for injected uniforms."""
    nodes = torch.empty(batch, dtype=torch.int64, device=u.device)
    w = torch.empty(batch, dtype=torch.float32, device=u.device)
    rows = torch.empty(batch, row_floats, dtype=torch.float32, device=u.device)
    check(lib().rmc_per_sample(replay, batch, beta, ptr(u), 0, 0, ptr(nodes), ptr(w), ptr(rows), stream_ptr(u.device.index)))
    return nodes, w, rows


@torch.library.custom_op("rmc_b200::per_update", mutates_args=())
def per_update(replay: int, nodes: Tensor, priorities: Tensor) -> Tensor:
    """SumTree.update applied in batch order (dqn/replay_memory.py:97-98); returns the batch size."""
    check(lib().rmc_per_update(replay, ptr(nodes), ptr(priorities), nodes.numel(), stream_ptr(nodes.device.index)))
    return torch.tensor(nodes.numel())


@torch.library.custom_op("rmc_b200::act", mutates_args=())
def act(learner: int, obs: Tensor) -> Tensor:
    """Network.actions (dqn/network.py:67-74,110-117) on a device tensor [n, D]."""
    out = torch.empty(obs.shape[0], dtype=torch.int64, device=obs.device)
    check(lib().rmc_learner_act(learner, ptr(obs), obs.shape[0], ptr(out), stream_ptr(obs.device.index)))
    return out


@torch.library.custom_op("rmc_b200::q_values", mutates_args=())
def q_values(learner: int, which: int, obs: Tensor, n_actions: int) -> Tensor:
    """Network.forward (dqn/network.py:59-63,90-96)."""
    out = torch.empty(obs.shape[0], n_actions, dtype=torch.float32, device=obs.device)
    check(lib().rmc_learner_q_values(learner, which, ptr(obs), obs.shape[0], ptr(out), stream_ptr(obs.device.index)))
    return out


@torch.library.custom_op("rmc_b200::learner_step", mutates_args=())
def learner_step(learner: int, replay: int, device_index: int, batch: int, phases: int, beta: float, u: Optional[Tensor], idx: Optional[Tensor],
                 seed: int, counter: int, adam_t: int) -> Tensor:
    """{Simple,Double,PerDouble}Agent.learn + update_target_network (dqn/agent.py:101-110,166-272); returns the loss."""
    a = _lib.StepArgs()
    a.batch, a.phases, a.per_beta, a.seed, a.counter, a.adam_t = batch, phases, beta, seed, counter, adam_t
    a.u_dev = None if u is None else u.data_ptr()
    a.idx_dev = None if idx is None else idx.data_ptr()
    check(lib().rmc_learner_step(learner, replay, C.byref(a), stream_ptr(device_index)))
    p, n = C.c_void_p(), C.c_int64()
    check(lib().rmc_learner_output(learner, b"loss", C.byref(p), C.byref(n)))
    from .network import _tensor_from_ptr
    return _tensor_from_ptr(p.value, 1, torch.float32, device_index).clone()


# shape-only implementations (tracing, torch.library.opcheck) for every op
@per_update.register_fake
def _(replay, nodes, priorities):
    return torch.empty((), dtype=torch.int64)


@learner_step.register_fake
def _(learner, replay, device_index, batch, phases, beta, u, idx, seed, counter, adam_t):
    return torch.empty(1, dtype=torch.float32, device=torch.device("cuda", device_index))


@per_sample.register_fake
def _(replay, batch, beta, u, row_floats):
    return (u.new_empty(batch, dtype=torch.int64), u.new_empty(batch, dtype=torch.float32),
            u.new_empty(batch, row_floats, dtype=torch.float32))


@act.register_fake
def _(learner, obs):
    return obs.new_empty(obs.shape[0], dtype=torch.int64)


@q_values.register_fake
def _(learner, which, obs, n_actions):
    return obs.new_empty(obs.shape[0], n_actions)
