"""Q-networks with the reference's ``Network`` surface (dqn/network.py) whose compute runs in
librmc_b200 (fused CUDA kernels) instead of torch eager.

The ``nn.Module`` tree (``net``, ``fc_val``/``fc_adv`` or ``fc_out``) is kept so that
``state_dict()`` keys, ``load_state_dict``, ``parameters()`` and the msgpack ``.pack`` checkpoint
(network.py:27-47) are byte-compatible; the authoritative copy of the weights lives in the
learner's HBM blob and the module tensors are synchronised lazily in either direction.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch as T
import torch.nn as nn

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from . import packfmt


def _hidden_activation(mod):
    """0 for nn.ReLU, 1 for nn.ELU(alpha=1) (the activation of env/dqn_config.py:175), else None."""
    if isinstance(mod, nn.ReLU):
        return _lib.ACT_RELU
    if isinstance(mod, nn.ELU) and float(mod.alpha) == 1.0:
        return _lib.ACT_ELU
    return None


def match_hybrid_body(net, fc_out_dim):
    """The repo-HEAD body (env/dqn_config.py:66-143 ``TwoStreamHybridNetwork``), recognised structurally: attributes
    ``macro_len``, ``micro_shape`` (C, H, W), ``cnn_stream`` = Sequential of (Conv2d 3x3 padding 1, act) pairs and
    ``dense_stream`` = Sequential of (Linear, act) pairs, one activation kind throughout, parameters registered in that
    order.  Returns a dict for ``rmc_hybrid_spec_t`` or None when ``net`` is not of this family."""
    if not all(hasattr(net, a) for a in ("macro_len", "micro_shape", "cnn_stream", "dense_stream")):
        return None
    cnn, dense = net.cnn_stream, net.dense_stream
    if not (isinstance(cnn, nn.Sequential) and isinstance(dense, nn.Sequential) and len(cnn) % 2 == 0 and len(dense) % 2 == 0
            and 2 <= len(cnn) <= 8 and 2 <= len(dense) <= 6):
        return None
    acts = set()
    convs, denses = [], []
    for i in range(0, len(cnn), 2):
        c, a = cnn[i], cnn[i + 1]
        if not (isinstance(c, nn.Conv2d) and tuple(c.kernel_size) == (3, 3) and tuple(c.padding) == (1, 1) and tuple(c.dilation) == (1, 1)
                and c.groups == 1 and c.bias is not None and c.padding_mode == "zeros" and _hidden_activation(a) is not None):
            return None
        acts.add(_hidden_activation(a))
        convs.append(c)
    for i in range(0, len(dense), 2):
        d, a = dense[i], dense[i + 1]
        if not (isinstance(d, nn.Linear) and d.bias is not None and _hidden_activation(a) is not None):
            return None
        acts.add(_hidden_activation(a))
        denses.append(d)
    ch, hh, ww = (int(v) for v in net.micro_shape)
    keys = [k for k, _ in net.named_parameters()]
    want = ["cnn_stream.%d.%s" % (2 * i, n) for i in range(len(convs)) for n in ("weight", "bias")] + \
           ["dense_stream.%d.%s" % (2 * i, n) for i in range(len(denses)) for n in ("weight", "bias")]
    if len(acts) != 1 or keys != want or convs[0].in_channels != ch or denses[-1].out_features != fc_out_dim:
        return None
    return dict(macro_len=int(net.macro_len), grid=(ch, hh, ww), conv_out=[c.out_channels for c in convs],
                conv_stride=[tuple(int(v) for v in c.stride) for c in convs], dense_out=[d.out_features for d in denses],
                activation=acts.pop(), obs_dim=int(net.macro_len) + ch * hh * ww)


def match_macro_body(net, fc_out_dim, optim_func, loss_func):
    """The fused path is built for the macro-state body of the reference
    (env/custom_env/macro with lane/dqn_config.py:58-104): Sequential(Linear(D,256), act,
    Linear(256,128), act) with act = ReLU or ELU(alpha=1) (the same in both places) + optim.Adam +
    nn.SmoothL1Loss.  Anything else raises: there is no eager / CPU fallback by design.
    Returns (obs_dim, activation code)."""
    ok = (isinstance(net, nn.Sequential) and len(net) == 4 and isinstance(net[0], nn.Linear)
          and isinstance(net[2], nn.Linear) and _hidden_activation(net[1]) is not None
          and _hidden_activation(net[1]) == _hidden_activation(net[3])
          and net[0].out_features == 256 and net[2].in_features == 256 and net[2].out_features == 128
          and fc_out_dim == 128 and net[0].bias is not None and net[2].bias is not None)
    if not ok:
        raise NotImplementedError("librmc_b200 implements the macro-state MLP body Linear(D,256)-act-Linear(256,128)-act with "
                                  "act = ReLU or ELU(alpha=1) only (got %r); no fallback path exists" % (net,))
    if optim_func is not T.optim.Adam:
        raise NotImplementedError("librmc_b200 fuses torch.optim.Adam only (got %r)" % (optim_func,))
    if loss_func is not nn.SmoothL1Loss:
        raise NotImplementedError("librmc_b200 fuses nn.SmoothL1Loss only (got %r)" % (loss_func,))
    return net[0].in_features, _hidden_activation(net[1])


class LearnerHandle:
    """Owner of one ``rmc_learner_t`` (online + target + Adam state + scratch)."""

    def __init__(self, obs_dim, n_actions, dueling, double_dqn, prioritized, max_batch, device_index, hyper: _lib.Hyper,
                 activation=0, hybrid=None):
        _lib.require_cuda()
        self.hyper = hyper
        self.device_index = int(device_index)
        self.max_batch = int(max_batch)
        self.hybrid = hybrid
        h = C.c_void_p()
        if hybrid is not None:       # the repo-HEAD CNN + MLP body (match_hybrid_body)
            sp = _lib.HybridSpec()
            sp.macro_len = hybrid["macro_len"]
            sp.grid_c, sp.grid_h, sp.grid_w = hybrid["grid"]
            sp.n_conv, sp.n_dense = len(hybrid["conv_out"]), len(hybrid["dense_out"])
            for i, (o, st) in enumerate(zip(hybrid["conv_out"], hybrid["conv_stride"])):
                sp.conv_out[i], sp.conv_sh[i], sp.conv_sw[i] = int(o), int(st[0]), int(st[1])
            for i, o in enumerate(hybrid["dense_out"]):
                sp.dense_out[i] = int(o)
            sp.n_actions, sp.dueling, sp.double_dqn, sp.prioritized = int(n_actions), int(dueling), int(double_dqn), int(prioritized)
            sp.activation = int(hybrid["activation"])
            self.spec = sp
            check(lib().rmc_learner_create_hybrid(C.byref(h), C.byref(sp), C.byref(self.hyper), self.max_batch, self.device_index))
        else:
            self.spec = _lib.NetSpec(int(obs_dim), 256, 128, int(n_actions), int(dueling), int(double_dqn),
                                     int(prioritized), int(activation))
            check(lib().rmc_learner_create(C.byref(h), C.byref(self.spec), C.byref(self.hyper), self.max_batch,
                                           self.device_index))
        self.handle = h
        self.n_params = int(lib().rmc_learner_param_count(h))
        self.version = [0, 0]   # bumped whenever the blob of kind ONLINE / TARGET changes on the device
        self._flush_hook = None  # set by the owning Agent: launches a learn() step it has recorded but not issued yet

    def flush_pending(self):
        """Agent.learn() is lazy (agent.py); every read or write of the learner's state goes through here first."""
        if self._flush_hook is not None:
            self._flush_hook()

    def __del__(self):
        try:
            if self.handle is not None:
                lib().rmc_learner_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_params(self, kind, flat):
        self.flush_pending()
        flat = flat.contiguous()
        check(lib().rmc_learner_set_params(self.handle, kind, ptr(flat), flat.numel(), int(not flat.is_cuda), stream_ptr(self.device_index)))

    def get_params(self, kind, device=None):
        self.flush_pending()
        dev = T.device("cuda", self.device_index) if device is None else device
        out = T.empty(self.n_params, dtype=T.float32, device=dev)
        check(lib().rmc_learner_get_params(self.handle, kind, ptr(out), out.numel(), int(not out.is_cuda), stream_ptr(self.device_index)))
        return out

    def output(self, name, dtype=T.float32):
        """Zero-copy torch view of a per-step product (valid until the next step)."""
        self.flush_pending()
        p, n = C.c_void_p(), C.c_int64()
        check(lib().rmc_learner_output(self.handle, name.encode(), C.byref(p), C.byref(n)))
        return _tensor_from_ptr(p.value, n.value, dtype, self.device_index)


def _tensor_from_ptr(addr, n, dtype, device_index):
    """Wrap device memory owned by the library as a torch tensor (no copy) via __cuda_array_interface__."""
    np_t = {T.float32: "<f4", T.int64: "<i8", T.float64: "<f8"}[dtype]

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(n),), "typestr": np_t, "data": (int(addr), False), "version": 3,
                                  "strides": None}
    return T.as_tensor(h, device=T.device("cuda", device_index))


class Network(nn.Module):
    """dqn/network.py:11-47."""

    def __init__(self, device, nn_conf_func, input_dim):
        super().__init__()
        self.net, self.fc_out_dim, optim_func, loss_func = nn_conf_func(input_dim)
        self._hybrid = match_hybrid_body(self.net, self.fc_out_dim)
        if self._hybrid is not None:
            if optim_func is not T.optim.Adam or loss_func is not nn.SmoothL1Loss:
                raise NotImplementedError("librmc_b200 fuses torch.optim.Adam + nn.SmoothL1Loss only")
            self._obs_dim, self._activation = self._hybrid["obs_dim"], self._hybrid["activation"]
        else:
            self._obs_dim, self._activation = match_macro_body(self.net, self.fc_out_dim, optim_func, loss_func)
        self.optim_func = (lambda params, lr: optim_func(params, lr=lr))
        self.loss_func = (lambda reduction: loss_func(reduction=reduction))
        self.device = device
        # binding to a learner blob (set by the Agent, or created on demand for stand-alone use)
        self._lh = None
        self._kind = _lib.ONLINE
        self._seen_version = -1     # blob version the module tensors correspond to
        self._module_dirty = True   # module tensors changed since they were last pushed

    # ---- binding ---------------------------------------------------------------------------
    def _bind(self, handle: LearnerHandle, kind: int):
        self._lh, self._kind = handle, kind
        self._module_dirty = True
        self._push()

    def _standalone_handle(self):
        if self._lh is None:
            dev = T.device(self.device)
            if dev.type != "cuda":
                raise RuntimeError("Network compute needs a CUDA device (no CPU fallback); got device=%s" % (self.device,))
            index = dev.index if dev.index is not None else T.cuda.current_device()
            hyper = _lib.Hyper(1e-4, 0.9, 0.999, 1e-8, 0.99, 1e-3, 1e-4, 0.6, 1.0)
            self._bind(LearnerHandle(self._obs_dim, self._n_actions, self._dueling, True, False, 1, index, hyper,
                                     activation=self._activation, hybrid=self._hybrid), _lib.ONLINE)
        return self._lh

    def _flat_module_params(self):
        return T.cat([v.detach().reshape(-1).to(T.float32) for v in nn.Module.state_dict(self).values()])

    def _push(self):
        """module tensors -> device blob."""
        if self._lh is not None:
            self._lh.flush_pending()
        if self._lh is not None and self._module_dirty:
            self._lh.set_params(self._kind, self._flat_module_params())
            self._lh.version[self._kind] += 1
            self._seen_version = self._lh.version[self._kind]
            self._module_dirty = False

    def _pull(self):
        """device blob -> module tensors (after learner steps)."""
        if self._lh is not None:
            self._lh.flush_pending()
        if self._lh is not None and not self._module_dirty and self._seen_version != self._lh.version[self._kind]:
            flat = self._lh.get_params(self._kind)
            off = 0
            with T.no_grad():
                for v in nn.Module.state_dict(self).values():
                    n = v.numel()
                    v.copy_(flat[off:off + n].reshape(v.shape))
                    off += n
            self._seen_version = self._lh.version[self._kind]

    # ---- nn.Module surface kept coherent with the blob -----------------------------------------
    def state_dict(self, *args, **kwargs):
        self._pull()
        return super().state_dict(*args, **kwargs)

    def parameters(self, recurse=True):
        self._pull()
        return super().parameters(recurse)

    def named_parameters(self, *args, **kwargs):
        self._pull()
        return super().named_parameters(*args, **kwargs)

    def load_state_dict(self, state_dict, *args, **kwargs):
        res = super().load_state_dict(state_dict, *args, **kwargs)
        self._module_dirty = True
        self._push()
        return res

    def mark_modified(self):
        """Call after editing module tensors in place (e.g. ``p.data.copy_``)."""
        self._module_dirty = True
        self._push()

    def forward(self, s):
        """Q values (network.py:59-63 / 90-96) computed by the CUDA inference kernel."""
        lh = self._standalone_handle()
        self._push()
        dev = T.device("cuda", lh.device_index)
        x = T.as_tensor(s, dtype=T.float32, device=dev).contiguous()
        squeeze = x.dim() == 1
        x = x.reshape(-1, self._obs_dim)
        q = T.empty(x.shape[0], self._n_actions, dtype=T.float32, device=dev)
        check(lib().rmc_learner_q_values(lh.handle, self._kind, ptr(x), x.shape[0], ptr(q), stream_ptr(lh.device_index)))
        return q[0] if squeeze else q

    def actions(self, obses, precision="fp32"):
        """Greedy actions as a python list (network.py:67-74 / 110-117); host buffers in and out.
        ``precision="bf16"`` selects the tcgen05 tensor-core kernel (dense batches; looser stated bound)."""
        lh = self._standalone_handle()
        self._push()
        if self._kind != _lib.ONLINE:
            raise RuntimeError("actions() is served from the online blob")
        if precision == "bf16":
            dev = T.device("cuda", lh.device_index)
            x = T.as_tensor(obses, dtype=T.float32, device=dev).contiguous().reshape(-1, self._obs_dim)
            out = T.empty(x.shape[0], dtype=T.int64, device=dev)
            check(lib().rmc_learner_act_tc(lh.handle, ptr(x), x.shape[0], ptr(out), stream_ptr(lh.device_index)))
            return out.tolist()
        if precision != "fp32":
            raise ValueError("precision must be 'fp32' or 'bf16'")
        if T.is_tensor(obses) and obses.is_cuda:
            x = obses.to(T.float32).contiguous().reshape(-1, self._obs_dim)
            out = T.empty(x.shape[0], dtype=T.int64, device=x.device)
            check(lib().rmc_learner_act(lh.handle, ptr(x), x.shape[0], ptr(out), stream_ptr(lh.device_index)))
            return out.tolist()
        x = np.ascontiguousarray(np.asarray(obses, dtype=np.float32)).reshape(-1, self._obs_dim)
        out = np.empty(x.shape[0], np.int64)
        check(lib().rmc_learner_act_host_sync(lh.handle, x.ctypes.data, x.shape[0], out.ctypes.data, stream_ptr(lh.device_index)))
        return out.tolist()

    # ---- checkpoints: format of network.py:27-47, byte compatible ------------------------------
    def save(self, save_path, step, episode_count, rew_mean, len_mean):
        params_dict = {
            'parameters': {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()},
            'step': step, 'episode_count': episode_count, 'rew_mean': rew_mean, 'len_mean': len_mean
        }
        os.makedirs(os.path.dirname(save_path), exist_ok=True)
        with open(save_path, 'wb') as f:
            f.write(packfmt.dumps(params_dict))

    def load(self, load_path):
        if not os.path.exists(load_path):
            raise FileNotFoundError(load_path)
        with open(load_path, 'rb') as f:
            params_dict = packfmt.loads(f.read())
        parameters = {k: T.as_tensor(np.array(v), device=self.device) for k, v in params_dict['parameters'].items()}
        self.load_state_dict(parameters)
        return params_dict['step'], params_dict['episode_count'], params_dict['rew_mean'], params_dict['len_mean']


class DeepQNetwork(Network):
    """dqn/network.py:50-74."""

    def __init__(self, device, lr, nn_conf_func, input_dim, output_dim, reduction='mean'):
        super().__init__(device, nn_conf_func, input_dim)
        self._n_actions, self._dueling = int(output_dim), False
        self.fc_out = nn.Linear(self.fc_out_dim, output_dim)
        self.optimizer = self.optim_func(nn.Module.parameters(self), lr=lr)
        self.loss = self.loss_func(reduction=reduction)
        self.to(self.device)


class DuelingDeepQNetwork(Network):
    """dqn/network.py:77-117."""

    def __init__(self, device, lr, nn_conf_func, input_dim, output_dim, reduction='mean'):
        super().__init__(device, nn_conf_func, input_dim)
        self._n_actions, self._dueling = int(output_dim), True
        self.fc_val = nn.Linear(self.fc_out_dim, 1)
        self.fc_adv = nn.Linear(self.fc_out_dim, output_dim)
        self.aggregate_layer = (lambda val, adv: T.add(val, (adv - adv.mean(dim=1, keepdim=True))))
        self.optimizer = self.optim_func(nn.Module.parameters(self), lr=lr)
        self.loss = self.loss_func(reduction=reduction)
        self.to(self.device)

    def _heads(self, s):
        lh = self._standalone_handle()
        self._push()
        dev = T.device("cuda", lh.device_index)
        x = T.as_tensor(s, dtype=T.float32, device=dev).contiguous().reshape(-1, self._obs_dim)
        out = T.empty(x.shape[0], self._n_actions + 1, dtype=T.float32, device=dev)
        check(lib().rmc_learner_heads(lh.handle, self._kind, ptr(x), x.shape[0], ptr(out), stream_ptr(lh.device_index)))
        return out

    def value(self, s):
        """network.py:98-102."""
        return self._heads(s)[:, :1]

    def advantages(self, s):
        """network.py:104-108."""
        return self._heads(s)[:, 1:]
