"""ctypes binding of librmc_b200.so (C ABI in include/rmc_b200.h) + the nvcc build recipe.

There is deliberately no CPU fallback: ``lib()`` raises if the shared library is missing,
and every compute entry of the library itself fails without an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "librmc_b200.so")
HEADER = os.path.join(REPO_ROOT, "include", "rmc_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

# enums of include/rmc_b200.h
ONLINE, TARGET, ADAM_M, ADAM_V, GRADS = 0, 1, 2, 3, 4
PH_SAMPLE, PH_FORWARD, PH_PRIORITY, PH_BACKWARD, PH_ADAM, PH_POLYAK, PH_HARDSYNC = 1, 2, 4, 8, 16, 32, 64
PREC_FP32, PREC_BF16_TC = 0, 1
ACT_RELU, ACT_ELU = 0, 1
PH_LEARN = PH_SAMPLE | PH_FORWARD | PH_PRIORITY | PH_BACKWARD | PH_ADAM


class NetSpec(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("hidden1", C.c_int32), ("hidden2", C.c_int32), ("n_actions", C.c_int32),
                ("dueling", C.c_int32), ("double_dqn", C.c_int32), ("prioritized", C.c_int32),
                ("activation", C.c_int32)]


class HybridSpec(C.Structure):
    """rmc_hybrid_spec_t: the repo-HEAD two-stream network (env/dqn_config.py:66-193)."""
    _fields_ = [("macro_len", C.c_int32), ("grid_c", C.c_int32), ("grid_h", C.c_int32), ("grid_w", C.c_int32),
                ("n_conv", C.c_int32), ("conv_out", C.c_int32 * 4), ("conv_sh", C.c_int32 * 4), ("conv_sw", C.c_int32 * 4),
                ("n_dense", C.c_int32), ("dense_out", C.c_int32 * 3),
                ("n_actions", C.c_int32), ("dueling", C.c_int32), ("double_dqn", C.c_int32), ("prioritized", C.c_int32),
                ("activation", C.c_int32)]


class Hyper(C.Structure):
    _fields_ = [("lr", C.c_double), ("adam_beta1", C.c_double), ("adam_beta2", C.c_double), ("adam_eps", C.c_double),
                ("gamma", C.c_double), ("polyak_k", C.c_double), ("per_eps", C.c_double), ("per_alpha", C.c_double),
                ("per_pmax", C.c_double)]


class ReplayStats(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("size", C.c_int64), ("data_pointer", C.c_int64),
                ("total_priority", C.c_double), ("max_priority", C.c_double), ("min_priority", C.c_double),
                ("rejected_nodes", C.c_int64)]


class StepArgs(C.Structure):
    _fields_ = [("batch", C.c_int64), ("phases", C.c_int32), ("precision", C.c_int32), ("per_beta", C.c_double),
                ("u_dev", C.c_void_p), ("idx_dev", C.c_void_p), ("seed", C.c_uint64), ("counter", C.c_uint64),
                ("adam_t", C.c_int64), ("grads_in_dev", C.c_void_p), ("shard_offset", C.c_int64),
                ("global_batch", C.c_int64)]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [HEADER]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/rmc_b200.cu for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH,
                                                                            os.path.join(CSRC, "rmc_b200.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


def declared_symbols():
    """Every function name declared in include/rmc_b200.h."""
    with open(HEADER) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rmc_[a-z0-9_]+)\s*\(", text)))


_lock = threading.Lock()
_lib = None

_i32, _i64, _u64, _f32, _f64, _vp, _cp = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double, C.c_void_p, C.c_char_p
_SIGS = {
    "rmc_abi_version": (_i32, []),
    "rmc_last_error": (_cp, []),
    "rmc_launch_count": (_i64, []),
    "rmc_replay_create": (_i32, [C.POINTER(_vp), _i64, _i32, _i32, _i32]),
    "rmc_replay_destroy": (_i32, [_vp]),
    "rmc_replay_push": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "rmc_replay_push_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "rmc_replay_set_priorities": (_i32, [_vp, _vp, _i64, _vp]),
    "rmc_replay_stats_sync": (_i32, [_vp, C.POINTER(ReplayStats), _vp]),
    "rmc_replay_read_tree_sync": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "rmc_replay_read_rows_sync": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "rmc_replay_row_floats": (_i32, [_vp]),
    "rmc_replay_load_host": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "rmc_per_sample": (_i32, [_vp, _i64, _f64, _vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "rmc_uniform_sample": (_i32, [_vp, _i64, _vp, _u64, _u64, _vp, _vp, _vp]),
    "rmc_tree_get_leaf": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "rmc_per_update": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "rmc_per_update_from_td": (_i32, [_vp, _vp, _vp, _i64, _f32, _f32, _f32, _vp, _vp]),
    "rmc_learner_create": (_i32, [C.POINTER(_vp), C.POINTER(NetSpec), C.POINTER(Hyper), _i64, _i32]),
    "rmc_learner_create_hybrid": (_i32, [C.POINTER(_vp), C.POINTER(HybridSpec), C.POINTER(Hyper), _i64, _i32]),
    "rmc_learner_destroy": (_i32, [_vp]),
    "rmc_learner_param_count": (_i64, [_vp]),
    "rmc_learner_set_params": (_i32, [_vp, _i32, _vp, _i64, _i32, _vp]),
    "rmc_learner_get_params": (_i32, [_vp, _i32, _vp, _i64, _i32, _vp]),
    "rmc_learner_set_hyper": (_i32, [_vp, C.POINTER(Hyper)]),
    "rmc_learner_step": (_i32, [_vp, _vp, C.POINTER(StepArgs), _vp]),
    "rmc_learner_step_push": (_i32, [_vp, _vp, C.POINTER(StepArgs), _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "rmc_learner_output": (_i32, [_vp, _cp, C.POINTER(_vp), C.POINTER(_i64)]),
    "rmc_learner_loss_sync": (_i32, [_vp, C.POINTER(_f32), _vp]),
    "rmc_learner_status": (_i32, [_vp, C.POINTER(C.c_uint32)]),
    "rmc_learner_q_values": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp]),
    "rmc_learner_heads": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp]),
    "rmc_learner_act": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "rmc_learner_act_tc": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "rmc_learner_heads_tc": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "rmc_learner_act_host_sync": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "rmc_learner_act_eps_host_sync": (_i32, [_vp, _vp, _i64, _vp, _f32, _u64, _u64, _vp]),
    "rmc_learner_debug_timing": (_i32, [_vp, _i32]),
    "rmc_debug_spans": (_i32, [_i32, _i32]),
    "rmc_debug_spans_read_sync": (_i32, [_i32, _vp, _vp, _i32]),
    "rmc_learner_debug_gaps_sync": (_i32, [_vp, _vp, _vp]),
    "rmc_learner_debug_read_sync": (_i32, [_vp, _vp, _i32, C.POINTER(_i32), _vp]),
    "rmc_group_create": (_i32, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i32]),
    "rmc_group_destroy": (_i32, [_vp]),
    "rmc_group_step": (_i32, [_vp, C.POINTER(StepArgs), _vp]),
    "rmc_group_push_host": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i64, _vp]),
    "rmc_comm_create": (_i32, [C.POINTER(_vp), _vp, _i32, _i32, _i64]),
    "rmc_comm_export": (_i32, [_vp, _vp, C.POINTER(_vp)]),
    "rmc_comm_connect": (_i32, [_vp, _vp, C.POINTER(_vp)]),
    "rmc_comm_destroy": (_i32, [_vp]),
    "rmc_comm_status_sync": (_i32, [_vp, C.POINTER(C.c_uint32), _vp]),
    "rmc_learner_step_sharded": (_i32, [_vp, _vp, _vp, C.POINTER(StepArgs), _i32, _vp]),
}
ABI_VERSION = 3


def lib() -> C.CDLL:
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "librmc_b200.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'`"
                    " -- this package has no CPU fallback" % LIB_PATH)
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
            if handle.rmc_abi_version() != ABI_VERSION:
                raise RuntimeError("librmc_b200.so ABI version mismatch; rebuild")
            _lib = handle
    return _lib


class RmcError(RuntimeError):
    pass


def check(code: int) -> None:
    if code != 0:
        raise RmcError("librmc_b200 error %d: %s" % (code, lib().rmc_last_error().decode(errors="replace")))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("multimodal_drl_rmc_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
    return torch


def ptr(t) -> int:
    """data pointer of a torch tensor / 0 for None."""
    return 0 if t is None else t.data_ptr()


_raw_stream = None


def stream_ptr(device_index=None) -> int:
    """cudaStream_t of torch's current stream on the current (or given) device."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device() if device_index is None else device_index)
    return torch.cuda.current_stream().cuda_stream
