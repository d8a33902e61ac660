"""On-disk ``.pack`` checkpoint codec.

The reference writes checkpoints with msgpack + the msgpack-numpy convention
(dqn/network.py:27-47 through dqn/utils/msgpack_numpy.py:74-130): every ndarray becomes a map
with *bytes* keys ``nd, type, kind, shape, data`` (in that order), numpy scalars a map
``nd=False, type, data``; strings/bytes use the modern ``use_bin_type`` split.  This module
re-states that wire format without monkey-patching the global ``msgpack`` module, so files
written here load in the reference and vice versa, byte for byte (tests/test_checkpoint.py).
"""
from __future__ import annotations

import msgpack
import numpy as np


def _to_wire(obj):
    if isinstance(obj, np.ndarray):
        if obj.dtype.kind == "V":
            raise TypeError("structured arrays are not part of the checkpoint format")
        buf = obj if obj.flags["C_CONTIGUOUS"] else np.ascontiguousarray(obj)
        return {b"nd": True, b"type": obj.dtype.str, b"kind": b"", b"shape": obj.shape, b"data": buf.tobytes()}
    if isinstance(obj, (np.bool_, np.number)):
        return {b"nd": False, b"type": obj.dtype.str, b"data": obj.tobytes()}
    if isinstance(obj, complex):
        return {b"complex": True, b"data": repr(obj)}
    raise TypeError("cannot serialise %r" % type(obj))


def _from_wire(obj):
    if b"nd" in obj:
        if obj[b"nd"] is True:
            return np.frombuffer(obj[b"data"], dtype=np.dtype(obj[b"type"])).reshape(obj[b"shape"])
        return np.frombuffer(obj[b"data"], dtype=np.dtype(obj[b"type"]))[0]
    if b"complex" in obj:
        return complex(obj[b"data"])
    return obj


def dumps(obj) -> bytes:
    return msgpack.packb(obj, default=_to_wire, use_bin_type=True)


def loads(data: bytes):
    return msgpack.unpackb(data, object_hook=_from_wire, raw=False, strict_map_key=False)
