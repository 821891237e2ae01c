"""Tensors in compressible device memory (rt_obs_alloc / rt_obs_free, include/rt_b200.h).

Observation maps are mostly zeros and rollout readings are small integers stored as float32: both are
patterns Blackwell's L2 compresses on their way to HBM when the allocation is compressible (driver
virtual-memory API; ``cudaMalloc`` / the torch caching allocator never is).  Measured on B200
(profiles/r02t, r02x): rasteriser 0.138 -> 0.115 ms per 805 MB, batch moments update 0.179 -> 0.148 ms per
GiB.  The tensors behave like any CUDA tensor; the block is unmapped when the last tensor on it dies.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib

_TYPESTR = {"float32": "<f4", "float64": "<f8", "int32": "<i4", "int64": "<i8", "uint8": "|u1", "int16": "<i2"}


class _Block:
    """Owner of one rt_obs_alloc block; tensors built on it keep it alive (``__cuda_array_interface__``)."""

    def __init__(self, nbytes: int, shape, typestr: str) -> None:
        self._L = _lib.lib()
        ptr, comp = C.c_void_p(), C.c_int32(0)
        _lib.check(self._L.rt_obs_alloc(nbytes, C.byref(ptr), C.byref(comp)), "rt_obs_alloc")
        self._ptr = ptr
        self.compressible = bool(comp.value)
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr.value, False),
                                          "version": 2, "strides": None}

    def __del__(self) -> None:  # pragma: no cover - runs when the last tensor on the block dies
        try:
            if self._ptr.value:
                self._L.rt_obs_free(self._ptr)
                self._ptr = C.c_void_p()
        except Exception:
            pass


def compressible_tensor(shape, dtype=None, device=None, zero: bool = True):
    """CUDA tensor of ``dtype`` (default float32) in compressible device memory.  ``tensor.rt_compressible``
    says whether the device granted compressible backing (plain device memory otherwise — same results)."""
    torch = _lib.require_cuda()
    dtype = torch.float32 if dtype is None else dtype
    name = str(dtype).replace("torch.", "")
    _lib.require(name in _TYPESTR, f"unsupported dtype {dtype}")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    shape = tuple(int(x) for x in shape)
    n = int(np.prod(shape))
    _lib.require(n > 0, "compressible_tensor needs a non-empty shape")
    with torch.cuda.device(dev):
        try:
            block = _Block(n * np.dtype(_TYPESTR[name]).itemsize, shape, _TYPESTR[name])
        except _lib.RtError as exc:  # no virtual-memory API on this driver: ordinary device memory serves as well
            warnings.warn(f"rt_obs_alloc unavailable ({exc}); using the torch allocator", RuntimeWarning, stacklevel=2)
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)
            t.rt_compressible = False
            return t
        t = torch.as_tensor(block, device=dev)
        if zero:
            t.zero_()
    t.rt_compressible = block.compressible
    return t


def obs_tensor(shape, device=None, zero: bool = True):
    """float32 observation-map tensor ([..., 3, G, G]) in compressible memory: the target of
    ``env.step(out_maps=...)`` and of rollout observation buffers."""
    return compressible_tensor(shape, None, device, zero)


def readings_tensor(shape, device=None, zero: bool = True):
    """float32 rollout buffer of Poisson readings ([T, E, A], ``env.set_rollout``) in compressible memory."""
    return compressible_tensor(shape, None, device, zero)
