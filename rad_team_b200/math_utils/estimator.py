"""IntensityResamplingEstimator on the GPU (reference: src/math_utils/estimator.py:6-109).

The reference keys a dict by arbitrary (x, y) tuples; the device structure has a fixed number of key
slots, so the facade maps each new key to the next free slot (host dict) and the kernels keep, per slot,
a value-sorted linked list of readings (median = walk to the middle).  (The env step kernel keeps its per-cell
readings in 32-byte sorted bucket chains instead — rt_env_kernels.cuh: bucket_insert_median; this stand-alone
object is a parity surface for the reference's class, not a hot path.)"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

from .. import _lib
from .standardize import _view


class BatchedEstimator:
    """n independent estimators, n_keys key slots and `capacity` readings each."""

    def __init__(self, n: int, n_keys: int, capacity: int, device=None) -> None:
        torch = _lib.require_cuda()
        self._torch = torch
        self.n, self.n_keys, self.capacity = int(n), int(n_keys), int(capacity)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._L = _lib.lib()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.rt_estimator_create(self.n, self.n_keys, self.capacity, C.byref(self._h)))
        self._err = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        p = C.c_void_p()
        _lib.check(self._L.rt_estimator_maxmin(self._h, C.byref(p)))
        self.maxmin = _view(torch, p.value, (self.n, 2), "<f8", self.device)

    def __del__(self):  # pragma: no cover
        try:
            if self._h.value:
                self._L.rt_estimator_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _stream(self):
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def _keys(self, key):
        t = self._torch
        k = t.as_tensor(key, dtype=t.int32, device=self.device).contiguous()
        _lib.require(k.numel() == self.n, "one key per estimator expected")
        return k

    def reset(self) -> None:
        _lib.check(self._L.rt_estimator_reset(self._h, self._stream()))

    def update(self, key, value):
        t = self._torch
        k = self._keys(key)
        v = t.as_tensor(value, dtype=t.float64, device=self.device).contiguous()
        est = t.empty(self.n, dtype=t.float64, device=self.device)
        _lib.check(self._L.rt_estimator_update(self._h, k.data_ptr(), v.data_ptr(), est.data_ptr(),
                                               self._err.data_ptr(), self._stream()))
        return est, self._err

    def estimate(self, key):
        t = self._torch
        k = self._keys(key)
        est = t.empty(self.n, dtype=t.float64, device=self.device)
        _lib.check(self._L.rt_estimator_estimate(self._h, k.data_ptr(), est.data_ptr(), self._err.data_ptr(),
                                                 self._stream()))
        return est, self._err

    def buffer(self, key, max_out: int):
        t = self._torch
        k = self._keys(key)
        out = t.zeros((self.n, max_out), dtype=t.float64, device=self.device)
        n_out = t.zeros(self.n, dtype=t.int32, device=self.device)
        _lib.check(self._L.rt_estimator_buffer(self._h, k.data_ptr(), out.data_ptr(), n_out.data_ptr(), max_out,
                                               self._err.data_ptr(), self._stream()))
        return out, n_out, self._err


class _ReadingsView(dict):
    """`estimator.readings` as the reference exposes it: {key: [values in insertion order]}."""


class IntensityResamplingEstimator:
    """Scalar facade with the reference's method names (estimator.py:25-95)."""

    def __init__(self, readings: Optional[Dict[Tuple[int, int], List[float]]] = None, _min: float = 0.0,
                 _max: float = 0.0, n_keys: int = 4096, capacity: int = 4096) -> None:
        """`readings`, `_min`, `_max`: the reference dataclass's init fields (estimator.py:17-23; there is no
        __post_init__, so a passed table seeds the estimator and the passed extrema stand).  `n_keys` / `capacity` size
        the device pools (the reference's dict and lists are unbounded; exhaustion raises MemoryError naming the limit)."""
        self._b = BatchedEstimator(1, n_keys, capacity)
        self._slot: Dict[Tuple[int, int], int] = {}
        self._n_readings = 0
        if readings:
            for key, values in readings.items():
                for v in values:
                    self.update(key, v)
        if readings or _min != 0.0 or _max != 0.0:  # seeding does not move the extrema in the reference
            self._b.maxmin[0, 0] = float(_max)
            self._b.maxmin[0, 1] = float(_min)

    # reference attributes
    @property
    def readings(self) -> Dict[Tuple[int, int], List[float]]:
        return _ReadingsView({k: self.get_buffer(k) for k in self._slot})

    @property
    def _max(self) -> float:
        return float(self._b.maxmin[0, 0].item())

    @property
    def _min(self) -> float:
        return float(self._b.maxmin[0, 1].item())

    def reset(self) -> None:
        self._b.reset()
        self._slot = {}
        self._n_readings = 0

    def update(self, key: Tuple[int, int], value: float) -> None:
        if key not in self._slot:
            if len(self._slot) >= self._b.n_keys:
                raise MemoryError(f"estimator key slots exhausted (n_keys={self._b.n_keys}; the reference's dict is "
                                  "unbounded — construct with a larger n_keys)")
            self._slot[key] = len(self._slot)
        if self._n_readings >= self._b.capacity:  # the device log is a fixed pool, the reference's lists are unbounded
            raise MemoryError(f"estimator reading capacity exhausted (capacity={self._b.capacity} readings in total "
                              "over all keys — construct with a larger capacity)")
        self._n_readings += 1
        _, err = self._b.update([self._slot[key]], [float(value)])
        _lib.check(int(err[0].item()), "estimator update")

    def get_buffer(self, key: Tuple[int, int]) -> List:
        if not self.check_key(key=key):
            raise ValueError("Key does not exist")
        out, n_out, err = self._b.buffer([self._slot[key]], self._b.capacity)
        _lib.check(int(err[0].item()), "Key does not exist")
        vals = out[0, : int(n_out[0].item())].tolist()
        return [int(v) if float(v).is_integer() else v for v in vals]

    def get_estimate(self, key: Tuple[int, int]) -> float:
        if not self.check_key(key=key):
            raise ValueError("Key does not exist")
        est, err = self._b.estimate([self._slot[key]])
        _lib.check(int(err[0].item()), "Key does not exist")
        return float(est[0].item())

    def get_max(self) -> float:
        return self._max

    def get_min(self) -> float:
        return self._min

    def check_key(self, key: Tuple[int, int]) -> bool:
        return True if key in self._slot else False
