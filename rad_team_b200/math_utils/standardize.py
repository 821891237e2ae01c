"""Welford standardisation on the GPU (reference: src/math_utils/standardize.py:7-107).

``WelfordsOnlineStandardization`` keeps the reference's scalar interface (one stream);
``BatchedWelford`` is the same recurrence over n independent streams; ``RunningMoments`` is the
cross-environment normaliser (K4b): warp-shuffle batch moments that merge across GPUs.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from .. import _lib


class BatchedWelford:
    """n independent WelfordsOnlineStandardization streams (fp64 state, exact reference recurrence)."""

    def __init__(self, n_streams: int, device=None) -> None:
        torch = _lib.require_cuda()
        self._torch = torch
        self.n = int(n_streams)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._L = _lib.lib()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.rt_welford_create(self.n, C.byref(self._h)), "rt_welford_create")
        self._err = torch.zeros(self.n, dtype=torch.int32, device=self.device)

    def __del__(self):  # pragma: no cover
        try:
            if self._h.value:
                self._L.rt_welford_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _stream(self):
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def _as_f64(self, x):
        t = self._torch
        x = t.as_tensor(x, dtype=t.float64, device=self.device).contiguous()
        _lib.require(x.numel() == self.n, "one reading per stream expected")
        return x

    def reset(self) -> None:
        _lib.check(self._L.rt_welford_reset(self._h, self._stream()))

    def update(self, reading):
        """One reading per stream; returns the int32 status tensor (RT_ERR_ASSERT where reading < 0)."""
        r = self._as_f64(reading)
        _lib.check(self._L.rt_welford_update(self._h, r.data_ptr(), self._err.data_ptr(), self._stream()))
        return self._err

    def standardize(self, reading):
        r = self._as_f64(reading)
        z = self._torch.empty_like(r)
        _lib.check(self._L.rt_welford_standardize(self._h, r.data_ptr(), z.data_ptr(), self._err.data_ptr(),
                                                  self._stream()))
        return z, self._err

    def state(self):
        """(float64 [n, 6] = mean, M2, sample_variance, std, _max, _min ; int32 [n] count) device views."""
        t = self._torch
        sp, cp = C.c_void_p(), C.c_void_p()
        _lib.check(self._L.rt_welford_buffers(self._h, C.byref(sp), C.byref(cp)))
        return (_view(t, sp.value, (self.n, 6), "<f8", self.device), _view(t, cp.value, (self.n,), "<i4", self.device))


def _view(torch, ptr, shape, typestr, device):
    class _A:
        pass

    a = _A()
    a.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2,
                                  "strides": None}
    return torch.as_tensor(a, device=device)


class WelfordsOnlineStandardization:
    """Scalar facade with the reference's attribute and method names (standardize.py:17-107)."""

    def __init__(self, mean: float = 0.0, square_dist_mean: float = 0.0, sample_variance: float = 0.0,
                 std: float = 1.0, count: int = 0) -> None:
        # The reference is a dataclass with these init fields (standardize.py:17-26) whose __post_init__ calls
        # reset() (:33-44): whatever is passed is overwritten by the defaults.  Same here: accepted, not used.
        self._b = BatchedWelford(1)

    def _s(self):
        st, cnt = self._b.state()
        return st[0].tolist(), int(cnt[0].item())

    mean = property(lambda self: self._s()[0][0])
    square_dist_mean = property(lambda self: self._s()[0][1])
    sample_variance = property(lambda self: self._s()[0][2])
    std = property(lambda self: self._s()[0][3])
    _max = property(lambda self: self._s()[0][4])
    _min = property(lambda self: self._s()[0][5])
    count = property(lambda self: self._s()[1])

    def reset(self) -> None:
        self._b.reset()

    def update(self, reading: float) -> None:
        err = self._b.update([float(reading)])
        _lib.check(int(err[0].item()), "reading >= 0")

    def standardize(self, reading: float) -> float:
        z, err = self._b.standardize([float(reading)])
        _lib.check(int(err[0].item()), "reading >= 0")
        return float(z[0].item())

    def get_max(self) -> float:
        return self._max

    def get_min(self) -> float:
        return self._min


class RunningMoments:
    """Running (count, mean, M2) over everything fed to ``update`` — the batch form of the Welford
    standardiser for rollout-sized buffers, HBM-bound (4 B per reading), and mergeable across ranks
    (see rad_team_b200.distributed.merge_moments)."""

    def __init__(self, device=None) -> None:
        torch = _lib.require_cuda()
        self._torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._L = _lib.lib()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.rt_moments_create(C.byref(self._h)), "rt_moments_create")
        p = C.c_void_p()
        _lib.check(self._L.rt_moments_state(self._h, C.byref(p)))
        self.state = _view(torch, p.value, (3,), "<f8", self.device)  # count, mean, M2 (device view)

    def __del__(self):  # pragma: no cover
        try:
            if self._h.value:
                self._L.rt_moments_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _stream(self):
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def reset(self) -> None:
        _lib.check(self._L.rt_moments_reset(self._h, self._stream()))

    def update(self, x) -> None:
        t = self._torch
        _lib.require(x.is_cuda and x.dtype == t.float32 and x.is_contiguous(),
                     "readings must be a contiguous float32 CUDA tensor")
        _lib.check(self._L.rt_moments_update(self._h, x.data_ptr(), x.numel(), self._stream()), "rt_moments_update")

    def standardize(self, x, out=None):
        t = self._torch
        _lib.require(x.is_cuda and x.dtype == t.float32 and x.is_contiguous(),
                     "readings must be a contiguous float32 CUDA tensor")
        out = t.empty_like(x) if out is None else out
        _lib.require(out.is_cuda and out.dtype == t.float32 and out.is_contiguous() and out.numel() == x.numel(),
                     "out must be a contiguous float32 CUDA tensor of the input's size")
        _lib.check(self._L.rt_moments_standardize(self._h, x.data_ptr(), out.data_ptr(), x.numel(), self._stream()))
        return out

    def merge_gathered(self, gathered) -> None:
        """state <- Chan merge, in rank order, of gathered float64 [n_ranks, >=3] (device): row r starts
        with rank r's (count, mean, M2)."""
        t = self._torch
        _lib.require(gathered.is_cuda and gathered.dtype == t.float64 and gathered.is_contiguous()
                     and gathered.dim() == 2, "gathered must be a contiguous float64 CUDA tensor [n_ranks, >=3]")
        _lib.check(self._L.rt_moments_merge(self._h, gathered.data_ptr(), gathered.shape[0], gathered.shape[1],
                                            self._stream()))

    @property
    def count(self) -> float:
        return float(self.state[0].item())

    @property
    def mean(self) -> float:
        return float(self.state[1].item())

    @property
    def square_dist_mean(self) -> float:
        return float(self.state[2].item())
