"""Normalizer / BensLogNormalizer on the GPU (reference: src/math_utils/normalize.py:7-115).
Scalar facades keep the reference's signatures, exceptions and warnings; ``normalize_batch`` /
``lognormalize_batch`` are the elementwise tensor forms."""
from __future__ import annotations

import warnings
from typing import Optional, Union

from .. import _lib


def normalize_batch(current, max, min=None):
    """Elementwise Normalizer.normalize over float64 CUDA tensors -> (out, err int32)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    dev = current.device
    cur = current.to(torch.float64).contiguous()
    mx = max.to(torch.float64).contiguous()
    mn = None if min is None else min.to(torch.float64).contiguous()
    out = torch.empty_like(cur)
    err = torch.zeros(cur.numel(), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(L.rt_normalize(cur.data_ptr(), mx.data_ptr(), None if mn is None else mn.data_ptr(), out.data_ptr(),
                              err.data_ptr(), cur.numel(), st), "rt_normalize")
    return out, err


def lognormalize_batch(current, max: float, step_size: int = 2):
    torch = _lib.require_cuda()
    L = _lib.lib()
    cur = current.to(torch.float64).contiguous()
    out = torch.empty_like(cur)
    err = torch.zeros(cur.numel(), dtype=torch.int32, device=cur.device)
    st = torch.cuda.current_stream(cur.device).cuda_stream
    _lib.check(L.rt_lognormalize(cur.data_ptr(), float(max), int(step_size), out.data_ptr(), err.data_ptr(),
                                 cur.numel(), st), "rt_lognormalize")
    return out, err


class Normalizer:
    """Min-max normalisation to [0, 1] (normalize.py:7-55)."""

    def __init__(self, _base_max=None, _base_step_size=None) -> None:
        self._base_max = _base_max
        self._base_step_size = _base_step_size

    def normalize(self, current_value: float, max: float, min: Optional[float] = None) -> float:
        torch = _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        t = lambda v: torch.tensor([float(v)], dtype=torch.float64, device=dev)
        out, err = normalize_batch(t(current_value), t(max), None if min is None else t(min))
        _lib.check(int(err[0].item()), "Value error")
        return float(out[0].item())


class BensLogNormalizer:
    """Log-scale normalisation of an incrementing counter (normalize.py:58-115)."""

    def __init__(self, _base_max=None, _base_step_size=None) -> None:
        self._base_max: Union[float, int, None] = _base_max
        self._base_step_size: Optional[int] = _base_step_size

    def normalize_incremental_logscale(self, current_value, max, step_size: int = 2) -> float:
        torch = _lib.require_cuda()
        # host-side argument checks and warnings, in the reference's order (normalize.py:90-108)
        assert current_value >= 0 and max > 0 and step_size > 0, "Value error - input was negative that should not be"
        if not self._base_max:
            self._base_max = max
        elif self._base_max != max:
            warnings.warn(UserWarning("mismatch, Max mismatch from first use of normalize_incremental_logscale "
                                      "function! Ensure this was intentional! "))
        if not self._base_step_size:
            self._base_step_size = step_size
        if self._base_step_size != step_size:
            warnings.warn(UserWarning("mismatch, Step size mismatch from first use of "
                                      "normalize_incremental_logscale function! Ensure this was intentional"))
        if max == 1:
            raise ZeroDivisionError("float division by zero")  # math.log(x, 1) in the reference (:110)
        dev = torch.device("cuda", torch.cuda.current_device())
        cur = torch.tensor([float(current_value)], dtype=torch.float64, device=dev)
        out, err = lognormalize_batch(cur, float(max), int(step_size))
        _lib.check(int(err[0].item()), f"Normalization error, step_size: {step_size}, Current value: "
                                       f"{current_value}, Max: {max}")
        return float(out[0].item())
