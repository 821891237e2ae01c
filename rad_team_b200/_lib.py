"""ctypes binding of librt_b200.so (include/rt_b200.h).  No torch types cross this boundary:
tensors are passed as `tensor.data_ptr()` and the stream as `torch.cuda.current_stream().cuda_stream`.

There is deliberately no CPU fallback.  If the shared library is missing it is rebuilt with nvcc
(rad_team_b200/build.py); if that is impossible the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import build as _build

# --- status codes / flags (include/rt_b200.h) -------------------------------------------------
RT_OK, RT_ERR_BAD_ARG, RT_ERR_CUDA, RT_ERR_NOMEM, RT_ERR_BAD_ACTION, RT_ERR_ASSERT, RT_ERR_KEY = range(7)
FLAG_BAD_ACTION, FLAG_LOG_FULL, FLAG_LAMBDA, FLAG_DRAWS, FLAG_NORMALIZE = 0x01, 0x02, 0x04, 0x08, 0x10
MAX_AGENTS, MAX_POLYS, MAX_EDGES = 8, 5, 20
COMM_MAX_RANKS, COMM_HANDLE_BYTES = 16, 64

BUF = dict(
    AGENT_XY=0, PREV_DIST=1, START_XY=2, SRC_XY=3, SRC_INT=4, BKG=5, N_POLY=6, EDGES=7, TICK=8, EPISODE=9,
    VISIT=10, INTENSITY=11, CELLREC=12, LOG=13, WELFORD=15, WELFORD_N=16, EST_MAXMIN=17, FLAGS=18,
    LAST_COUNT=19, LAST_LOS=20, LAST_LAMBDA=21, LAST_EST=22, LAST_Z=23, LAST_VALUE=24, VISIT_LUT=25,
    EP_RETURN=26, EP_DONE=27, VREC=29,
)


class EnvCfg(C.Structure):
    """rt_env_cfg"""
    _fields_ = [
        ("n_envs", C.c_int32), ("n_agents", C.c_int32), ("grid", C.c_int32), ("max_steps", C.c_int32),
        ("poly_min", C.c_int32), ("poly_max", C.c_int32), ("rect_min", C.c_int32), ("rect_max", C.c_int32),
        ("lognorm_step", C.c_int32), ("fixed_scenario", C.c_int32),
        ("env_id_base", C.c_int64), ("seed", C.c_uint64),
        ("cell_pitch_cm", C.c_double), ("min_dist_cm", C.c_double), ("transmission", C.c_double),
        ("src_lo", C.c_double), ("src_hi", C.c_double), ("bkg_lo", C.c_double), ("bkg_hi", C.c_double),
    ]


# every symbol include/rt_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "rt_abi_version": (C.c_int, []),
    "rt_status_string": (C.c_char_p, [C.c_int]),
    "rt_env_create": (C.c_int, [C.POINTER(EnvCfg), _pp]),
    "rt_env_destroy": (C.c_int, [_vp]),
    "rt_env_set_scenario": (C.c_int, [_vp] * 7),
    "rt_env_reset": (C.c_int, [_vp] * 5),
    "rt_env_step": (C.c_int, [_vp] * 7),
    "rt_env_reset_host": (C.c_int, [_vp] * 5),
    "rt_env_step_host": (C.c_int, [_vp] * 7),
    "rt_env_rasterize": (C.c_int, [_vp, _vp, _vp]),
    "rt_env_set_incremental": (C.c_int, [_vp, _i32]),
    "rt_env_inject_uniforms": (C.c_int, [_vp, _vp, _i32]),
    "rt_env_set_rollout": (C.c_int, [_vp, _vp]),
    "rt_env_episode_stats": (C.c_int, [_vp, _vp, _vp]),
    "rt_env_state_bytes": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "rt_env_get_state": (C.c_int, [_vp, _vp, _vp]),
    "rt_env_set_state": (C.c_int, [_vp, _vp, _vp]),
    "rt_env_buffer": (C.c_int, [_vp, C.c_int, _pp, C.POINTER(C.c_size_t)]),
    "rt_env_errors": (C.c_int, [_vp, C.POINTER(_i32), _vp]),
    "rt_env_launch_count": (_i64, [_vp]),
    "rt_obs_alloc": (C.c_int, [C.c_size_t, _pp, C.POINTER(_i32)]),
    "rt_obs_free": (C.c_int, [_vp]),
    "rt_welford_create": (C.c_int, [_i64, _pp]),
    "rt_welford_destroy": (C.c_int, [_vp]),
    "rt_welford_reset": (C.c_int, [_vp, _vp]),
    "rt_welford_update": (C.c_int, [_vp, _vp, _vp, _vp]),
    "rt_welford_standardize": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "rt_welford_buffers": (C.c_int, [_vp, _pp, _pp]),
    "rt_moments_create": (C.c_int, [_pp]),
    "rt_moments_destroy": (C.c_int, [_vp]),
    "rt_moments_reset": (C.c_int, [_vp, _vp]),
    "rt_moments_update": (C.c_int, [_vp, _vp, _i64, _vp]),
    "rt_moments_standardize": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "rt_moments_state": (C.c_int, [_vp, _pp]),
    "rt_moments_merge": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "rt_comm_create": (C.c_int, [_i32, _i32, _pp, _vp]),
    "rt_comm_connect": (C.c_int, [_vp, _vp]),
    "rt_comm_create_shm": (C.c_int, [C.c_char_p, _i32, _i32, _pp]),
    "rt_comm_disconnect": (C.c_int, [_vp]),
    "rt_comm_destroy": (C.c_int, [_vp]),
    "rt_comm_errors": (C.c_int, [_vp, C.POINTER(_i32), _vp]),
    "rt_stats_allgather_merge": (C.c_int, [_vp] * 7),
    "rt_stats_pack": (C.c_int, [_vp] * 4),
    "rt_stats_merge_gathered": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "rt_normalize": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "rt_lognormalize": (C.c_int, [_vp, _dbl, _i32, _vp, _vp, _i64, _vp]),
    "rt_estimator_create": (C.c_int, [_i64, _i32, _i32, _pp]),
    "rt_estimator_destroy": (C.c_int, [_vp]),
    "rt_estimator_reset": (C.c_int, [_vp, _vp]),
    "rt_estimator_update": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "rt_estimator_estimate": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "rt_estimator_buffer": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "rt_estimator_maxmin": (C.c_int, [_vp, _pp]),
}

_lib = None


def lib_path() -> Path:
    return _build.LIB


def lib() -> C.CDLL:
    """Load (building first if stale or missing) librt_b200.so.  Raises if that is impossible."""
    global _lib
    if _lib is None:
        path = _build.LIB
        alt = os.environ.get("RT_B200_LIB")  # A/B builds of the SAME library (tools/, profiles/ sweeps), never a fallback
        if alt:
            path = Path(alt)
        elif not path.exists():
            _build.build()  # raises RuntimeError without nvcc: no silent fallback
        try:
            L = C.CDLL(str(path))
        except OSError as e:  # pragma: no cover
            raise ImportError(f"rad_team_b200: cannot load {path}: {e}") from e
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the ABI drifted
            fn.restype = res
            fn.argtypes = args
        if L.rt_abi_version() != 1:
            raise ImportError("rad_team_b200: ABI version mismatch")
        _lib = L
    return _lib


class RtError(RuntimeError):
    pass


def status_string(rc: int) -> str:
    return lib().rt_status_string(rc).decode()


def check(rc: int, what: str = "") -> None:
    """Map rt_status to the exception type the reference raises for the same condition."""
    if rc == RT_OK:
        return
    msg = f"{what}: {status_string(rc)}" if what else status_string(rc)
    if rc in (RT_ERR_BAD_ACTION, RT_ERR_KEY, RT_ERR_BAD_ARG):
        raise ValueError(msg)
    if rc == RT_ERR_ASSERT:
        raise AssertionError(msg)
    if rc == RT_ERR_NOMEM:
        raise MemoryError(msg)
    raise RtError(msg)


def require(cond: bool, what: str) -> None:
    """Argument validation in front of raw-pointer calls: a real exception, not an `assert` (python -O keeps it)."""
    if not cond:
        raise ValueError(what)


def require_cuda():
    """The product has no CPU path: fail loudly when there is no device."""
    import torch

    if not torch.cuda.is_available():
        raise RtError("rad_team_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch
