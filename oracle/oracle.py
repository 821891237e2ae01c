"""ctypes wrapper around the CPU oracle (oracle/rt_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_ref" / "librt_oracle.so"

# ids of include/rt_b200.h:rt_buf -> (dtype, per-env shape builder)
BUF = dict(
    AGENT_XY=0, PREV_DIST=1, START_XY=2, SRC_XY=3, SRC_INT=4, BKG=5, N_POLY=6, EDGES=7, TICK=8,
    EPISODE=9, VISIT=10, INTENSITY=11, WELFORD=15, WELFORD_N=16, EST_MAXMIN=17, FLAGS=18,
    LAST_COUNT=19, LAST_LOS=20, LAST_LAMBDA=21, LAST_EST=22, LAST_Z=23, LAST_VALUE=24, VISIT_LUT=25,
)


class OrcCfg(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("n_agents", C.c_int32), ("grid", C.c_int32), ("max_steps", C.c_int32),
        ("poly_min", C.c_int32), ("poly_max", C.c_int32), ("rect_min", C.c_int32), ("rect_max", C.c_int32),
        ("lognorm_step", C.c_int32), ("fixed_scenario", C.c_int32),
        ("env_id_base", C.c_int64), ("seed", C.c_uint64),
        ("cell_pitch_cm", C.c_double), ("min_dist_cm", C.c_double), ("transmission", C.c_double),
        ("src_lo", C.c_double), ("src_hi", C.c_double), ("bkg_lo", C.c_double), ("bkg_hi", C.c_double),
    ]


class OrcWelford(C.Structure):
    _fields_ = [("count", C.c_int32), ("mean", C.c_double), ("m2", C.c_double), ("var", C.c_double),
                ("std", C.c_double), ("zmax", C.c_double), ("zmin", C.c_double)]


DEFAULTS = dict(
    n_envs=1, n_agents=1, grid=32, max_steps=120, poly_min=0, poly_max=0, rect_min=2, rect_max=6,
    lognorm_step=2, fixed_scenario=0, env_id_base=0, seed=0, cell_pitch_cm=100.0, min_dist_cm=50.0,
    transmission=0.0, src_lo=1e6, src_hi=1e7, bkg_lo=10.0, bkg_hi=51.0,
)


def build(force: bool = False) -> Path:
    """Compile the oracle with oracle/Makefile (gcc) if the .so is missing or stale."""
    src = _HERE / "rt_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(OrcCfg)]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_threads.argtypes = [vp, C.c_int]
        L.orc_set_threads.restype = C.c_int
        L.orc_set_scenario.argtypes = [vp] + [vp] * 6
        L.orc_inject_uniforms.argtypes = [vp, vp, i32]
        L.orc_rasterize.argtypes = [vp, vp]
        L.orc_reset.argtypes = [vp, vp, vp, vp]
        L.orc_step.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_errors.argtypes = [vp]
        L.orc_errors.restype = i32
        L.orc_export.argtypes = [vp, C.c_int, vp]
        L.orc_export.restype = C.c_size_t
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        for f in ("orc_log", "orc_exp", "orc_loggam"):
            getattr(L, f).argtypes = [dbl]
            getattr(L, f).restype = dbl
        L.orc_poisson.argtypes = [dbl, vp, i32, C.POINTER(i32), C.POINTER(i32)]
        L.orc_poisson.restype = i64
        L.orc_uniforms.argtypes = [u64, i64, i32, i32, i32, i32, i32, vp]
        L.orc_welford_reset.argtypes = [C.POINTER(OrcWelford)]
        L.orc_welford_update.argtypes = [C.POINTER(OrcWelford), dbl]
        L.orc_welford_update.restype = C.c_int
        L.orc_welford_standardize.argtypes = [C.POINTER(OrcWelford), dbl]
        L.orc_welford_standardize.restype = dbl
        L.orc_normalize.argtypes = [dbl, dbl, C.c_int, dbl, C.POINTER(dbl)]
        L.orc_normalize.restype = C.c_int
        L.orc_lognormalize.argtypes = [dbl, dbl, i32, C.POINTER(dbl)]
        L.orc_lognormalize.restype = C.c_int
        L.orc_median.argtypes = [vp, i32]
        L.orc_median.restype = dbl
        L.orc_est_maxmin.argtypes = [dbl, C.POINTER(dbl), C.POINTER(dbl)]
        L.orc_moments_update.argtypes = [vp, vp, i64]
        L.orc_moments_merge.argtypes = [vp, vp]
        L.orc_moments_standardize.argtypes = [vp, vp, vp, i64]
        L.orc_segment_hits.argtypes = [vp, vp]
        L.orc_segment_hits.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """Batched view over E scalar oracle envs; numpy in, numpy out."""

    def __init__(self, **kw):
        cfg = dict(DEFAULTS)
        cfg.update(kw)
        self.cfg = cfg
        self.E, self.A, self.G = cfg["n_envs"], cfg["n_agents"], cfg["grid"]
        self._c = OrcCfg(**cfg)
        self._h = lib().orc_create(C.byref(self._c))
        self._inj = None

    def __del__(self):
        try:  # may run during interpreter shutdown, when module globals are already gone
            if getattr(self, "_h", None):
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_threads(self, n: int) -> int:
        return lib().orc_set_threads(self._h, n)

    def set_scenario(self, src_xy=None, src_int=None, bkg=None, start_xy=None, n_poly=None, edges=None):
        def prep(a, dt, shape):
            if a is None:
                return None
            a = np.ascontiguousarray(np.asarray(a, dtype=dt).reshape(shape))
            return a
        E, A = self.E, self.A
        args = [prep(src_xy, np.int32, (E, 2)), prep(src_int, np.float64, (E,)), prep(bkg, np.float64, (E,)),
                prep(start_xy, np.int32, (E, A, 2)), prep(n_poly, np.int32, (E,)),
                prep(edges, np.float32, (E, 20, 4))]
        lib().orc_set_scenario(self._h, *[_p(a) for a in args])

    def inject_uniforms(self, u):
        if u is None:
            self._inj = None
            lib().orc_inject_uniforms(self._h, None, 0)
            return
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(self.E, self.A, -1)
        self._inj = u
        lib().orc_inject_uniforms(self._h, _p(u), u.shape[2])

    def reset(self, mask=None, maps: bool = True):
        E, A, G = self.E, self.A, self.G
        xy = np.zeros((E, A, 2), np.int32)
        m = np.zeros((E, 3, G, G), np.float32) if maps else None
        mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().orc_reset(self._h, _p(mk), _p(xy), _p(m))
        return xy, m

    def step(self, actions, maps: bool = True):
        E, A, G = self.E, self.A, self.G
        act = np.ascontiguousarray(actions, dtype=np.int32).reshape(E, A)
        xy = np.zeros((E, A, 2), np.int32)
        rew = np.zeros((E, A), np.int32)
        done = np.zeros((E, A), np.uint8)
        m = np.zeros((E, 3, G, G), np.float32) if maps else None
        lib().orc_step(self._h, _p(act), _p(xy), _p(m), _p(rew), _p(done))
        return xy, m, rew, done

    def step_into(self, act, xy, maps, rew, done):
        """No-allocation step for timing (arrays preallocated by the caller)."""
        lib().orc_step(self._h, _p(act), _p(xy), _p(maps), _p(rew), _p(done))

    def rasterize(self):
        m = np.zeros((self.E, 3, self.G, self.G), np.float32)
        lib().orc_rasterize(self._h, _p(m))
        return m

    def errors(self) -> int:
        return lib().orc_errors(self._h)

    def export(self, name: str) -> np.ndarray:
        E, A, G = self.E, self.A, self.G
        cap = self.cfg["max_steps"] * A
        spec = {
            "AGENT_XY": (np.int32, (E, A, 2)), "PREV_DIST": (np.int32, (E, A)), "START_XY": (np.int32, (E, A, 2)),
            "SRC_XY": (np.int32, (E, 2)), "SRC_INT": (np.float64, (E,)), "BKG": (np.float64, (E,)),
            "N_POLY": (np.int32, (E,)), "EDGES": (np.float32, (E, 20, 4)), "TICK": (np.int32, (E,)),
            "EPISODE": (np.int32, (E,)), "VISIT": (np.uint16, (E, G * G)), "INTENSITY": (np.float32, (E, G * G)),
            "WELFORD": (np.float64, (E, 6)), "WELFORD_N": (np.int32, (E,)), "EST_MAXMIN": (np.float64, (E, 2)),
            "FLAGS": (np.int32, (E,)), "LAST_COUNT": (np.int32, (E, A)), "LAST_LOS": (np.uint8, (E, A)),
            "LAST_LAMBDA": (np.float64, (E, A)), "LAST_EST": (np.float64, (E, A)), "LAST_Z": (np.float64, (E, A)),
            "LAST_VALUE": (np.float64, (E, A)), "VISIT_LUT": (np.float32, (cap + 1,)),
        }[name]
        out = np.zeros(spec[1], spec[0])
        n = lib().orc_export(self._h, BUF[name], _p(out))
        assert n == out.nbytes, (name, n, out.nbytes)
        return out


# ---- scalar helpers -------------------------------------------------------------------------
def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def uniforms(seed, gid, stream, agent, tick, episode, n):
    out = np.zeros(n, np.float64)
    lib().orc_uniforms(seed, gid, stream, agent, tick, episode, n, _p(out))
    return out


def poisson(lam, u):
    u = np.ascontiguousarray(u, np.float64)
    used, flags = C.c_int32(0), C.c_int32(0)
    r = lib().orc_poisson(float(lam), _p(u), len(u), C.byref(used), C.byref(flags))
    return int(r), used.value, flags.value


class Welford:
    def __init__(self):
        self.s = OrcWelford()
        lib().orc_welford_reset(C.byref(self.s))

    def update(self, r):
        return lib().orc_welford_update(C.byref(self.s), float(r))

    def standardize(self, r):
        return lib().orc_welford_standardize(C.byref(self.s), float(r))


def normalize(cur, mx, mn=None):
    out = C.c_double()
    rc = lib().orc_normalize(float(cur), float(mx), 0 if mn is None else 1, 0.0 if mn is None else float(mn),
                             C.byref(out))
    return rc, out.value


def lognormalize(cur, mx, step=2):
    out = C.c_double()
    rc = lib().orc_lognormalize(float(cur), float(mx), int(step), C.byref(out))
    return rc, out.value


def median(v):
    v = np.ascontiguousarray(v, np.float64)
    return lib().orc_median(_p(v), len(v))


def est_maxmin(est, mx, mn):
    a, b = C.c_double(mx), C.c_double(mn)
    lib().orc_est_maxmin(float(est), C.byref(a), C.byref(b))
    return a.value, b.value


def moments_update(state, x):
    s = np.ascontiguousarray(state, np.float64).copy()
    x = np.ascontiguousarray(x, np.float32)
    lib().orc_moments_update(_p(s), _p(x), x.size)
    return s


def moments_merge(a, b):
    a = np.ascontiguousarray(a, np.float64).copy()
    b = np.ascontiguousarray(b, np.float64)
    lib().orc_moments_merge(_p(a), _p(b))
    return a


def moments_standardize(state, x):
    s = np.ascontiguousarray(state, np.float64)
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(x)
    lib().orc_moments_standardize(_p(s), _p(x), _p(out), x.size)
    return out


def rect_blocks(segs, rect):
    """uint8[n]: does segment i = (px, py, qx, qy) hit at least one edge of the rectangle (x0, y0, x1, y1)?"""
    s = np.ascontiguousarray(segs, np.float32).reshape(-1, 4)
    r = np.ascontiguousarray(rect, np.float32)
    out = np.zeros(s.shape[0], np.uint8)
    f = lib().orc_rect_blocks_batch
    f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    f.restype = None
    f(_p(s), s.shape[0], _p(r), _p(out))
    return out


def segment_hits(seg, edge):
    s = np.ascontiguousarray(seg, np.float32)
    e = np.ascontiguousarray(edge, np.float32)
    return bool(lib().orc_segment_hits(_p(s), _p(e)))
