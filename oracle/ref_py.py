"""The UNMODIFIED Python reference as a timed CPU baseline (test infrastructure, like everything under oracle/).

The reference (bentotten/RAD-TEAM, mounted read-only at /root/reference in the build container) is pure Python
and does not exist on the GPU box.  `compile_reference()` — run by `make -C oracle ref_py` and by
`__graft_entry__.build()` wherever /root/reference is present — byte-compiles the four source files of the hot
path FROM WHERE THEY LIE into oracle/_ref/refpy/ (compiled outputs only; the directory is git-ignored and travels
with the repo snapshot exactly like oracle/_ref/librt_oracle.so).  No reference source is copied.

`load_reference()` imports those modules (source tree when mounted, the byte-code otherwise) behind the same
import shims tests/golden/make_golden.py uses: stub `gymnasium` / `pygame` modules (neither is installed; only
`Env.reset(seed)`, the two space constructors and `register` are touched) and `np.NINF` (removed in NumPy 2,
still used by simple_gridworld.py:106).  Nothing of the reference's arithmetic is replaced.

Timed workloads:
  * `time_simplegrid`  SimpleGrid(start=(5,5), terminal=(0,0), size=10): reset + 120 random-action steps per
                       episode (src/envs/simple_gridworld.py:114-167), one process or one process per core;
  * `time_pipeline`    the composed observation statistics per reading (src/math_utils): estimator.update ->
                       get_estimate -> Welford.update -> standardize -> Normalizer.normalize, and
                       BensLogNormalizer on the visit count.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import py_compile
import sys
import time
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
BYTECODE = HERE / "_ref" / "refpy"
FILES = ["src/envs/__init__.py", "src/envs/simple_gridworld.py", "src/math_utils/standardize.py",
         "src/math_utils/normalize.py", "src/math_utils/estimator.py"]


def _bc_path(rel: str) -> Path:
    """Byte-code file of one reference module.  Not named *.pyc: snapshot tools drop those as caches."""
    return BYTECODE / (rel[:-3].replace("/", "__") + ".bytecode")


def compile_reference(force: bool = False) -> bool:
    """Byte-compile the reference's hot-path modules into oracle/_ref/refpy/ (sourceless byte-code files).
    Returns False when the reference is not mounted (the GPU box): the prebuilt files are used as they are."""
    if not REF.exists():
        return False
    for rel in FILES:
        src, dst = REF / rel, _bc_path(rel)
        if force or not dst.exists() or dst.stat().st_mtime < src.stat().st_mtime:
            dst.parent.mkdir(parents=True, exist_ok=True)
            py_compile.compile(str(src), cfile=str(dst), doraise=True, optimize=0)
    return True


def available() -> bool:
    return REF.exists() or all(_bc_path(rel).exists() for rel in FILES)


def _import_bytecode() -> None:
    """Load the five compiled modules under their reference names (src.envs, src.envs.simple_gridworld,
    src.math_utils.*) with sourceless loaders; `src` and `src.math_utils` are namespace packages in the reference."""
    import importlib.machinery
    import importlib.util

    for ns in ("src", "src.math_utils"):
        if ns not in sys.modules:
            m = types.ModuleType(ns)
            m.__path__ = []
            sys.modules[ns] = m
    for rel in FILES:
        name = rel[:-3].replace("/", ".")
        is_pkg = name.endswith(".__init__")
        if is_pkg:
            name = name[:-len(".__init__")]
        if name in sys.modules and getattr(sys.modules[name], "__rt_bytecode__", False):
            continue
        loader = importlib.machinery.SourcelessFileLoader(name, str(_bc_path(rel)))
        spec = importlib.util.spec_from_loader(name, loader, is_package=is_pkg)
        mod = importlib.util.module_from_spec(spec)
        mod.__rt_bytecode__ = True
        sys.modules[name] = mod
        loader.exec_module(mod)
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], leaf, mod)


def _install_shims() -> None:
    gym = types.ModuleType("gymnasium")

    class Env:  # the two members SimpleGrid touches: np_random and reset(seed, options)
        _np_random = None

        @property
        def np_random(self):
            return self._np_random

        @np_random.setter
        def np_random(self, v):
            self._np_random = v

        def reset(self, seed=None, options=None):
            if seed is not None:
                self._np_random = np.random.default_rng(seed)

    gym.Env = Env
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.MultiDiscrete = lambda *a, **k: None
    spaces.Box = lambda *a, **k: None
    gym.spaces = spaces
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda *a, **k: None
    envs = types.ModuleType("gymnasium.envs")
    envs.registration = reg
    pg = types.ModuleType("pygame")
    pg.Surface = object
    pg.time = types.SimpleNamespace(Clock=object)
    for name, mod in {"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.envs": envs,
                      "gymnasium.envs.registration": reg, "pygame": pg}.items():
        sys.modules.setdefault(name, mod)
    if not hasattr(np, "NINF"):
        np.NINF = -np.inf


def load_reference():
    """(SimpleGrid, IntensityResamplingEstimator, BensLogNormalizer, Normalizer, WelfordsOnlineStandardization,
    origin) of the unmodified reference; origin is "source" (mounted tree) or "bytecode" (oracle/_ref/refpy)."""
    if not available():
        raise FileNotFoundError("reference neither mounted at /root/reference nor byte-compiled under oracle/_ref/refpy")
    _install_shims()
    origin = "source" if REF.exists() else "bytecode"
    if origin == "source":
        if str(REF) not in sys.path:
            sys.path.insert(0, str(REF))
    else:
        _import_bytecode()
    from src.envs.simple_gridworld import SimpleGrid
    from src.math_utils.estimator import IntensityResamplingEstimator
    from src.math_utils.normalize import BensLogNormalizer, Normalizer
    from src.math_utils.standardize import WelfordsOnlineStandardization

    return SimpleGrid, IntensityResamplingEstimator, BensLogNormalizer, Normalizer, WelfordsOnlineStandardization, origin


def time_simplegrid(seconds: float = 2.0):
    """env-steps/s of reset + 120-step random-action episodes on this process (fixture of tests/test_envs.py:23-25)."""
    SimpleGrid = load_reference()[0]
    env = SimpleGrid(np.array([5, 5], np.int32), np.array([0, 0], np.int32), size=10)
    acts = [int(a) for a in np.random.default_rng(0).integers(1, 5, size=120)]
    n, t0 = 0, time.perf_counter()
    while True:
        env.reset()
        for a in acts:
            env.step(a)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n * 120 / dt, n


def _worker(seconds):
    return time_simplegrid(seconds)


def time_simplegrid_all_cores(seconds: float = 2.0, procs: int | None = None):
    """One independent process per core (the reference has no vectorisation); aggregate env-steps/s."""
    procs = procs or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_worker, [seconds] * procs)
        wall = time.perf_counter() - t0
    steps = sum(int(round(r * seconds)) for r, _ in res)  # each worker ran for `seconds`
    return sum(r for r, _ in res), procs, wall, steps


def time_pipeline(seconds: float = 2.0):
    """readings/s through the reference's estimator -> Welford -> Normalizer objects and BensLogNormalizer, with
    the episode structure of the benchmark (a fresh set of objects every 360 readings = 120 steps x 3 agents)."""
    _, Est, BensLog, Norm, Welford, _ = load_reference()
    rng = np.random.default_rng(0)
    n_pre = 36000
    cells = rng.integers(0, 32, size=(n_pre, 2)).tolist()
    counts = rng.poisson(200.0, n_pre).tolist()
    n, t0 = 0, time.perf_counter()
    while True:
        for ep in range(n_pre // 360):
            est, w, nz, bl, visits = Est(), Welford(), Norm(), BensLog(), {}
            for i in range(ep * 360, ep * 360 + 360):
                k = (cells[i][0], cells[i][1])
                est.update(key=k, value=counts[i])
                m = est.get_estimate(k)
                w.update(m)
                z = w.standardize(m)
                nz.normalize(current_value=z, max=w.get_max(), min=w.get_min())
                v = visits[k] = visits.get(k, 0) + 1
                bl.normalize_incremental_logscale(v, max=360, step_size=2)
            n += 360
            if time.perf_counter() - t0 >= seconds:
                return n / (time.perf_counter() - t0), n


if __name__ == "__main__":
    print("compiled" if compile_reference(force="--force" in sys.argv) else "reference not mounted; using prebuilt byte-code")
    r1, _ = time_simplegrid(1.0)
    ra, procs, _, _ = time_simplegrid_all_cores(1.0)
    rp, _ = time_pipeline(1.0)
    print(f"{load_reference()[-1]}: SimpleGrid {r1:,.0f} env-steps/s on 1 core, {ra:,.0f} on {procs} processes; "
          f"math_utils pipeline {rp:,.0f} readings/s on 1 core")
