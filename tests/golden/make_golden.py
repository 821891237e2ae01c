#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference (mounted read-only at
/root/reference) in the build container.  The reference cannot travel to the GPU box, so its outputs
are committed as fixtures next to this script.

    python tests/golden/make_golden.py

Floats are stored as float.hex() strings so that comparisons are bit-exact.

What is pinned
  ref_math_utils.json  WelfordsOnlineStandardization, Normalizer, BensLogNormalizer and
                       IntensityResamplingEstimator driven with seeded inputs (src/math_utils/*.py)
  ref_simplegrid.json  SimpleGrid random-action episodes (src/envs/simple_gridworld.py), imported
                       behind stub gymnasium / pygame modules (neither is installed) and np.NINF
  ref_pipeline.json    the composed observation pipeline: (cell, count) streams produced by the CPU
                       oracle's physics, pushed through the reference's estimator -> Welford ->
                       Normalizer objects and BensLogNormalizer, step by step
"""
from __future__ import annotations

import json
import sys
import types
import warnings
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))


def hx(x) -> str:
    return float(x).hex()


def load_reference():
    """Import the reference modules without modifying them."""
    if not REF.exists():
        raise SystemExit("reference not mounted at /root/reference — fixtures can only be regenerated in the build container")
    # stubs for the two third-party imports of simple_gridworld.py that are not installed here
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
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces, "gymnasium.envs": envs,
                        "gymnasium.envs.registration": reg, "pygame": pg})
    if not hasattr(np, "NINF"):
        np.NINF = -np.inf  # removed in NumPy 2; simple_gridworld.py:106 still uses it
    sys.path.insert(0, str(REF))
    from src.envs.simple_gridworld import SimpleGrid  # noqa: E402
    from src.math_utils.estimator import IntensityResamplingEstimator  # noqa: E402
    from src.math_utils.normalize import BensLogNormalizer, Normalizer  # noqa: E402
    from src.math_utils.standardize import WelfordsOnlineStandardization  # noqa: E402

    return SimpleGrid, IntensityResamplingEstimator, BensLogNormalizer, Normalizer, WelfordsOnlineStandardization


def gen_math_utils(Est, BensLog, Norm, Welford):
    rng = np.random.default_rng(20261019)
    out = {}
    # --- Welford streams: counts-like, tiny, huge and constant readings
    streams = []
    for kind in ("poisson50", "poisson5e4", "uniform1e7", "tiny", "constant", "zeros_then_big"):
        n = 200
        if kind == "poisson50":
            xs = rng.poisson(50.0, n).astype(float)
        elif kind == "poisson5e4":
            xs = rng.poisson(5e4, n).astype(float)
        elif kind == "uniform1e7":
            xs = rng.random(n) * 1e7
        elif kind == "tiny":
            xs = rng.random(n) * 1e-3
        elif kind == "constant":
            xs = np.full(n, 1234.0)
        else:
            xs = np.concatenate([np.zeros(20), rng.random(n - 20) * 1e5])
        w = Welford()
        rows = []
        for x in xs:
            w.update(float(x))
            rows.append([hx(x), w.count, hx(w.mean), hx(w.square_dist_mean), hx(w.sample_variance), hx(w.std),
                         hx(w.get_max()), hx(w.get_min()), hx(w.standardize(float(x))), hx(w.standardize(17.0))])
        streams.append({"kind": kind, "rows": rows})
    out["welford"] = streams
    # --- Normalizer table, including every assertion branch
    cases = []
    vals = [0.0, 1.0, -1.0, 50.0, -50.0, 100.0, 0.5, -0.25, 1e-9, 3.5, -501.0, 10.0, 11.0]
    nz = Norm()
    for cur in vals:
        for mx in (100.0, 1.0, 0.0, -1.0, 3.5):
            for mn in (None, 0.0, 10.0, -10.0, -0.25, 11.0):
                try:
                    r = nz.normalize(current_value=cur, max=mx, min=mn)
                    cases.append([hx(cur), hx(mx), None if mn is None else hx(mn), hx(r)])
                except AssertionError:
                    cases.append([hx(cur), hx(mx), None if mn is None else hx(mn), "assert"])
    for _ in range(300):  # random z-score-like triples min <= cur <= max
        lo, hi = sorted(rng.normal(0, 2, 2))
        lo, hi = min(lo, 0.0), max(hi, 0.0)
        cur = rng.uniform(lo, hi)
        r = nz.normalize(current_value=float(cur), max=float(hi), min=float(lo))
        cases.append([hx(cur), hx(hi), hx(lo), hx(r)])
    out["normalizer"] = cases
    # --- BensLogNormalizer: whole reachable table for the bench config (max = 120*3) and small ones
    logs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for mx, step in ((360, 2), (120, 2), (10, 2), (65534, 2), (7, 3)):
            b = BensLog()
            vs = sorted(set(list(range(0, min(mx * step - step, 800) + 1)) + [mx * step - step]))
            rows = []
            for v in vs:
                try:
                    rows.append([v, hx(b.normalize_incremental_logscale(current_value=v, max=mx, step_size=step))])
                except AssertionError:
                    rows.append([v, "assert"])
            for v in (mx * step - step + 1, mx * step + 5, -1):
                try:
                    rows.append([v, hx(b.normalize_incremental_logscale(current_value=v, max=mx, step_size=step))])
                except AssertionError:
                    rows.append([v, "assert"])
            logs.append({"max": mx, "step": step, "rows": rows})
    out["benslog"] = logs
    # --- estimator: random keys from a small pool so buffers grow; counts-like and fractional values
    est_runs = []
    for vals_kind in ("counts", "floats"):
        e = Est()
        rows = []
        for i in range(400):
            key = (int(rng.integers(0, 4)), int(rng.integers(0, 3)))
            v = float(rng.poisson(40.0)) if vals_kind == "counts" else float(rng.random() * 100.0)
            if i in (50, 51):
                v = 0.0  # exercises the 0-as-unset sentinel (estimator.py:45-48)
            e.update(key=key, value=v)
            rows.append([list(key), hx(v), hx(e.get_estimate(key)), hx(e.get_max()), hx(e.get_min()),
                         len(e.get_buffer(key))])
        final = {f"{k[0]},{k[1]}": [hx(v) for v in buf] for k, buf in e.readings.items()}
        est_runs.append({"kind": vals_kind, "rows": rows, "buffers": final})
    out["estimator"] = est_runs
    return out


def gen_simplegrid(SimpleGrid):
    rng = np.random.default_rng(7)
    episodes = []
    setups = [((5, 5), (0, 0), 10), ((5, 5), (6, 6), 10), ((0, 0), (9, 9), 10), ((3, 7), (3, 2), 8),
              ((15, 16), (2, 30), 32), ((1, 1), (0, 0), 4)]
    for start, terminal, size in setups:
        for rep in range(2):
            env = SimpleGrid(np.array(start, dtype=np.int32), np.array(terminal, dtype=np.int32), size=size)
            obs, info = env.reset()
            acts = rng.integers(1, 5, size=120)
            rows = []
            for a in acts:
                obs, rew, done, trunc, info = env.step(int(a))
                rows.append([int(a), int(obs[0]), int(obs[1]), int(rew), bool(done), bool(trunc),
                             float(info["distance"]), float(env.agent_previous_dist)])
            episodes.append({"start": list(start), "terminal": list(terminal), "size": size, "rows": rows})
    return {"episodes": episodes}


def gen_pipeline(Est, BensLog, Norm, Welford):
    """Oracle physics -> reference math objects.  Pins the oracle's composed statistics."""
    from oracle import oracle as O

    runs = []
    for cfg in (dict(n_envs=6, n_agents=3, grid=32, max_steps=120, poly_min=1, poly_max=5, seed=11, transmission=0.0),
                dict(n_envs=4, n_agents=1, grid=10, max_steps=60, seed=5),
                dict(n_envs=3, n_agents=2, grid=6, max_steps=200, poly_min=0, poly_max=2, rect_min=1, rect_max=2,
                     seed=3, transmission=0.25)):
        E, A, G, T = cfg["n_envs"], cfg["n_agents"], cfg["grid"], cfg["max_steps"]
        env = O.OracleEnv(**cfg)
        env.reset()
        rng = np.random.default_rng(cfg["seed"] + 100)
        est = [Est() for _ in range(E)]
        wel = [Welford() for _ in range(E)]
        nz, bl = Norm(), BensLog()
        visits = np.zeros((E, G * G), np.int64)
        rows = []
        acts_all = rng.integers(1, 5, size=(T, E, A)).astype(np.int32)
        for t in range(T):
            xy, _, rew, done = env.step(acts_all[t], maps=False)
            cnt = env.export("LAST_COUNT")
            step_rows = []
            for e in range(E):
                for a in range(A):
                    x, y = int(xy[e, a, 0]), int(xy[e, a, 1])
                    c = int(cnt[e, a])
                    est[e].update(key=(x, y), value=c)
                    m = est[e].get_estimate((x, y))
                    wel[e].update(m)
                    z = wel[e].standardize(m)
                    v = nz.normalize(current_value=z, max=wel[e].get_max(), min=wel[e].get_min())
                    visits[e, y * G + x] += 1
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        lv = bl.normalize_incremental_logscale(int(visits[e, y * G + x]), max=T * A, step_size=2)
                    step_rows.append([x, y, c, hx(m), hx(z), hx(v), hx(lv), hx(est[e].get_max()), hx(est[e].get_min())])
            rows.append(step_rows)
        runs.append({"cfg": cfg, "actions": acts_all.tolist(), "rows": rows})
    return {"runs": runs}


def main():
    SimpleGrid, Est, BensLog, Norm, Welford = load_reference()
    (HERE / "ref_math_utils.json").write_text(json.dumps(gen_math_utils(Est, BensLog, Norm, Welford)))
    (HERE / "ref_simplegrid.json").write_text(json.dumps(gen_simplegrid(SimpleGrid)))
    (HERE / "ref_pipeline.json").write_text(json.dumps(gen_pipeline(Est, BensLog, Norm, Welford)))
    for f in sorted(HERE.glob("*.json")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
