#!/usr/bin/env python
"""Time the UNMODIFIED Python reference in the build container (it cannot travel to the GPU box):
SimpleGrid 120-step random-action episodes and the composed math_utils pipeline per
reading.  Context for bench.py's cpu_baseline, which times the oracle's C port instead.

    python tools/ref_python_timing.py      (needs /root/reference)
"""
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests" / "golden"))
from make_golden import load_reference  # noqa: E402  (installs the gymnasium / pygame stubs)


def episodes(n):
    SimpleGrid, *_ = load_reference()
    env = SimpleGrid(np.array([5, 5], np.int32), np.array([0, 0], np.int32), size=10)
    acts = np.random.default_rng(0).integers(1, 5, size=120)
    t0 = time.perf_counter()
    for _ in range(n):
        env.reset()
        for a in acts:
            env.step(int(a))
    return n * 120 / (time.perf_counter() - t0)


def pipeline(n):
    _, Est, BensLog, Norm, Welford = load_reference()
    rng = np.random.default_rng(0)
    est, w, nz, bl = Est(), Welford(), Norm(), BensLog()
    cells = rng.integers(0, 32, size=(n, 2))
    counts = rng.poisson(200.0, n)
    visits = {}
    t0 = time.perf_counter()
    for (x, y), c in zip(cells, counts):
        k = (int(x), int(y))
        est.update(key=k, value=int(c))
        m = est.get_estimate(k)
        w.update(m)
        z = w.standardize(m)
        nz.normalize(current_value=z, max=w.get_max(), min=w.get_min())
        visits[k] = visits.get(k, 0) + 1
        bl.normalize_incremental_logscale(visits[k], max=360, step_size=2)
    return n / (time.perf_counter() - t0)


if __name__ == "__main__":
    cores = os.cpu_count()
    one = episodes(500)
    with mp.Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(episodes, [300] * cores)
        allc = cores * 300 * 120 / (time.perf_counter() - t0)
    pipe = pipeline(20000)
    print(f"container CPU, {cores} cores (NOT the GPU box)")
    print(f"reference SimpleGrid reset+120 steps: {one:,.0f} env-steps/s on 1 core; {allc:,.0f} env-steps/s on {cores} processes")
    print(f"reference math_utils pipeline (estimator -> Welford -> Normalizer, BensLog): {pipe:,.0f} readings/s on 1 core")
    print(f"=> composed reference agent-step ~ {1 / (1 / one + 1 / pipe):,.0f} agent-steps/s/core (env step + observation statistics)")
