"""Property tests of the oracle's building blocks against independent implementations (statistics,
NumPy) — hypothesis generates the inputs.  CPU only."""
import math
import statistics

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import oracle as O

counts = st.lists(st.integers(min_value=0, max_value=2 ** 30), min_size=1, max_size=200)
floats = st.lists(st.floats(min_value=0.0, max_value=1e9, allow_nan=False, allow_infinity=False), min_size=2, max_size=300)


@settings(max_examples=200, deadline=None)
@given(counts)
def test_median_is_statistics_median(xs):
    assert O.median([float(x) for x in xs]) == statistics.median(xs)


@settings(max_examples=150, deadline=None)
@given(floats)
def test_welford_stream_matches_two_pass_moments(xs):
    w = O.Welford()
    for x in xs:
        assert w.update(x) == 0
    mean = math.fsum(xs) / len(xs)
    var = math.fsum((x - mean) ** 2 for x in xs) / (len(xs) - 1)
    scale = max(abs(mean), 1.0)
    assert abs(w.s.mean - mean) <= 1e-9 * scale
    assert abs(w.s.var - var) <= 1e-7 * max(var, scale * scale * 1e-6, 1e-12)
    assert w.s.std == max(math.sqrt(w.s.var), 1.0)          # standardize.py:74
    assert w.s.zmax >= 0.0 >= w.s.zmin                       # both start at 0 (standardize.py:43-44)
    z = w.standardize(xs[-1])
    assert w.s.zmin <= z <= w.s.zmax or len(xs) == 1
    rc, v = O.normalize(z, w.s.zmax, w.s.zmin)               # the pipeline's normalisation never asserts
    assert rc == 0 and 0.0 <= v <= 1.0


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(min_value=0.0, max_value=1e7, allow_nan=False, width=32), min_size=1, max_size=400),
       st.integers(min_value=1, max_value=6))
def test_chan_merge_of_shards_matches_one_pass(xs, parts):
    x = np.asarray(xs, np.float32)
    whole = O.moments_update(np.zeros(3), x)
    merged = np.zeros(3)
    for shard in np.array_split(x, parts):
        merged = O.moments_merge(merged, O.moments_update(np.zeros(3), shard))
    assert merged[0] == whole[0] == len(xs)
    assert abs(merged[1] - whole[1]) <= 1e-9 * max(abs(whole[1]), 1.0)
    assert abs(merged[2] - whole[2]) <= 1e-8 * max(whole[2], 1.0)
    assert abs(whole[1] - float(np.mean(x.astype(np.float64)))) <= 1e-9 * max(abs(whole[1]), 1.0)


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=0, max_value=718), st.sampled_from([(360, 2), (120, 2), (10, 2), (7, 3)]))
def test_benslog_is_monotone_and_bounded(v, cfg):
    mx, step = cfg
    rc, r = O.lognormalize(v, mx, step)
    if v <= mx * step - step:
        assert rc == 0 and 0.0 <= r <= 1.0
        if v > 0:
            assert r >= O.lognormalize(v - 1, mx, step)[1]
    else:
        assert rc == 1
