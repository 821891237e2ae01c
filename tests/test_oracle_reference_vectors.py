"""The oracle against the known-answer vectors of the reference's OWN tests:
tests/test_standardization.py:12-102, tests/test_normalization.py:7-76, tests/test_estimator.py:53-107,
tests/test_envs.py:68-162.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O


def test_welford_update_vectors():  # tests/test_standardization.py:12-44
    w = O.Welford()
    assert w.update(-1.0) == 1  # AssertionError in the reference
    assert w.s.count == 0
    w.update(1000.0)
    assert (w.s.mean, w.s.count, w.s.zmax, w.s.zmin) == (1000.0, 1, 0.0, 0.0)
    w.update(2000.0)
    assert w.s.count == 2 and w.s.mean == 1500.0 and w.s.m2 == 500000 and w.s.var == 500000
    assert w.s.std == pytest.approx(707.10678)
    assert w.s.zmax == pytest.approx(0.70710678) and w.s.zmin == 0
    w.update(100.0)
    assert w.s.count == 3
    assert w.s.mean == pytest.approx(1033.33333)
    assert w.s.m2 == pytest.approx(1806666.66666)
    assert w.s.var == pytest.approx(903333.33333)
    assert w.s.std == pytest.approx(950.43849)
    assert w.s.zmax == pytest.approx(0.70710678)
    assert w.s.zmin == pytest.approx(-0.9820028733646521)


def test_welford_standardize_vectors():  # tests/test_standardization.py:46-67
    w = O.Welford()
    w.update(1000.0)
    assert w.standardize(1) == -999 and w.standardize(1000) == 0 and w.standardize(10000) == 9000
    w.update(2000.0)
    assert w.standardize(1) == pytest.approx(-2.11990612999)
    assert w.standardize(1000) == pytest.approx(-0.7071067811865475)
    assert w.standardize(10000) == pytest.approx(12.020815280171307)
    assert w.s.zmax == pytest.approx(0.70710678) and w.s.zmin == 0


def test_normalizer_vectors():  # tests/test_normalization.py:7-37
    assert O.normalize(1.0, 0.0)[0] == 1
    assert O.normalize(1.0, -1.0)[0] == 1
    assert O.normalize(100, 1.0)[0] == 1
    assert O.normalize(-50.0, 0.0) == (0, 0.0)
    assert O.normalize(10, 100, 11)[0] == 1
    assert O.normalize(50, 100) == (0, 0.5)
    assert O.normalize(-501.0, 100) == (0, 0.0)
    assert O.normalize(0, 100) == (0, 0.0)
    assert O.normalize(50, 100, 10)[1] == pytest.approx(0.4444444444444444)
    assert O.normalize(50, 100, -10)[1] == pytest.approx(0.5454545454545454)
    assert O.normalize(0, 100, -10) == (0, 0.0)
    assert O.normalize(50, 100, 0)[1] == pytest.approx(0.5)


def test_benslog_vectors():  # tests/test_normalization.py:40-76
    for args in ((-1.0, 10, 2), (1.0, -10, 2), (1.0, 0, 2), (1.0, 10, 0), (1.0, 10, -1)):
        assert O.lognormalize(*args)[0] == 1
    assert O.lognormalize(4.0, 10, 2)[1] == pytest.approx(0.598104004)
    assert O.lognormalize(18.0, 10, 2) == (0, 1.0)
    assert O.lognormalize(1.0, 10, 2)[1] == pytest.approx(0.366725791)
    assert O.lognormalize(30.0, 10, 2)[0] == 1


def test_estimator_vectors():  # tests/test_estimator.py:53-107
    assert O.median([1000, 2000]) == 1500
    assert O.median([1000, 2000, 500]) == 1000
    mx = mn = 0.0
    buf = []
    seq = [(1000, 1000, 1000), (2000, 1500, 1000), (300, None, None), (300, 1500, 650)]
    for v, emx, emn in seq:
        buf.append(v)
        mx, mn = O.est_maxmin(O.median(buf), mx, mn)
        if emx is not None:
            assert (mx, mn) == (emx, emn)
    mx, mn = O.est_maxmin(O.median([50]), mx, mn)
    assert (mx, mn) == (1500, 50)
    mx, mn = O.est_maxmin(O.median([3000]), mx, mn)
    assert (mx, mn) == (3000, 50)


def _grid(start, terminal, size=10):
    env = O.OracleEnv(n_envs=1, n_agents=1, grid=size, max_steps=500, fixed_scenario=1)
    env.set_scenario(src_xy=terminal, src_int=[1e6], bkg=[10.0], start_xy=start, n_poly=[0])
    return env


def test_env_reset_and_step_vectors():  # tests/test_envs.py:68-162
    env = _grid((5, 5), (0, 0))
    xy, _ = env.reset()
    assert xy.tolist() == [[[5, 5]]] and env.export("PREV_DIST")[0, 0] == 10
    dirs = {1: (0, 1), 2: (1, 0), 3: (-1, 0), 4: (0, -1)}  # simple_gridworld.py:26-31
    for act, (dx, dy) in dirs.items():
        xy, _, rew, done = env.step([[act]])
        want = (5 + dx, 5 + dy)
        assert tuple(xy[0, 0]) == want and not done[0, 0]
        d = want[0] + want[1]
        assert rew[0, 0] == (-1 if d > 10 else 0)
        assert env.export("PREV_DIST")[0, 0] == d
        xy, _ = env.reset()
        assert xy.tolist() == [[[5, 5]]]
    # invalid action: the reference raises ValueError and leaves the state alone (:104-105,138)
    xy, _, rew, done = env.step([[0]])
    assert env.errors() & 0x01 and tuple(xy[0, 0]) == (5, 5) and env.export("TICK")[0] == 0
    # terminal step (:140-162)
    env = _grid((5, 5), (6, 6))
    env.reset()
    xy, _, rew, done = env.step([[1]])
    assert tuple(xy[0, 0]) == (5, 6) and done[0, 0] == 1 and rew[0, 0] == 1
    # wall bump: clipped, distance unchanged => -1
    env = _grid((0, 0), (9, 9))
    env.reset()
    xy, _, rew, done = env.step([[3]])
    assert tuple(xy[0, 0]) == (0, 0) and rew[0, 0] == -1
