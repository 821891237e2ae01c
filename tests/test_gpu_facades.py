"""The reference-facing facades on the GPU: the reference's own tests (tests/test_envs.py,
test_standardization.py, test_normalization.py, test_estimator.py), transcribed against
rad_team_b200.envs / rad_team_b200.math_utils, plus the fixtures generated from the unmodified reference."""
import warnings

import numpy as np
import pytest

from .conftest import fh

pytestmark = pytest.mark.gpu


def arrays_match(a, b) -> bool:
    return bool(np.all(np.asarray(a) == np.asarray(b)))


# ---- tests/test_envs.py ----------------------------------------------------------------------
class Test_SimpleGrid:
    @pytest.fixture(autouse=True)
    def setup(self):
        from rad_team_b200.envs.simple_gridworld import SimpleGrid

        self.start = np.array([5, 5], dtype=np.int32)
        self.terminal = np.array([0, 0], dtype=np.int32)
        self.env = SimpleGrid(self.start, self.terminal, render_mode="rgb_array")
        self.original_distance = np.linalg.norm(self.start - self.terminal, ord=1)
        self.env.reset()
        yield
        self.env.close()

    def test_reset(self):
        obs, info = self.env.reset()
        assert arrays_match(obs, self.start)
        assert set(info.keys()) == {"agent", "target", "distance"}
        assert arrays_match(info["agent"], self.start) and arrays_match(info["target"], self.terminal)
        assert info["distance"] == 10
        assert self.env.agent_previous_dist == self.original_distance
        _ = self.env.step(1)
        obs, info = self.env.reset()
        assert arrays_match(obs, self.start) and info["distance"] == 10

    def test_step(self):
        from rad_team_b200.envs.simple_gridworld import Action, ActionDirectionMap, SimpleGrid

        _ = self.env.reset()
        with pytest.raises(ValueError, match="0 is not a valid Action"):
            _ = self.env.step(0)
        for act in Action:
            position_check = self.start + ActionDirectionMap[act]
            distance_check = np.linalg.norm(position_check - self.terminal, ord=1)
            obs, rew, done, trunc, info = self.env.step(act.value)
            assert not done and not trunc
            assert arrays_match(obs, position_check)
            assert rew == (-1 if distance_check > self.original_distance else 0)
            assert type(rew) is int and type(done) is bool
            assert set(info.keys()) == {"agent", "target", "distance"}
            assert arrays_match(info["agent"], position_check) and arrays_match(info["target"], self.terminal)
            assert info["distance"] == distance_check
            assert arrays_match(self.env.agent, position_check)
            assert self.env.agent_previous_dist == distance_check
            _ = self.env.reset()
        temp_terminal = self.start + 1
        terminal_env = SimpleGrid(self.start, temp_terminal)
        _ = terminal_env.reset()
        position_check = self.start + ActionDirectionMap[1]
        obs, rew, done, trunc, info = terminal_env.step(1)
        assert done and not trunc and arrays_match(obs, position_check) and rew == 1
        assert info["distance"] == np.linalg.norm(position_check - temp_terminal, ord=1)

    def test_helpers_and_render(self):
        assert self.env._get_distance() == self.original_distance
        o = self.env._get_oracle_obs()
        assert set(o.keys()) == {"agent", "target"} and arrays_match(o["agent"], self.start)
        assert arrays_match(self.env._get_obs(), self.start)
        assert self.env._get_info()["distance"] == self.original_distance
        frame = self.env.render()
        assert frame.shape == (512, 512, 3) and frame.dtype == np.uint8


def test_simplegrid_golden_episodes(golden_grid):
    from rad_team_b200.envs.simple_gridworld import SimpleGrid

    for ep in golden_grid["episodes"][::2]:
        env = SimpleGrid(np.array(ep["start"], np.int32), np.array(ep["terminal"], np.int32), size=ep["size"])
        obs, _ = env.reset()
        assert obs.tolist() == ep["start"]
        for a, x, y, rew, done, trunc, dist, prev in ep["rows"]:
            obs, r, d, tr, info = env.step(a)
            assert (int(obs[0]), int(obs[1]), r, d, tr) == (x, y, rew, done, trunc)
            assert float(info["distance"]) == dist and float(env.agent_previous_dist) == prev
        env.close()


# ---- tests/test_standardization.py -----------------------------------------------------------
class Test_Standardization:
    @pytest.fixture(autouse=True)
    def setup(self):
        from rad_team_b200.math_utils.standardize import WelfordsOnlineStandardization

        self.stats = WelfordsOnlineStandardization()

    def test_Update(self):
        with pytest.raises(AssertionError):
            self.stats.update(reading=-1.0)
        self.stats.update(reading=1000.0)
        assert self.stats.mean == 1000.0 and self.stats.count == 1 and self.stats._max == 0 and self.stats._min == 0
        self.stats.update(reading=2000.0)
        assert self.stats.count == 2 and self.stats.mean == 1500.0
        assert self.stats.square_dist_mean == 500000 and self.stats.sample_variance == 500000
        assert self.stats.std == pytest.approx(707.10678)
        assert self.stats._max == pytest.approx(0.70710678) and self.stats._min == 0
        self.stats.update(reading=100.0)
        assert self.stats.count == 3
        assert self.stats.mean == pytest.approx(1033.33333)
        assert self.stats.square_dist_mean == pytest.approx(1806666.66666)
        assert self.stats.sample_variance == pytest.approx(903333.33333)
        assert self.stats.std == pytest.approx(950.43849)
        assert self.stats._max == pytest.approx(0.70710678)
        assert self.stats._min == pytest.approx(-0.9820028733646521)

    def test_Standardize(self):
        with pytest.raises(AssertionError):
            self.stats.standardize(reading=-1.0)
        self.stats.update(reading=1000.0)
        assert self.stats.standardize(1) == -999 and self.stats.standardize(1000) == 0
        assert self.stats.standardize(10000) == 9000
        self.stats.update(reading=2000.0)
        assert self.stats.standardize(1) == pytest.approx(-2.11990612999)
        assert self.stats.standardize(1000) == pytest.approx(-0.7071067811865475)
        assert self.stats.standardize(10000) == pytest.approx(12.020815280171307)
        assert self.stats._max == pytest.approx(0.70710678) and self.stats._min == 0

    def test_GetMaxMin_and_Reset(self):
        self.stats.update(reading=1000.0)
        assert self.stats.get_max() == 0 and self.stats.get_min() == 0
        self.stats.update(reading=2000.0)
        self.stats.update(reading=100.0)
        assert self.stats.get_max() == pytest.approx(0.70710678)
        assert self.stats.get_min() == pytest.approx(-0.9820028733646521)
        self.stats.reset()
        assert (self.stats.mean, self.stats.square_dist_mean, self.stats.sample_variance, self.stats.std,
                self.stats.count, self.stats._max, self.stats._min) == (0.0, 0.0, 0.0, 1.0, 0, 0.0, 0.0)


def test_welford_streams_bit_exact_vs_reference(golden_math):
    """All golden streams at once through the batched kernel: fp64 state is bit-identical."""
    import torch
    from rad_team_b200.math_utils.standardize import BatchedWelford

    streams = golden_math["welford"]
    n = len(streams)
    w = BatchedWelford(n)
    for i in range(len(streams[0]["rows"])):
        rows = [s["rows"][i] for s in streams]
        x = torch.tensor([fh(r[0]) for r in rows], dtype=torch.float64)
        err = w.update(x)
        assert int(err.abs().sum().item()) == 0
        st, cnt = w.state()
        st, cnt = st.cpu().numpy(), cnt.cpu().numpy()
        z, _ = w.standardize(x)
        z17, _ = w.standardize(torch.full((n,), 17.0, dtype=torch.float64))
        for j, r in enumerate(rows):
            want = [fh(v) for v in r[2:8]]
            assert cnt[j] == r[1] and st[j].tolist() == want, (streams[j]["kind"], i)
            assert z[j].item() == fh(r[8]) and z17[j].item() == fh(r[9])


# ---- tests/test_normalization.py ---------------------------------------------------------------
class Test_Normalizers:
    def test_Normalize(self):
        from rad_team_b200.math_utils.normalize import Normalizer

        normalizer = Normalizer()
        for kw in (dict(current_value=1.0, max=0.0), dict(current_value=1.0, max=-1.0), dict(current_value=100, max=1.0),
                   dict(current_value=10, max=100, min=11)):
            with pytest.raises(AssertionError):
                normalizer.normalize(**kw)
        assert normalizer.normalize(current_value=-50.0, max=0.0) == 0
        assert normalizer.normalize(current_value=50, max=100) == 0.5
        assert normalizer.normalize(current_value=-501.0, max=100) == 0.0
        assert normalizer.normalize(current_value=0, max=100) == 0.0
        assert normalizer.normalize(current_value=50, max=100, min=10) == pytest.approx(0.4444444444444444)
        assert normalizer.normalize(current_value=50, max=100, min=-10) == pytest.approx(0.5454545454545454)
        assert normalizer.normalize(current_value=0, max=100, min=-10) == 0.0
        assert normalizer.normalize(current_value=50, max=100, min=0) == pytest.approx(0.5)

    def test_LogNormalize(self):
        from rad_team_b200.math_utils.normalize import BensLogNormalizer

        normalizer = BensLogNormalizer()
        for kw in (dict(current_value=-1.0, max=10), dict(current_value=1.0, max=-10), dict(current_value=1.0, max=0),
                   dict(current_value=1.0, max=10, step_size=0), dict(current_value=1.0, max=10, step_size=-1)):
            with pytest.raises(AssertionError):
                normalizer.normalize_incremental_logscale(**kw)
        f = normalizer.normalize_incremental_logscale
        assert f(current_value=4.0, max=10, step_size=2) == pytest.approx(0.598104004)
        assert f(current_value=4.0, max=10, step_size=2) == pytest.approx(f(current_value=4.0, max=10))
        assert f(current_value=18.0, max=10, step_size=2) == 1
        assert f(current_value=1.0, max=10, step_size=2) == pytest.approx(0.366725791)
        with pytest.raises(AssertionError):
            f(current_value=30.0, max=10, step_size=2)
        with pytest.warns(UserWarning):
            f(current_value=10.0, max=100, step_size=2)
        with pytest.warns(UserWarning):
            f(current_value=10.0, max=10, step_size=10)


def test_normalizer_tables_vs_reference(golden_math):
    import torch
    from rad_team_b200 import _lib
    from rad_team_b200.math_utils.normalize import lognormalize_batch, normalize_batch

    dev = torch.device("cuda")
    for with_min in (False, True):
        rows = [r for r in golden_math["normalizer"] if (r[2] is not None) == with_min]
        cur = torch.tensor([fh(r[0]) for r in rows], dtype=torch.float64, device=dev)
        mx = torch.tensor([fh(r[1]) for r in rows], dtype=torch.float64, device=dev)
        mn = torch.tensor([fh(r[2]) for r in rows], dtype=torch.float64, device=dev) if with_min else None
        out, err = normalize_batch(cur, mx, mn)
        out, err = out.cpu().tolist(), err.cpu().tolist()
        for r, o, e in zip(rows, out, err):
            if r[3] == "assert":
                assert e == _lib.RT_ERR_ASSERT, r
            else:
                assert e == 0 and o == fh(r[3]), r  # bit-exact
    for tab in golden_math["benslog"]:
        vs = torch.tensor([float(r[0]) for r in tab["rows"]], dtype=torch.float64, device=dev)
        out, err = lognormalize_batch(vs, tab["max"], tab["step"])
        out, err = out.cpu().tolist(), err.cpu().tolist()
        for r, o, e in zip(tab["rows"], out, err):
            if r[1] == "assert":
                assert e == _lib.RT_ERR_ASSERT, (tab["max"], r)
            else:  # device log is the deterministic <=1 ulp log, libm on the reference side
                assert e == 0 and o == pytest.approx(fh(r[1]), rel=1e-13, abs=1e-15), (tab["max"], r)


def test_env_visit_lut_is_the_reference_table(golden_math):
    from rad_team_b200.envs import BatchedRadSearchEnv

    tab = next(t for t in golden_math["benslog"] if t["max"] == 360 and t["step"] == 2)
    env = BatchedRadSearchEnv(n_envs=2, n_agents=3, grid=32, max_steps=120)
    lut = env.buffer("VISIT_LUT").cpu().numpy()
    assert lut[0] == 0
    for v, res in tab["rows"]:
        if 1 <= v <= 360:
            assert lut[v] == np.float32(fh(res))  # host libm, same bits as CPython's math.log


# ---- tests/test_estimator.py -------------------------------------------------------------------
class Test_IntensityResamplingEstimator:
    @pytest.fixture(autouse=True)
    def setup(self):
        from rad_team_b200.math_utils.estimator import IntensityResamplingEstimator

        self.estimator = IntensityResamplingEstimator(n_keys=64, capacity=256)

    def test_Update_and_GetBuffer(self):
        with pytest.raises(ValueError):
            self.estimator.get_buffer(key=(1, 2))
        self.estimator.update(key=(1, 2), value=1000)
        assert (1, 2) in self.estimator.readings.keys()
        assert [1000] in self.estimator.readings.values()
        self.estimator.update(key=(1, 2), value=2000)
        assert self.estimator.get_buffer(key=(1, 2)) == [1000, 2000]
        self.estimator.update(key=(3, 3), value=350)
        assert self.estimator.get_buffer(key=(1, 2)) == [1000, 2000]
        assert self.estimator.get_buffer(key=(3, 3)) == [350]

    def test_GetEstimate(self):
        with pytest.raises(ValueError):
            self.estimator.get_estimate(key=(1, 2))
        self.estimator.update(key=(1, 2), value=1000)
        self.estimator.update(key=(1, 2), value=2000)
        assert self.estimator.get_estimate(key=(1, 2)) == 1500
        self.estimator.update(key=(1, 2), value=500)
        assert self.estimator.get_estimate(key=(1, 2)) == 1000

    def test_GetMinMax(self):
        e = self.estimator
        assert e.get_max() == 0.0 and e.get_min() == 0.0
        e.update(key=(1, 2), value=1000)
        assert (e.get_max(), e.get_min()) == (1000, 1000)
        e.update(key=(1, 2), value=2000)
        assert (e.get_max(), e.get_min()) == (1500, 1000)
        e.update(key=(1, 2), value=300)
        e.update(key=(1, 2), value=300)
        assert (e.get_max(), e.get_min()) == (1500, 650)
        e.update(key=(3, 3), value=50)
        assert (e.get_max(), e.get_min()) == (1500, 50)
        e.update(key=(4, 4), value=3000)
        assert (e.get_max(), e.get_min()) == (3000, 50)

    def test_CheckKey_and_reset(self):
        assert self.estimator.check_key((1, 1)) is False
        self.estimator.update(key=(4, 4), value=3000)
        assert self.estimator.check_key((1, 1)) is False and self.estimator.check_key((4, 4)) is True
        self.estimator.reset()
        assert self.estimator.readings == {} and self.estimator._max == 0.0 and self.estimator._min == 0.0


def test_estimator_runs_vs_reference(golden_math):
    from rad_team_b200.math_utils.estimator import IntensityResamplingEstimator

    for run in golden_math["estimator"]:
        e = IntensityResamplingEstimator(n_keys=16, capacity=512)
        for key, v, est, emx, emn, blen in run["rows"][:150]:
            e.update(tuple(key), fh(v))
            assert (e.get_estimate(tuple(key)), e.get_max(), e.get_min()) == (fh(est), fh(emx), fh(emn))
        k0 = tuple(run["rows"][0][0])
        want = [fh(r[1]) for r in run["rows"][:150] if tuple(r[0]) == k0]
        assert [float(x) for x in e.get_buffer(k0)] == want


# ---- Gymnasium-style vector adapter -----------------------------------------------
def test_vector_env_adapter_autoreset_and_shapes():
    from oracle import oracle as O
    from rad_team_b200.envs import RadSearchVectorEnv

    kw = dict(poly_min=0, poly_max=2, rect_min=1, rect_max=3, seed=41)
    venv = RadSearchVectorEnv(num_envs=64, n_agents=2, grid=8, max_steps=12, **kw)
    ok = dict(O.DEFAULTS); ok.update(n_envs=64, n_agents=2, grid=8, max_steps=12, **kw)
    orc = O.OracleEnv(**ok)
    obs, info = venv.reset()
    xy_o, m_o = orc.reset()
    assert obs.shape == (64, 2, 2) and obs.dtype == np.int32 and np.array_equal(obs, xy_o)
    assert np.array_equal(info["maps"].cpu().numpy(), m_o)
    rng = np.random.default_rng(41)
    saw_trunc = saw_term = False
    tick = np.zeros(64, np.int32)
    for t in range(40):
        act = rng.integers(1, 5, size=(64, 2))
        obs, rew, term, trunc, info = venv.step(act)
        xy_o, m_o, rew_o, done_o = orc.step(act.astype(np.int32))
        tick += 1
        want_term = done_o.astype(bool).any(1)
        want_trunc = (tick >= 12) & ~want_term
        assert np.array_equal(term, want_term) and np.array_equal(trunc, want_trunc)
        assert rew.shape == (64,) and np.array_equal(rew, rew_o.sum(1)) and np.array_equal(info["agent_reward"], rew_o)
        fin = want_term | want_trunc
        if fin.any():
            assert np.array_equal(info["final_observation"], xy_o) and np.array_equal(info["_final_observation"], fin)
            xy_o, m_o = orc.reset(mask=fin.astype(np.uint8))
            tick[fin] = 0
            saw_trunc |= bool(want_trunc.any()); saw_term |= bool(want_term.any())
        assert np.array_equal(obs, xy_o)
        assert np.array_equal(info["maps"].cpu().numpy(), m_o)
    assert saw_trunc and saw_term
    with pytest.raises(ValueError, match="is not a valid Action"):
        venv.step(np.zeros((64, 2), np.int64))
    venv.close()


@pytest.mark.gpu
def test_batched_env_renders_an_rgb_frame_from_the_device_maps():
    import torch
    from rad_team_b200.envs import BatchedRadSearchEnv

    env = BatchedRadSearchEnv(n_envs=4, n_agents=2, grid=16, max_steps=30, poly_min=1, poly_max=3, seed=5)
    env.reset()
    for t in range(6):
        env.step(torch.full((4, 2), 1 + t % 4, dtype=torch.int32, device=env.device))
    frame = env.render(env_index=2, window_size=512)
    assert frame.shape == (512, 512, 3) and frame.dtype == np.uint8
    scale = 512 // 16
    xy = env.obs_xy.cpu().numpy()[2]
    src = env.buffer("SRC_XY").cpu().numpy()[2]
    for (x, y) in xy:  # agents are blue; row 0 of the frame is the top of the world
        px = frame[(15 - y) * scale + scale // 2, x * scale + scale // 2]
        assert tuple(px) == (0, 0, 255), (x, y, px)
    if not any((src == a).all() for a in xy):
        assert tuple(frame[(15 - src[1]) * scale + scale // 2, src[0] * scale + scale // 2]) == (255, 0, 0)
    assert (frame[0, :, :] == 0).all() and (frame[:, 0, :] == 0).all()  # grid lines


@pytest.mark.gpu
def test_reference_dataclass_constructor_fields_are_honoured_like_the_reference_does():
    from rad_team_b200.math_utils.estimator import IntensityResamplingEstimator
    from rad_team_b200.math_utils.standardize import WelfordsOnlineStandardization

    w = WelfordsOnlineStandardization(mean=5.0, std=3.0, count=7)   # the reference's __post_init__ resets them
    assert (w.mean, w.std, w.count) == (0.0, 1.0, 0)
    e = IntensityResamplingEstimator(readings={(1, 2): [1000, 2000, 500], (0, 0): [7]}, _max=9.0, _min=-1.0)
    assert e.get_buffer((1, 2)) == [1000, 2000, 500] and e.get_estimate((1, 2)) == 1000
    assert e.get_estimate((0, 0)) == 7 and (e.get_max(), e.get_min()) == (9.0, -1.0)
    e.update((0, 0), 9)
    assert e.get_estimate((0, 0)) == 8 and e.get_min() == -1.0 and e.get_max() == 9.0  # 8 < _max, 8 > _min: unchanged
    small = IntensityResamplingEstimator(capacity=3)
    for v in (1, 2, 3):
        small.update((0, 0), v)
    with pytest.raises(MemoryError, match="capacity=3"):
        small.update((0, 0), 4)


@pytest.mark.gpu
def test_simplegrid_episode_cap_is_reported_as_a_named_limit():
    from rad_team_b200.envs.simple_gridworld import SimpleGrid

    env = SimpleGrid(np.array([5, 5], np.int32), np.array([0, 0], np.int32), size=10, max_steps=3)
    env.reset()
    for _ in range(3):
        env.step(1)
    with pytest.raises(MemoryError, match="max_steps=3"):
        env.step(1)
    env.reset()          # a reset opens a new log; the sticky flag of the old episode stays readable but the env works
