"""Seeded differential fuzzing: random configurations, random actions (with occasional invalid ones),
random masked resets, injected uniforms on some runs — CUDA vs oracle, every state array, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATE = ["AGENT_XY", "PREV_DIST", "START_XY", "SRC_XY", "SRC_INT", "BKG", "N_POLY", "EDGES", "TICK", "EPISODE",
         "VISIT", "INTENSITY", "WELFORD", "WELFORD_N", "EST_MAXMIN", "FLAGS", "LAST_COUNT", "LAST_LOS",
         "LAST_LAMBDA", "LAST_EST", "LAST_Z", "LAST_VALUE"]


def random_cfg(rng):
    grid = int(rng.choice([3, 5, 8, 11, 16, 24, 32, 40]))
    agents = int(rng.integers(1, 9))
    pmax = int(rng.integers(0, 6))
    pmin = int(rng.integers(0, pmax + 1))
    rmax = int(rng.integers(1, max(2, grid // 2)))
    lo = 10 ** rng.uniform(1, 6)
    return dict(n_envs=int(rng.integers(1, 300)), n_agents=agents, grid=grid, max_steps=int(rng.integers(3, 90)),
                poly_min=pmin, poly_max=pmax, rect_min=int(rng.integers(1, rmax + 1)), rect_max=rmax,
                seed=int(rng.integers(0, 2 ** 62)), env_id_base=int(rng.integers(0, 2 ** 40)),
                cell_pitch_cm=float(rng.uniform(10, 500)), min_dist_cm=float(rng.uniform(1, 200)),
                transmission=float(rng.choice([0.0, 0.1, 0.5, 1.0])), src_lo=float(lo), src_hi=float(lo * rng.uniform(1, 20)),
                bkg_lo=float(rng.uniform(0, 5)), bkg_hi=float(rng.uniform(5, 60)))


@pytest.mark.parametrize("case", range(16))
def test_random_configuration(case):
    import torch
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    rng = np.random.default_rng(1000 + case)
    kw = dict(O.DEFAULTS); kw.update(random_cfg(rng))
    E, A, T = kw["n_envs"], kw["n_agents"], kw["max_steps"]
    g, o = BatchedRadSearchEnv(**kw), O.OracleEnv(**kw)
    inject = case % 4 == 3

    def compare(where):
        for k in STATE:
            got, want = g.buffer(k).cpu().numpy(), o.export(k)
            assert np.array_equal(got, want), f"case {case} {k} {where} cfg={kw}"

    xy, _ = g.reset(); xy_o, m_o = o.reset()
    assert np.array_equal(xy.cpu().numpy(), xy_o) and np.array_equal(g.obs_maps.cpu().numpy(), m_o)
    for ep in range(2):
        for t in range(T + (2 if ep == 0 else 0)):       # episode 0 runs past max_steps: LOG_FULL on both sides
            act = rng.integers(1, 5, size=(E, A)).astype(np.int32)
            if rng.random() < 0.1:
                act[rng.integers(0, E), rng.integers(0, A)] = int(rng.choice([0, 5, -1, 99]))
            if inject:
                u = rng.random((E, A, 48))
                g.inject_uniforms(u); o.inject_uniforms(u)
            xy, rew, done, _, _ = g.step(torch.as_tensor(act, device=g.device))
            xy_o, m_o, rew_o, done_o = o.step(act)
            assert np.array_equal(xy.cpu().numpy(), xy_o) and np.array_equal(rew.cpu().numpy(), rew_o)
            assert np.array_equal(done.cpu().numpy(), done_o)
            assert np.array_equal(g.obs_maps.cpu().numpy(), m_o), f"case {case} maps step {t} cfg={kw}"
            if t % 7 == 0:
                compare(f"ep {ep} step {t}")
            if rng.random() < 0.08:
                mask = (rng.random(E) < 0.5).astype(np.uint8)
                xy, _ = g.reset(mask=torch.as_tensor(mask, device=g.device)); xy_o, m_o = o.reset(mask=mask)
                assert np.array_equal(xy.cpu().numpy(), xy_o) and np.array_equal(g.obs_maps.cpu().numpy(), m_o)
        compare(f"end of ep {ep}")
        assert g.errors() == o.errors()
        xy, _ = g.reset(); xy_o, m_o = o.reset()
        assert np.array_equal(g.obs_maps.cpu().numpy(), m_o)
