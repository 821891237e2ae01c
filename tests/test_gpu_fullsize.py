"""BASELINE.json's full sizes, through size-independent properties (the oracle cannot step 65,536 envs
in seconds, so it checks a random SAMPLE of envs — envs are independent and their random streams are
functions of the global env id, so a 64-env oracle with env_id_base = k reproduces envs [k, k+64) of the
big batch exactly)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

C3 = dict(n_envs=65536, n_agents=3, grid=32, max_steps=120, poly_min=1, poly_max=5, seed=21)


def test_c3_full_size_properties_and_sampled_oracle_parity():
    import torch
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    env = BatchedRadSearchEnv(**C3)
    E, A, G, T = 65536, 3, 32, 40
    gen = torch.Generator(device=env.device); gen.manual_seed(5)
    acts = torch.randint(1, 5, (T, E, A), dtype=torch.int32, device=env.device, generator=gen)
    bases = [0, 12345, 40000, 65536 - 64]
    orcs = []
    for b in bases:
        kw = dict(O.DEFAULTS); kw.update(C3); kw.update(n_envs=64, env_id_base=b)
        orcs.append(O.OracleEnv(**kw))
    env.reset()
    for o in orcs:
        o.reset()
    ret = torch.zeros((E, A), dtype=torch.int64, device=env.device)
    for t in range(T):
        xy, rew, done, _, _ = env.step(acts[t])
        ret += rew
        if t % 13 == 0 or t == T - 1:
            m = env.obs_maps
            # every map value is in [0, 1]; location channel marks exactly the agents' cells
            assert float(m.min()) >= 0.0 and float(m.max()) <= 1.0
            loc = m[:, 0].reshape(E, -1)
            cells = (xy[..., 1] * G + xy[..., 0]).long()
            assert torch.all(loc.gather(1, cells) == 1.0)
            n_distinct = torch.tensor([1], device=env.device)  # at most A ones per env
            assert int(loc.sum(1).max()) <= A and int(loc.sum(1).min()) >= 1
            # visits: every env has counted A readings per step; visit channel is non-zero exactly there
            visit = env.buffer("VISIT").to(torch.int32)
            assert torch.all(visit.sum(1) == (t + 1) * A)
            assert torch.equal(m[:, 2].reshape(E, -1) > 0, visit > 0)
            # record list length == number of visited cells
            assert torch.equal(env.buffer("VREC")[:, 0, 0], (visit > 0).sum(1).to(torch.int32))
        for b, o in zip(bases, orcs):
            a_np = acts[t, b:b + 64].cpu().numpy()
            xy_o, m_o, rew_o, done_o = o.step(a_np)
            assert np.array_equal(xy[b:b + 64].cpu().numpy(), xy_o)
            assert np.array_equal(rew[b:b + 64].cpu().numpy(), rew_o) and np.array_equal(done[b:b + 64].cpu().numpy(), done_o)
            assert np.array_equal(env.buffer("LAST_COUNT")[b:b + 64].cpu().numpy(), o.export("LAST_COUNT"))
            assert np.array_equal(env.buffer("LAST_LOS")[b:b + 64].cpu().numpy(), o.export("LAST_LOS"))
            assert np.array_equal(env.obs_maps[b:b + 64].cpu().numpy(), m_o)
    assert env.errors() == 0
    # episode statistics kernel == plain reductions
    st = env.episode_stats().cpu().numpy()
    assert st[0] == float(ret.sum()) and st[1] == E * A * T
    assert st[2] == float(env.buffer("EP_DONE").sum())
    # determinism: a second handle with the same seed reproduces the tensor bit for bit
    twin = BatchedRadSearchEnv(**C3)
    twin.reset()
    for t in range(T):
        twin.step(acts[t])
    assert torch.equal(twin.obs_maps, env.obs_maps) and torch.equal(twin.buffer("WELFORD"), env.buffer("WELFORD"))


def test_c2_size_single_agent_whole_episode_vs_oracle():
    """BASELINE configs[1] exactly: 4,096 envs x 1 agent, no obstructions, a whole 120-step episode."""
    import torch
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    kw = dict(O.DEFAULTS); kw.update(n_envs=4096, n_agents=1, grid=32, max_steps=120, seed=22)
    g, o = BatchedRadSearchEnv(**kw), O.OracleEnv(**kw)
    g.reset(); o.reset()
    rng = np.random.default_rng(22)
    for t in range(120):
        a = rng.integers(1, 5, size=(4096, 1)).astype(np.int32)
        xy, rew, done, _, _ = g.step(torch.as_tensor(a, device=g.device))
        xy_o, m_o, rew_o, done_o = o.step(a, maps=(t % 17 == 0 or t == 119))
        assert np.array_equal(xy.cpu().numpy(), xy_o) and np.array_equal(rew.cpu().numpy(), rew_o)
        assert np.array_equal(g.buffer("LAST_COUNT").cpu().numpy(), o.export("LAST_COUNT"))
        if m_o is not None:
            assert np.array_equal(g.obs_maps.cpu().numpy(), m_o)
    for k in ("VISIT", "INTENSITY", "WELFORD", "EST_MAXMIN"):
        assert np.array_equal(g.buffer(k).cpu().numpy(), o.export(k)), k


@pytest.mark.parametrize("kw", [
    dict(n_envs=3, n_agents=2, grid=2, max_steps=5),                       # smallest grid
    dict(n_envs=2, n_agents=8, grid=64, max_steps=8191, poly_min=5, poly_max=5, rect_min=1, rect_max=9),  # T*A = 65,528: near the u16 limit; dense fallback rasteriser? no: 4096 cells
    dict(n_envs=2, n_agents=2, grid=248, max_steps=30),                    # 61,504 cells: tile too big for smem -> dense vector path
    dict(n_envs=1, n_agents=1, grid=255, max_steps=30),                    # largest grid, 65,025 cells -> generic path
    dict(n_envs=5, n_agents=3, grid=13, max_steps=25, poly_min=1, poly_max=3, rect_min=1, rect_max=3),   # 169 cells: generic rasteriser
])
def test_extreme_configurations_vs_oracle(kw):
    import torch
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    full = dict(O.DEFAULTS); full.update(kw); full["seed"] = 23
    g, o = BatchedRadSearchEnv(**full), O.OracleEnv(**full)
    xy, _ = g.reset(); xy_o, m_o = o.reset()
    assert np.array_equal(xy.cpu().numpy(), xy_o) and np.array_equal(g.obs_maps.cpu().numpy(), m_o)
    rng = np.random.default_rng(23)
    steps = min(full["max_steps"], 60)
    for t in range(steps):
        a = rng.integers(1, 5, size=(full["n_envs"], full["n_agents"])).astype(np.int32)
        xy, rew, done, _, _ = g.step(torch.as_tensor(a, device=g.device))
        xy_o, m_o, rew_o, done_o = o.step(a)
        assert np.array_equal(xy.cpu().numpy(), xy_o) and np.array_equal(rew.cpu().numpy(), rew_o)
        assert np.array_equal(g.obs_maps.cpu().numpy(), m_o), t
    assert np.array_equal(g.buffer("VISIT_LUT").cpu().numpy(), o.export("VISIT_LUT"))
    assert g.errors() == o.errors() == 0


def test_c4_shard_of_rank_7_of_8_sampled_oracle_parity():
    """BASELINE configs[3] as rank 7 of 8 sees it: 131,072 envs whose GLOBAL ids start at 917,504.  Sampled blocks of
    the shard are reproduced by 64-env oracles placed at the same global ids (RNG streams are functions of the global
    env id, so this is also what a one-GPU run of all 1,048,576 envs computes for them)."""
    import torch
    from oracle import oracle as O
    from rad_team_b200.distributed import shard_range
    from rad_team_b200.envs import BatchedRadSearchEnv

    base, count = shard_range(1_048_576, 7, 8)
    assert (base, count) == (917_504, 131_072)
    cfg = dict(n_envs=count, n_agents=3, grid=32, max_steps=120, poly_min=1, poly_max=5, seed=0, env_id_base=base)
    env = BatchedRadSearchEnv(**cfg)
    T = 24
    gen = torch.Generator(device=env.device); gen.manual_seed(9)
    acts = torch.randint(1, 5, (T, count, 3), dtype=torch.int32, device=env.device, generator=gen)
    offs = [0, 70_001, count - 64]
    orcs = []
    for off in offs:
        kw = dict(O.DEFAULTS); kw.update(cfg); kw.update(n_envs=64, env_id_base=base + off)
        orcs.append(O.OracleEnv(**kw))
    xy, _ = env.reset()
    for off, o in zip(offs, orcs):
        xy_o, m_o = o.reset()
        assert np.array_equal(xy[off:off + 64].cpu().numpy(), xy_o)
        assert np.array_equal(env.obs_maps[off:off + 64].cpu().numpy(), m_o)
    for t in range(T):
        xy, rew, done, _, _ = env.step(acts[t])
        for off, o in zip(offs, orcs):
            xy_o, m_o, rew_o, done_o = o.step(acts[t, off:off + 64].cpu().numpy())
            assert np.array_equal(xy[off:off + 64].cpu().numpy(), xy_o), (t, off)
            assert np.array_equal(rew[off:off + 64].cpu().numpy(), rew_o)
            assert np.array_equal(env.buffer("LAST_COUNT")[off:off + 64].cpu().numpy(), o.export("LAST_COUNT"))
            assert np.array_equal(env.buffer("LAST_LOS")[off:off + 64].cpu().numpy(), o.export("LAST_LOS"))
            assert np.array_equal(env.obs_maps[off:off + 64].cpu().numpy(), m_o), (t, off)
    visit = env.buffer("VISIT").to(torch.int32)
    assert torch.all(visit.sum(1) == T * 3) and env.errors() == 0
