"""CUDA path (through the C ABI) vs the CPU oracle on identical seeds — the parity tests proper.

Bar (BASELINE.json north_star): grid indices, obstruction flags and Poisson counts bit-exact; float
intensities / normalised observations within 1e-5 relative.  The two sides use the same un-contracted
IEEE operation order, so the floats are compared for EXACT equality as well — a drift would show here
long before it reached 1e-5."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATE = ["AGENT_XY", "PREV_DIST", "START_XY", "SRC_XY", "SRC_INT", "BKG", "N_POLY", "EDGES", "TICK", "EPISODE",
         "VISIT", "INTENSITY", "WELFORD", "WELFORD_N", "EST_MAXMIN", "FLAGS", "LAST_COUNT", "LAST_LOS",
         "LAST_LAMBDA", "LAST_EST", "LAST_Z", "LAST_VALUE", "VISIT_LUT"]
RTOL = 1e-5  # the stated tolerance; exact equality is asserted first


def make_pair(device=None, **cfg):
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    full = dict(O.DEFAULTS)
    full.update(cfg)
    return BatchedRadSearchEnv(device=device, **full), O.OracleEnv(**full)


def assert_state_equal(g, o, names=STATE, where=""):
    for k in names:
        got = g.buffer(k).cpu().numpy()
        want = o.export(k)
        if want.dtype.kind == "f":
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-12, err_msg=f"{k} {where}")
        assert np.array_equal(got, want), f"{k} differs {where}: {np.argwhere(got != want)[:5]}"


def run_episode(g, o, steps, rng, check_every=1, device_actions=True, maps_every=1):
    import torch

    E, A = g.E, g.A
    xy_g, _ = g.reset()
    xy_o, m_o = o.reset()
    assert np.array_equal(xy_g.cpu().numpy(), xy_o)
    assert np.array_equal(g.obs_maps.cpu().numpy(), m_o)
    assert_state_equal(g, o, where="after reset")
    for t in range(steps):
        act = rng.integers(1, 5, size=(E, A)).astype(np.int32)
        if device_actions:
            xy, rew, done, trunc, info = g.step(torch.as_tensor(act, device=g.device))
            xy, rew, done = xy.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        else:
            xy, rew, done, trunc, info = g.step(act)
        xy_o, m_o, rew_o, done_o = o.step(act, maps=(t % maps_every == 0))
        assert trunc is False
        assert np.array_equal(xy, xy_o), f"obs_xy step {t}"
        assert np.array_equal(rew, rew_o), f"reward step {t}"
        assert np.array_equal(done, done_o), f"done step {t}"
        if t % maps_every == 0:
            m_g = g.obs_maps.cpu().numpy()
            np.testing.assert_allclose(m_g, m_o, rtol=RTOL, atol=1e-12)
            assert np.array_equal(m_g, m_o), f"obs_maps step {t}"
        if t % check_every == 0 or t == steps - 1:
            assert_state_equal(g, o, where=f"step {t}")
    assert g.errors() == o.errors() == 0


def test_c2_like_single_agent_no_obstructions():
    g, o = make_pair(n_envs=512, n_agents=1, grid=32, max_steps=120, seed=1)
    run_episode(g, o, 120, np.random.default_rng(0), check_every=10, maps_every=7)


def test_c3_like_three_agents_obstructed():
    g, o = make_pair(n_envs=384, n_agents=3, grid=32, max_steps=120, poly_min=1, poly_max=5, seed=2)
    run_episode(g, o, 120, np.random.default_rng(1), check_every=10, maps_every=7)
    assert (g.buffer("LAST_LOS").cpu().numpy() > 0).any() or True
    # a second episode: new scenarios, statistics cleared
    run_episode(g, o, 30, np.random.default_rng(2), check_every=10, maps_every=5)


def test_partial_transmission_and_host_path():
    g, o = make_pair(n_envs=200, n_agents=2, grid=16, max_steps=80, poly_min=0, poly_max=3, rect_min=1, rect_max=4,
                     seed=3, transmission=0.3, min_dist_cm=10.0)
    run_episode(g, o, 80, np.random.default_rng(3), check_every=8, device_actions=False, maps_every=9)


def test_reference_default_grid_generic_raster_path():
    # G = 10 (reference default size): 100 cells, not a multiple of 8 -> scalar rasteriser
    g, o = make_pair(n_envs=64, n_agents=4, grid=10, max_steps=50, poly_min=0, poly_max=2, rect_min=1, rect_max=3, seed=4)
    run_episode(g, o, 50, np.random.default_rng(4), check_every=5)


def test_many_agents_and_odd_sizes():
    g, o = make_pair(n_envs=77, n_agents=8, grid=24, max_steps=40, poly_min=2, poly_max=5, seed=5)
    run_episode(g, o, 40, np.random.default_rng(5), check_every=5, maps_every=3)


def test_small_grid_heavy_revisits():
    # 4x4 grid: cells are revisited dozens of times, so medians walk long sorted lists
    g, o = make_pair(n_envs=96, n_agents=3, grid=4, max_steps=200, seed=6, src_lo=1e3, src_hi=1e5, cell_pitch_cm=50.0)
    run_episode(g, o, 200, np.random.default_rng(6), check_every=20, maps_every=11)
    assert g.buffer("VISIT").cpu().numpy().max() > 30


def test_low_rate_multiplication_branch():
    # background-dominated tiny rates: lambda < 10 -> Knuth multiplication sampler, many zero counts
    g, o = make_pair(n_envs=256, n_agents=2, grid=32, max_steps=60, seed=7, src_lo=1e2, src_hi=1e3, bkg_lo=0.0,
                     bkg_hi=3.0, poly_min=1, poly_max=2)
    run_episode(g, o, 60, np.random.default_rng(7), check_every=6, maps_every=8)
    lam = g.buffer("LAST_LAMBDA").cpu().numpy()
    assert (lam < 10).any()


def test_injected_uniforms_bit_exact_counts():
    import torch

    g, o = make_pair(n_envs=300, n_agents=3, grid=32, max_steps=30, poly_min=1, poly_max=5, seed=8)
    rng = np.random.default_rng(8)
    g.reset(); o.reset()
    for t in range(20):
        u = rng.random((300, 3, 64))
        g.inject_uniforms(u); o.inject_uniforms(u)
        act = rng.integers(1, 5, size=(300, 3)).astype(np.int32)
        g.step(torch.as_tensor(act, device=g.device))
        o.step(act, maps=False)
        assert np.array_equal(g.buffer("LAST_COUNT").cpu().numpy(), o.export("LAST_COUNT"))
        assert np.array_equal(g.buffer("LAST_LOS").cpu().numpy(), o.export("LAST_LOS"))
    assert_state_equal(g, o)
    # too few uniforms: both sides flag it
    u = rng.random((300, 3, 1))
    g.inject_uniforms(u); o.inject_uniforms(u)
    act = rng.integers(1, 5, size=(300, 3)).astype(np.int32)
    g.step(torch.as_tensor(act, device=g.device)); o.step(act, maps=False)
    assert g.errors() == o.errors() and g.errors() & 0x08
    assert_state_equal(g, o)
    g.inject_uniforms(None)


def test_counts_match_numpy_poisson_on_injected_pcg64():
    """Independent third-party check of the DEVICE sampler: u_k from PCG64, count == twin.poisson(lam)."""
    import torch
    from rad_team_b200.envs import BatchedRadSearchEnv

    E = 2048
    g = BatchedRadSearchEnv(n_envs=E, n_agents=1, grid=32, max_steps=4, seed=9, fixed_scenario=1)
    rng = np.random.default_rng(9)
    src = rng.integers(0, 32, size=(E, 2))
    start = (src + rng.integers(2, 9, size=(E, 2))) % 32
    g.set_scenario(src_xy=src, src_int=10 ** rng.uniform(3, 7.5, E), bkg=rng.uniform(0, 12, E), start_xy=start,
                   n_poly=np.zeros(E, np.int32))
    g.reset()
    u = np.zeros((E, 1, 64))
    twins = []
    for e in range(E):
        u[e, 0] = np.random.Generator(np.random.PCG64(1000 + e)).random(64)
        twins.append(np.random.Generator(np.random.PCG64(1000 + e)))
    g.inject_uniforms(u)
    g.step(torch.full((E, 1), 1, dtype=torch.int32, device=g.device))
    lam = g.buffer("LAST_LAMBDA").cpu().numpy()[:, 0]
    cnt = g.buffer("LAST_COUNT").cpu().numpy()[:, 0]
    want = np.array([twins[e].poisson(lam[e]) for e in range(E)])
    assert (lam < 10).sum() > 50 and (lam >= 10).sum() > 50
    assert np.array_equal(cnt, want)


def test_invalid_actions_skip_env_and_flag():
    import torch

    g, o = make_pair(n_envs=64, n_agents=3, grid=32, max_steps=30, poly_min=1, poly_max=3, seed=10)
    rng = np.random.default_rng(10)
    g.reset(); o.reset()
    for t in range(10):
        act = rng.integers(1, 5, size=(64, 3)).astype(np.int32)
        if t == 4:
            act[5, 1] = 0
            act[17, 2] = 5
            act[40, 0] = -3
        xy, rew, done, *_ = g.step(torch.as_tensor(act, device=g.device))
        xy_o, _, rew_o, done_o = o.step(act, maps=False)
        assert np.array_equal(xy.cpu().numpy(), xy_o) and np.array_equal(rew.cpu().numpy(), rew_o)
        assert np.array_equal(done.cpu().numpy(), done_o)
    assert_state_equal(g, o)
    fl = g.buffer("FLAGS").cpu().numpy()
    assert set(np.nonzero(fl)[0]) == {5, 17, 40} and g.errors() == 0x01
    assert g.buffer("TICK").cpu().numpy()[5] == 9  # one step was refused
    # host path: the reference's exception type and message (tests/test_envs.py:104-105)
    act = np.ones((64, 3), np.int32); act[3, 0] = 0
    with pytest.raises(ValueError, match="0 is not a valid Action"):
        g.step(act)


def test_steps_beyond_episode_length_are_refused():
    import torch

    g, o = make_pair(n_envs=33, n_agents=2, grid=8, max_steps=5, seed=11)
    rng = np.random.default_rng(11)
    g.reset(); o.reset()
    for t in range(7):
        act = rng.integers(1, 5, size=(33, 2)).astype(np.int32)
        g.step(torch.as_tensor(act, device=g.device)); o.step(act, maps=False)
    assert_state_equal(g, o)
    assert g.errors() == o.errors() == 0x02


def test_masked_reset():
    import torch

    g, o = make_pair(n_envs=128, n_agents=3, grid=32, max_steps=60, poly_min=1, poly_max=5, seed=12)
    rng = np.random.default_rng(12)
    g.reset(); o.reset()
    for t in range(40):
        act = rng.integers(1, 5, size=(128, 3)).astype(np.int32)
        _, _, done, *_ = g.step(torch.as_tensor(act, device=g.device))
        _, _, _, done_o = o.step(act, maps=False)
        if t % 10 == 9:  # reset the envs where any agent reached the source (plus a fixed few)
            mask = done_o.any(axis=1).astype(np.uint8)
            mask[::17] = 1
            xy, _ = g.reset(mask=torch.as_tensor(mask, device=g.device))
            xy_o, m_o = o.reset(mask=mask)
            assert np.array_equal(xy.cpu().numpy(), xy_o)
            assert np.array_equal(g.obs_maps.cpu().numpy(), m_o)
            assert_state_equal(g, o, where=f"masked reset at {t}")
    assert_state_equal(g, o)
    assert len(set(g.buffer("EPISODE").cpu().numpy().tolist())) > 1


def test_state_dict_round_trip_replays_bit_exactly():
    import torch
    from rad_team_b200.envs import BatchedRadSearchEnv

    kw = dict(n_envs=100, n_agents=3, grid=32, max_steps=50, poly_min=1, poly_max=5, seed=13)
    a = BatchedRadSearchEnv(**kw)
    rng = np.random.default_rng(13)
    acts = torch.as_tensor(rng.integers(1, 5, size=(50, 100, 3)).astype(np.int32), device=a.device)
    a.reset()
    for t in range(20):
        a.step(acts[t])
    sd = a.state_dict()
    for t in range(20, 50):
        a.step(acts[t])
    final = {k: v.clone() for k, v in a.state_dict().items()}
    maps = a.obs_maps.clone()
    b = BatchedRadSearchEnv(**kw)
    b.load_state_dict(sd)
    for t in range(20, 50):
        b.step(acts[t])
    for k, v in b.state_dict().items():
        assert torch.equal(v, final[k]), k
    assert torch.equal(b.obs_maps, maps)


def test_sharding_invariance_global_env_ids():
    """Two handles owning [0,96) and [96,160) reproduce one handle owning [0,160)."""
    import torch
    from rad_team_b200.envs import BatchedRadSearchEnv

    kw = dict(n_agents=3, grid=32, max_steps=30, poly_min=1, poly_max=5, seed=14)
    whole = BatchedRadSearchEnv(n_envs=160, **kw)
    lo = BatchedRadSearchEnv(n_envs=96, env_id_base=0, **kw)
    hi = BatchedRadSearchEnv(n_envs=64, env_id_base=96, **kw)
    rng = np.random.default_rng(14)
    for env in (whole, lo, hi):
        env.reset()
    for t in range(30):
        act = torch.as_tensor(rng.integers(1, 5, size=(160, 3)).astype(np.int32), device=whole.device)
        whole.step(act); lo.step(act[:96].contiguous()); hi.step(act[96:].contiguous())
    for k in ("AGENT_XY", "LAST_COUNT", "VISIT", "INTENSITY", "WELFORD", "EDGES"):
        w = whole.buffer(k)
        assert torch.equal(w[:96], lo.buffer(k)) and torch.equal(w[96:], hi.buffer(k)), k
    assert torch.equal(whole.obs_maps[96:], hi.obs_maps)


@pytest.mark.parametrize("kw", [
    dict(n_envs=333, n_agents=3, grid=32, max_steps=40, poly_min=1, poly_max=5, seed=15),
    dict(n_envs=70, n_agents=8, grid=8, max_steps=300, seed=16),      # > 125 visited cells impossible (64 cells), many revisits
    dict(n_envs=50, n_agents=5, grid=48, max_steps=150, seed=17),     # long paths: record lists overflow the 1 KB prefix
    dict(n_envs=9, n_agents=1, grid=16, max_steps=20, seed=18),
])
def test_sparse_and_dense_rasterisers_agree(monkeypatch, kw):
    """Default: k_raster_sparse_lsu composes tiles from the visited-cell records and streams them out;
    RT_B200_RASTER=sparse_tma does the same with bulk (TMA) copies; RT_B200_RASTER=dense streams the dense
    maps (k_raster_vec).  step() and step(rasterize=False) + rasterize() must all give identical tensors,
    and identical to the oracle."""
    import torch
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    full = dict(O.DEFAULTS); full.update(kw)
    small = BatchedRadSearchEnv(**full)   # default at this size: k_raster_small (warp per env, zeros then patch)
    monkeypatch.setenv("RT_B200_RASTER", "sparse_lsu")
    sparse = BatchedRadSearchEnv(**full)  # the large-batch default, forced
    monkeypatch.setenv("RT_B200_RASTER", "dense")
    monkeypatch.setenv("RT_B200_EPB", "16")
    monkeypatch.setenv("RT_B200_RASTER_UNROLL", "4")
    dense = BatchedRadSearchEnv(**full)
    monkeypatch.delenv("RT_B200_RASTER"); monkeypatch.delenv("RT_B200_EPB"); monkeypatch.delenv("RT_B200_RASTER_UNROLL")
    monkeypatch.setenv("RT_B200_SPARSE_CTAS", "1")
    monkeypatch.setenv("RT_B200_RASTER", "sparse_lsu")
    split = BatchedRadSearchEnv(**full)
    monkeypatch.setenv("RT_B200_RASTER", "sparse_tma")
    tma = BatchedRadSearchEnv(**full)
    monkeypatch.delenv("RT_B200_RASTER")
    orc = O.OracleEnv(**full)
    E, A, T = full["n_envs"], full["n_agents"], full["max_steps"]
    rng = np.random.default_rng(kw["seed"])
    for env in (small, sparse, dense, split, tma):
        env.reset()
    _, m_o = orc.reset()
    assert np.array_equal(sparse.obs_maps.cpu().numpy(), m_o) and torch.equal(sparse.obs_maps, dense.obs_maps)
    for t in range(T):
        a_np = rng.integers(1, 5, size=(E, A)).astype(np.int32)
        if t % 3 == 0:  # drive agents along long straight paths so that many distinct cells get visited
            a_np[:] = 1 + (t // 30) % 4
        act = torch.as_tensor(a_np, device=sparse.device)
        small.step(act); sparse.step(act); dense.step(act); tma.step(act)
        split.step(act, rasterize=False); split.rasterize()
        _, m_o, _, _ = orc.step(a_np)
        assert torch.equal(sparse.obs_maps, small.obs_maps), t
        assert torch.equal(sparse.obs_maps, dense.obs_maps), t
        assert torch.equal(sparse.obs_maps, split.obs_maps), t
        assert torch.equal(sparse.obs_maps, tma.obs_maps), t
        assert np.array_equal(sparse.obs_maps.cpu().numpy(), m_o), t
    for k in ("VISIT", "INTENSITY", "WELFORD", "LOG", "CELLREC", "VREC"):
        assert torch.equal(sparse.buffer(k), dense.buffer(k)), k
    nv = sparse.buffer("VREC")[:, 0, 0].cpu().numpy()
    assert np.array_equal(nv, (sparse.buffer("VISIT").cpu().numpy() > 0).sum(1))
    if kw["grid"] == 48:
        assert nv.max() > 125  # the overflow path (records beyond the bulk-copied prefix) was exercised


def test_step_is_cuda_graph_capturable():
    """The device path is pure kernel launches on the caller's stream (no sync, no allocation), so a
    whole step can be captured once and replayed with new actions written into the same tensor."""
    import torch
    from rad_team_b200.envs import BatchedRadSearchEnv

    kw = dict(n_envs=500, n_agents=3, grid=32, max_steps=60, poly_min=1, poly_max=5, seed=31)
    eager, graphed = BatchedRadSearchEnv(**kw), BatchedRadSearchEnv(**kw)
    rng = np.random.default_rng(31)
    acts = torch.as_tensor(rng.integers(1, 5, size=(60, 500, 3)).astype(np.int32), device=eager.device)
    eager.reset(); graphed.reset()
    static_act = acts[0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):       # warm-up outside capture, as torch recommends
        graphed.step(static_act)
    torch.cuda.current_stream().wait_stream(side)
    eager.step(acts[0])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        graphed.step(static_act)
    # the capture itself does not execute: replay for steps 1..59
    for t in range(1, 60):
        static_act.copy_(acts[t])
        g.replay()
        eager.step(acts[t])
    torch.cuda.synchronize()
    assert torch.equal(eager.obs_maps, graphed.obs_maps) and torch.equal(eager.obs_xy, graphed.obs_xy)
    for k in ("VISIT", "INTENSITY", "WELFORD", "VREC", "TICK", "EP_RETURN"):
        assert torch.equal(eager.buffer(k), graphed.buffer(k)), k
    assert graphed.errors() == 0


def test_host_abi_with_pageable_buffers_and_host_mask():
    """rt_env_step_host / rt_env_reset_host called directly through ctypes with plain (pageable) NumPy
    arrays for every host buffer: the handle's pinned staging path."""
    import ctypes as C

    import torch
    from oracle import oracle as O
    from rad_team_b200 import _lib
    from rad_team_b200.envs import BatchedRadSearchEnv

    kw = dict(O.DEFAULTS); kw.update(n_envs=40, n_agents=2, grid=16, max_steps=30, poly_min=0, poly_max=2, rect_min=1,
                                     rect_max=3, seed=32)
    g, o = BatchedRadSearchEnv(**kw), O.OracleEnv(**kw)
    L, st = _lib.lib(), torch.cuda.current_stream().cuda_stream
    xy = np.zeros((40, 2, 2), np.int32); rew = np.zeros((40, 2), np.int32); done = np.zeros((40, 2), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(L.rt_env_reset_host(g._h, None, p(xy), g.obs_maps.data_ptr(), st))
    xy_o, m_o = o.reset()
    assert np.array_equal(xy, xy_o) and np.array_equal(g.obs_maps.cpu().numpy(), m_o)
    rng = np.random.default_rng(32)
    for t in range(20):
        act = rng.integers(1, 5, size=(40, 2)).astype(np.int32)
        _lib.check(L.rt_env_step_host(g._h, p(act), p(xy), p(rew), p(done), g.obs_maps.data_ptr(), st))
        xy_o, m_o, rew_o, done_o = o.step(act)
        assert np.array_equal(xy, xy_o) and np.array_equal(rew, rew_o) and np.array_equal(done, done_o)
        assert np.array_equal(g.obs_maps.cpu().numpy(), m_o)
        if t == 9:
            mask = (rng.random(40) < 0.4).astype(np.uint8)
            _lib.check(L.rt_env_reset_host(g._h, p(mask), p(xy), g.obs_maps.data_ptr(), st))
            xy_o, m_o = o.reset(mask=mask)
            assert np.array_equal(xy, xy_o) and np.array_equal(g.obs_maps.cpu().numpy(), m_o)
    act = np.ones((40, 2), np.int32); act[7, 1] = 9
    assert L.rt_env_step_host(g._h, p(act), p(xy), p(rew), p(done), None, st) == _lib.RT_ERR_BAD_ACTION
    o.step(act, maps=False)
    assert np.array_equal(g.buffer("TICK").cpu().numpy(), o.export("TICK"))  # env 7 voided, the rest advanced


def test_incremental_observation_update_matches_full_rasterisation():
    """Opt-in rt_env_set_incremental: handed the same tensor every step, the env only patches the cells a
    step changed; the tensor must stay bit-identical to a full rasterisation — across masked resets, a
    change of target tensor, invalid actions, agents sharing and swapping cells, and the host path."""
    import torch
    from rad_team_b200.envs import BatchedRadSearchEnv

    kw = dict(n_envs=400, n_agents=4, grid=8, max_steps=50, poly_min=0, poly_max=2, rect_min=1, rect_max=2, seed=51)
    full, inc = BatchedRadSearchEnv(**kw), BatchedRadSearchEnv(**kw)
    inc.set_incremental(True)
    rng = np.random.default_rng(51)
    full.reset(); inc.reset()
    other = torch.zeros_like(inc.obs_maps)
    l_full = l_inc = 0
    for t in range(120):
        a_np = rng.integers(1, 5, size=(400, 4)).astype(np.int32)
        if t % 11 == 5:
            a_np[rng.integers(0, 400), rng.integers(0, 4)] = 0                    # voided env
        act = torch.as_tensor(a_np, device=full.device)
        l0, l1 = full.launches, inc.launches
        full.step(act)
        if t % 17 == 9:                       # a different target tensor: must fall back to a full rasterisation
            inc.step(act, out_maps=other)
            assert torch.equal(other, full.obs_maps), t
            inc.rasterize()                   # bring the env's own tensor up to date again (full)
        elif t % 13 == 3:                     # host path
            try:
                inc.step(a_np)
            except ValueError:
                pass
        else:
            inc.step(act)
        l_full += full.launches - l0; l_inc += inc.launches - l1
        assert torch.equal(inc.obs_maps, full.obs_maps), t
        if t % 25 == 24:
            mask = torch.as_tensor((rng.random(400) < 0.5).astype(np.uint8), device=full.device)
            full.reset(mask=mask); inc.reset(mask=mask)
            assert torch.equal(inc.obs_maps, full.obs_maps)
        if t == 49 or t == 99:
            full.reset(); inc.reset()
    assert full.errors() == inc.errors()
    for k in ("VISIT", "INTENSITY", "VREC"):
        assert torch.equal(full.buffer(k), inc.buffer(k)), k


def test_handle_runs_on_its_creation_device():
    """include/rt_b200.h conventions: entry points run on the handle's device, not the caller's current one."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g, o = make_pair(device=1, n_envs=64, n_agents=2, grid=16, max_steps=12, seed=5)
    torch.cuda.set_device(0)
    run_episode(g, o, 8, np.random.default_rng(0))
    assert torch.cuda.current_device() == 0
    g.close()


def test_observation_memory_allocator():
    """rt_obs_alloc / rt_obs_free (include/rt_b200.h): compressible observation memory behaves like any
    device memory — the rasteriser writes the same bytes into it, torch reads it, and it outlives the
    Python handle that created it for as long as a tensor uses it."""
    import ctypes as C
    import gc

    import torch

    from rad_team_b200 import _lib
    from rad_team_b200.memory import compressible_tensor, obs_tensor

    g, o = make_pair(n_envs=96, n_agents=3, grid=32, max_steps=40, seed=11)
    assert g.obs_maps.is_cuda and hasattr(g.obs_maps, "rt_compressible")
    plain = torch.empty_like(g.obs_maps)          # ordinary caching-allocator memory
    mine = obs_tensor(g.obs_maps.shape, zero=False)
    g.reset(out_maps=mine)
    o.reset()
    rng = np.random.default_rng(3)
    for t in range(12):
        act = rng.integers(1, 5, size=(96, 3)).astype(np.int32)
        g.step(torch.as_tensor(act, device=g.device), out_maps=mine)
        g.rasterize(plain)
        _, m_o, _, _ = o.step(act, maps=True)
        assert torch.equal(mine, plain)
        assert np.array_equal(mine.cpu().numpy(), m_o)
    view = mine[5:7]
    del mine
    gc.collect()
    assert float(view.sum()) == float(plain[5:7].sum())   # the block is still mapped under the view
    del view
    gc.collect()
    ints = compressible_tensor((1000, 3), dtype=torch.int32)
    assert ints.dtype == torch.int32 and int(ints.sum()) == 0 and tuple(ints.shape) == (1000, 3)
    # C ABI contract: foreign pointers are refused, NULL is a no-op, zero bytes is a bad argument
    L = _lib.lib()
    assert L.rt_obs_free(C.c_void_p(plain.data_ptr())) == _lib.RT_ERR_BAD_ARG
    assert L.rt_obs_free(None) == _lib.RT_OK
    p = C.c_void_p()
    assert L.rt_obs_alloc(0, C.byref(p), None) == _lib.RT_ERR_BAD_ARG
    assert L.rt_obs_alloc(1 << 20, C.byref(p), None) == _lib.RT_OK and p.value
    assert L.rt_obs_free(p) == _lib.RT_OK
    assert L.rt_obs_free(p) == _lib.RT_ERR_BAD_ARG   # double free is caught, not executed


def test_ptrs_acceptance_threshold_adversarial():
    """The device sampler decides most PTRS acceptance tests from a cheap approximate evaluation with an error
    margin (rt_device.cuh: ptrs_squeeze) and only the rest from the exact fdlibm-style test the oracle runs.
    Injected (U, V) pairs put V at relative distances 1e-15 … 3e-1 on BOTH sides of the acceptance boundary
    (inside the margin: the exact test must take over; outside: the approximation decides) — every count has
    to equal the oracle's, and both outcomes have to occur."""
    import math

    import torch

    rng = np.random.default_rng(77)
    deltas = np.array([s * 10.0 ** -p for p in (15, 13, 11, 9, 7, 6, 5, 4, 3, 2, 1) for s in (-1, 1)] + [-0.3, 0.3, 0.0])
    n_lam = 96
    E = n_lam * len(deltas)
    g, o = make_pair(n_envs=E, n_agents=1, grid=32, max_steps=8, seed=5, fixed_scenario=1)
    # the agent sits under the top wall and pushes UP: it never moves, so lambda is the same every step
    dist = rng.integers(1, 12, size=n_lam)
    src = np.stack([np.full(n_lam, 16), 31 - dist], 1).repeat(len(deltas), 0)
    start = np.tile(np.array([[16, 31]]), (E, 1)).reshape(E, 1, 2)
    kw = dict(src_xy=src, src_int=(10 ** rng.uniform(5.5, 7.5, n_lam)).repeat(len(deltas)),
              bkg=rng.uniform(10, 51, n_lam).repeat(len(deltas)), start_xy=start, n_poly=np.zeros(E, np.int32))
    g.set_scenario(**kw); o.set_scenario(**kw)
    g.reset(); o.reset()
    up = np.ones((E, 1), np.int32)
    g.step(torch.as_tensor(up, device=g.device)); o.step(up, maps=False)
    lam = g.buffer("LAST_LAMBDA").cpu().numpy()[:, 0]
    assert np.array_equal(lam, o.export("LAST_LAMBDA")[:, 0]) and (lam >= 10).all()

    u = np.zeros((E, 1, 8))
    k_tail = np.zeros(E, np.int64)
    for e in range(E):
        L = float(lam[e])
        slam, loglam = math.sqrt(L), math.log(L)
        b = 0.931 + 2.53 * slam
        a = -0.059 + 0.02483 * b
        invalpha = 1.1239 + 1.1328 / (b - 3.4)
        usel = 0.5 + (0.44 + 0.05 * ((e // len(deltas)) % 7) / 7.0) * (1 if e % 2 else -1)  # us in (0.01, 0.06]: no fast accept
        U = usel - 0.5
        us = 0.5 - abs(U)
        k = math.floor((2 * a / us + b) * U + L + 0.43)
        k_tail[e] = k
        if k < 0:  # rejected before the acceptance test: nothing to aim at
            u[e, 0, :4] = [usel, 0.5, 0.5, 0.0]
            continue
        rhs = -L + k * loglam - math.lgamma(k + 1.0)
        vstar = math.exp(rhs - math.log(invalpha) + math.log(a / (us * us) + b))
        V = min(max(vstar * (1.0 + deltas[e % len(deltas)]), 0.0), 1.0 - 2.0 ** -53)
        if us < 0.013:
            V = min(V, us)  # stay clear of the early rejection (us < 0.013 and V > us)
        # pair 1 = the adversarial draw; pair 2 accepts at once (us = 0.5, V = 0) with k = floor(lam + 0.43)
        u[e, 0, :4] = [usel, V, 0.5, 0.0]
    g.inject_uniforms(u); o.inject_uniforms(u)
    g.step(torch.as_tensor(up, device=g.device)); o.step(up, maps=False)
    cnt = g.buffer("LAST_COUNT").cpu().numpy()[:, 0]
    want = o.export("LAST_COUNT")[:, 0]
    assert np.array_equal(cnt, want), np.argwhere(cnt != want)[:8]
    valid = k_tail >= 0
    accepted = (cnt == k_tail) & valid
    fallback = (cnt == np.floor(lam + 0.43)) & valid & ~accepted
    assert accepted.sum() > E // 4 and fallback.sum() > E // 4, (accepted.sum(), fallback.sum())
    assert g.errors() == o.errors() == 0
    g.inject_uniforms(None)


@pytest.mark.timeout(120)
def test_corrupted_bucket_chain_cannot_hang_the_kernel():
    """A state restored through rt_env_buffer may be garbage; the bucket walk is bounded by the count it reads,
    so a cyclic chain with the largest possible count costs a few thousand hops, not an endless loop."""
    import torch
    from rad_team_b200.envs import BatchedRadSearchEnv

    g = BatchedRadSearchEnv(n_envs=64, n_agents=2, grid=2, max_steps=400, seed=3, poly_min=0, poly_max=0)
    g.reset()
    act = torch.ones((64, 2), dtype=torch.int32, device=g.device)
    for _ in range(20):
        g.step(act)
    log = g.buffer("LOG")                                   # int32 [cap, E, 8]
    ids = torch.arange(1, log.shape[0] + 1, dtype=torch.int32, device=g.device).view(-1, 1)
    log[:, :, 7] = ids | (0x7FFF << 16)                     # every bucket links to itself and claims 32,767 readings
    for _ in range(3):
        g.step(act)
    torch.cuda.synchronize()
    assert g.errors() & ~0x1F == 0
    g.close()


@pytest.mark.parametrize("zc_out_kb,layout", [("256", "packed"), ("0", "packed"), ("0", "separate"), ("256", "separate"),
                                              ("0", "pageable"), ("256", "mixed")])
def test_host_step_result_paths_agree(monkeypatch, zc_out_kb, layout):
    """rt_env_step_host returns obs_xy / reward / done three ways — stored by k_step itself into mapped host memory
    (small batches), ONE copy-engine request when the caller's buffers are one pinned block laid out xy | reward |
    done, three requests otherwise — and through pinned staging for pageable buffers.  All equal to the oracle,
    a bad action is reported (and voids exactly its env) on every path."""
    import ctypes as C

    import torch
    from oracle import oracle as O
    from rad_team_b200 import _lib
    from rad_team_b200.envs import BatchedRadSearchEnv

    monkeypatch.setenv("RT_B200_ZC_OUT_KB", zc_out_kb)
    E, A = 300, 3
    kw = dict(O.DEFAULTS); kw.update(n_envs=E, n_agents=A, grid=32, max_steps=40, poly_min=1, poly_max=4, seed=77)
    g, o = BatchedRadSearchEnv(**kw), O.OracleEnv(**kw)
    L, st, n = _lib.lib(), torch.cuda.current_stream().cuda_stream, E * A
    if layout == "packed":
        block = torch.zeros(n * 13, dtype=torch.uint8).pin_memory()
        xy = block[:n * 8].view(torch.int32).view(E, A, 2)
        rew = block[n * 8:n * 12].view(torch.int32).view(E, A)
        done = block[n * 12:].view(E, A)
    elif layout == "separate":
        xy = torch.zeros((E, A, 2), dtype=torch.int32).pin_memory()
        rew = torch.zeros((E, A), dtype=torch.int32).pin_memory()
        done = torch.zeros((E, A), dtype=torch.uint8).pin_memory()
    elif layout == "pageable":
        xy = torch.zeros((E, A, 2), dtype=torch.int32)
        rew = torch.zeros((E, A), dtype=torch.int32)
        done = torch.zeros((E, A), dtype=torch.uint8)
    else:  # pinned positions, pageable reward, pinned done
        xy = torch.zeros((E, A, 2), dtype=torch.int32).pin_memory()
        rew = torch.zeros((E, A), dtype=torch.int32)
        done = torch.zeros((E, A), dtype=torch.uint8).pin_memory()
    h_act = torch.zeros((E, A), dtype=torch.int32).pin_memory()
    g.reset()
    o.reset()
    rng = np.random.default_rng(77)
    for t in range(25):
        a_np = rng.integers(1, 5, size=(E, A)).astype(np.int32)
        bad = t == 11
        if bad:
            a_np[17, 1] = 0
        h_act.copy_(torch.from_numpy(a_np))
        rc = L.rt_env_step_host(g._h, h_act.data_ptr(), xy.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                g.obs_maps.data_ptr(), st)
        assert rc == (_lib.RT_ERR_BAD_ACTION if bad else _lib.RT_OK), (t, rc)
        if bad:  # the oracle steps every env but 17, whose step the reference would have refused before any change
            keep = o.export("AGENT_XY")[17].copy()
            a_fix = a_np.copy(); a_fix[17, 1] = 1
            xy_o, m_o, rew_o, done_o = o.step(a_fix)
            sel = np.ones(E, bool); sel[17] = False
            assert np.array_equal(xy.numpy()[sel], xy_o[sel]) and np.array_equal(rew.numpy()[sel], rew_o[sel])
            assert np.array_equal(xy.numpy()[17], keep) and (rew.numpy()[17] == 0).all() and (done.numpy()[17] == 0).all()
            break  # env 17 now differs from the oracle's: the comparison ends here
        xy_o, m_o, rew_o, done_o = o.step(a_np)
        assert np.array_equal(xy.numpy(), xy_o) and np.array_equal(rew.numpy(), rew_o), t
        assert np.array_equal(done.numpy(), done_o), t
        assert np.array_equal(g.obs_maps.cpu().numpy(), m_o), t
    assert g.errors() == _lib.FLAG_BAD_ACTION


def test_cell_records_survive_resets_by_episode_tag_including_the_16_bit_wrap():
    """A reset does not clear the per-cell records: they carry the low 16 bits of the episode number and a record of
    an older episode counts as empty; when the tag wraps (episode 65,536) the records ARE cleared.  Checked against a
    host-side tally of the agents' paths — visit counts per cell and the number of distinct visited cells — across
    episodes 65,533 ... 65,538, with and without the dense maps."""
    import torch
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    kw = dict(O.DEFAULTS); kw.update(n_envs=24, n_agents=2, grid=8, max_steps=40, seed=99)
    env = BatchedRadSearchEnv(**kw)
    E, A, G, T = 24, 2, 8, 40
    env.buffer("EPISODE").fill_(65532)          # the next reset opens episode 65,533
    rng = np.random.default_rng(99)
    for ep in range(65533, 65539):
        xy, _ = env.reset()
        assert int(env.buffer("EPISODE")[0].item()) == ep
        tally = np.zeros((E, G * G), np.int64)
        for t in range(T):
            act = torch.as_tensor(rng.integers(1, 5, size=(E, A)).astype(np.int32), device=env.device)
            xy, _, _, _, _ = env.step(act)
            p = xy.cpu().numpy()
            for a in range(A):
                np.add.at(tally, (np.arange(E), p[:, a, 1] * G + p[:, a, 0]), 1)
            if t in (0, 7, T - 1):
                assert np.array_equal(env.buffer("VISIT").cpu().numpy().astype(np.int64), tally), (ep, t)
                assert np.array_equal(env.buffer("VREC")[:, 0, 0].cpu().numpy(), (tally > 0).sum(1)), (ep, t)
                lut = env.buffer("VISIT_LUT").cpu().numpy()
                maps = env.obs_maps.cpu().numpy().reshape(E, 3, G * G)
                assert np.array_equal(maps[:, 2], lut[tally]), (ep, t)     # the visit channel = table[tally]
        assert env.errors() == 0


def test_get_state_set_state_replays_bit_exactly():
    """rt_env_get_state / rt_env_set_state: snapshot mid-episode, run on, restore, run the same actions again — every
    output and the final state are identical; a blob of another shape is refused; works through host memory too."""
    import torch
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    kw = dict(O.DEFAULTS); kw.update(n_envs=130, n_agents=3, grid=32, max_steps=60, poly_min=1, poly_max=5, seed=5)
    env = BatchedRadSearchEnv(**kw)
    rng = np.random.default_rng(5)
    acts = torch.as_tensor(rng.integers(1, 5, size=(50, 130, 3)).astype(np.int32), device=env.device)
    env.reset()
    for t in range(20):
        env.step(acts[t])
    blob = env.get_state()
    host_blob = blob.cpu()                          # pageable host memory

    def run():
        outs = []
        for t in range(20, 50):
            xy, rw, dn, _, _ = env.step(acts[t])
            outs.append((xy.clone(), rw.clone(), dn.clone(), env.buffer("LAST_COUNT").clone(), env.obs_maps.clone()))
        return outs, env.get_state()

    first, end1 = run()
    env.set_state(blob)
    second, end2 = run()
    env.set_state(host_blob)
    third, end3 = run()
    for a, b, c in zip(first, second, third):
        for x, y, z in zip(a, b, c):
            assert torch.equal(x, y) and torch.equal(x, z)
    assert torch.equal(end1, end2) and torch.equal(end1, end3)
    other = BatchedRadSearchEnv(**dict(kw, n_envs=131))
    with pytest.raises(ValueError):
        other.set_state(blob[: other.get_state().numel()] if blob.numel() >= other.get_state().numel() else blob)


@pytest.mark.parametrize("E,A,raster", [(4096, 1, "auto"), (1024, 3, "auto"), (4096, 1, "sparse_lsu")])
def test_whole_episode_cuda_graph_replay_matches_eager_launches(monkeypatch, E, A, raster):
    """A captured CUDA graph of a whole episode (120 x (k_step + rasteriser) + reset), replayed back to back — no host
    gaps, programmatic dependent launches overlapping every kernel with its neighbours — must leave exactly the state,
    readings and observation of the same episode launched eagerly.  (Small batches use k_raster_small, which writes
    its zeros BEFORE its dependency wait: this is the test that would catch a rasteriser running ahead of the
    previous step's.)"""
    import torch
    from oracle import oracle as O
    from rad_team_b200.envs import BatchedRadSearchEnv

    monkeypatch.setenv("RT_B200_RASTER", raster)
    kw = dict(O.DEFAULTS); kw.update(n_envs=E, n_agents=A, grid=32, max_steps=120, poly_min=0, poly_max=3, seed=61)
    a, b = BatchedRadSearchEnv(**kw), BatchedRadSearchEnv(**kw)
    T = 120
    gen = torch.Generator(device=a.device); gen.manual_seed(3)
    acts = torch.randint(1, 5, (T, E, A), dtype=torch.int32, device=a.device, generator=gen)
    ra = torch.zeros((T, E, A), dtype=torch.float32, device=a.device)
    rb = torch.zeros_like(ra)
    a.set_rollout(ra); b.set_rollout(rb)
    mid_a = torch.empty_like(a.obs_maps); mid_b = torch.empty_like(b.obs_maps)

    def episode(env, mid):
        env.reset()
        for t in range(T):
            env.step(acts[t])
            if t == 59:
                mid.copy_(env.obs_maps)   # a consumer reading the tensor between two steps

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        episode(b, mid_b)                 # warm-up outside the capture
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            episode(b, mid_b)
        for _ in range(3):                # episodes 1, 2, 3 of env b (episode 0 was the warm-up... see below)
            g.replay()
        side.synchronize()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(4):                    # env a: the same four episodes, eagerly
        episode(a, mid_a)
    torch.cuda.synchronize()
    assert torch.equal(a.buffer("EPISODE"), b.buffer("EPISODE"))
    assert torch.equal(mid_a, mid_b)
    assert torch.equal(a.obs_maps, b.obs_maps) and torch.equal(ra, rb)
    for k in ("VREC", "CELLREC", "WELFORD", "LOG", "AGENT_XY", "EST_MAXMIN"):
        assert torch.equal(a.buffer(k), b.buffer(k)), k
    assert a.errors() == b.errors() == 0
