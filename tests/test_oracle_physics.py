"""Oracle pieces the mounted reference does not contain (parity unpinned by the reference): pinned
instead to third-party known answers — Random123 Philox vectors, libm, numpy.random.Generator.poisson —
and to the written spec.  CPU only."""
import math

import numpy as np
import pytest

from oracle import oracle as O


def test_philox_known_answers():  # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in O.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in O.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == [
        "0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in O.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == [
        "0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_uniform_stream_layout():
    u = O.uniforms(seed=0x1234567800000042, gid=(5 << 32) | 7, stream=1, agent=2, tick=9, episode=3, n=6)
    assert np.all((u >= 0) & (u < 1))
    for k in range(6):
        o = O.philox([k >> 1, 2 | (1 << 8) | (9 << 16), 3, 7], [0x42, 0x12345678 ^ 5])
        lo, hi = int(o[2 * (k & 1)]), int(o[2 * (k & 1) + 1])
        assert u[k] == ((hi << 32 | lo) >> 11) * 2.0 ** -53


def _ulp_err(got, ref):
    return abs(got - ref) / math.ulp(ref) if ref != 0 else abs(got - ref)


def test_log_exp_within_one_ulp_of_libm():
    L = O.lib()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.random(20000), rng.random(20000) * 1e3, 10 ** rng.uniform(-300, 300, 20000),
                         1 + rng.normal(0, 1e-3, 20000), [1.0, 2.0, 0.5, 10.0, 5e-324, 1e-310]])
    assert max(_ulp_err(L.orc_log(float(x)), math.log(x)) for x in xs if x > 0) <= 1.0
    assert L.orc_log(0.0) == -math.inf and L.orc_log(1.0) == 0.0
    xs = -rng.random(40000) * 12
    assert max(_ulp_err(L.orc_exp(float(x)), math.exp(x)) for x in xs) <= 1.0
    assert L.orc_exp(0.0) == 1.0
    for x in (1.0, 2.0, 3.0, 7.5, 100.0, 12345.0):
        assert L.orc_loggam(x) == pytest.approx(math.lgamma(x), rel=1e-10, abs=1e-12)


def test_poisson_matches_numpy_draw_for_draw():
    """Inject u_k = PCG64.random(); the count must equal a twin generator's .poisson(lam) and the two
    streams must stay in lock-step (same number of uniforms consumed)."""
    rng = np.random.default_rng(2)
    lams = list(10 ** rng.uniform(-2, 9, 150)) + [0.0, 9.999, 10.0, 10.5, 1e9, 5.0, 1.0, 30.0, 1e4]
    n = 0
    for lam in lams:
        for s in range(40):
            seed = s * 7919 + int(lam * 13) % 1000
            g, twin = np.random.Generator(np.random.PCG64(seed)), np.random.Generator(np.random.PCG64(seed))
            u = g.random(256)
            want = twin.poisson(lam)
            got, used, flags = O.poisson(lam, u)
            assert (got, flags) == (want, 0), lam
            assert twin.random() == u[used]  # lock-step
            n += 1
    assert n > 6000


def test_poisson_guards():
    assert O.poisson(-1.0, np.zeros(4))[2] & 0x04
    assert O.poisson(float("nan"), np.zeros(4))[2] & 0x04
    assert O.poisson(2.0 ** 31, np.zeros(4))[2] & 0x04
    got, used, flags = O.poisson(50.0, np.array([0.5]))  # PTRS needs two uniforms
    assert flags & 0x08


def test_segment_edge_half_open_rule():
    e = [2, 0, 2, 4]  # vertical edge x = 2, y in [0, 4]
    assert O.segment_hits([0.5, 1.5, 3.5, 1.5], e)
    assert not O.segment_hits([0.5, 5.5, 3.5, 5.5], e)     # passes above
    assert not O.segment_hits([0.5, 1.5, 1.5, 1.5], e)     # stops short
    assert not O.segment_hits([2.0, 5.0, 2.0, 9.0], e)     # collinear, disjoint
    assert not O.segment_hits([0.5, 0.5, 0.5, 0.5], e)     # degenerate segment
    assert not O.segment_hits([0.5, 1.5, 3.5, 1.5], [0, 0, 0, 0])  # empty edge slot
    # a ray through the shared corner (2,2) of a rectangle [2,4]x[2,4]: exactly one of the two
    # edges meeting there counts, so the obstruction is seen exactly as for a ray just inside it
    bottom, left = [2, 2, 4, 2], [2, 4, 2, 2]
    hits = O.segment_hits([0.5, 0.5, 3.5, 3.5], bottom) + O.segment_hits([0.5, 0.5, 3.5, 3.5], left)
    assert hits == 1


def test_scenario_generation_invariants():
    env = O.OracleEnv(n_envs=512, n_agents=3, grid=32, max_steps=120, poly_min=1, poly_max=5, seed=9)
    env.reset()
    src, start, npoly, edges = (env.export(k) for k in ("SRC_XY", "START_XY", "N_POLY", "EDGES"))
    assert src.min() >= 0 and src.max() < 32 and start.min() >= 0 and start.max() < 32
    assert np.all(np.abs(start - src[:, None, :]).sum(-1) > 1)
    assert npoly.min() >= 1 and npoly.max() <= 5 and len(set(npoly.tolist())) == 5
    s_int, bkg = env.export("SRC_INT"), env.export("BKG")
    assert np.all((s_int >= 1e6) & (s_int < 1e7)) and np.all((bkg >= 10) & (bkg < 51))
    for e in range(512):
        for p in range(5):
            q = edges[e, 4 * p:4 * p + 4]
            if p >= npoly[e] or not q.any():
                assert p < npoly[e] or not q.any()
                continue
            x0, y0, x1, y1 = q[0, 0], q[0, 1], q[1, 2], q[1, 3]
            assert 2 <= x1 - x0 <= 6 and 2 <= y1 - y0 <= 6 and x0 >= 0 and y1 <= 32
            for c in [src[e]] + list(start[e]):
                assert not (x0 <= c[0] < x1 and y0 <= c[1] < y1)
    # a new episode draws a new scenario; the same (seed, env, episode) reproduces it
    env.reset()
    assert not np.array_equal(src, env.export("SRC_XY"))
    twin = O.OracleEnv(n_envs=512, n_agents=3, grid=32, max_steps=120, poly_min=1, poly_max=5, seed=9)
    twin.reset(); twin.reset()
    assert np.array_equal(env.export("EDGES"), twin.export("EDGES"))
    # sharding invariance: envs [100,164) of the big batch == a 64-env oracle with env_id_base=100
    shard = O.OracleEnv(n_envs=64, n_agents=3, grid=32, max_steps=120, poly_min=1, poly_max=5, seed=9, env_id_base=100)
    shard.reset(); shard.reset()
    assert np.array_equal(env.export("EDGES")[100:164], shard.export("EDGES"))
    assert np.array_equal(env.export("START_XY")[100:164], shard.export("START_XY"))


def test_obstruction_blocks_motion_and_line_of_sight():
    env = O.OracleEnv(n_envs=1, n_agents=1, grid=8, max_steps=50, fixed_scenario=1, transmission=0.5,
                      cell_pitch_cm=100.0)
    edges = np.zeros((1, 20, 4), np.float32)
    edges[0, :4] = [[3, 0, 4, 0], [4, 0, 4, 6], [4, 6, 3, 6], [3, 6, 3, 0]]  # wall x in [3,4], y in [0,6]
    env.set_scenario(src_xy=[6, 2], src_int=[4e6], bkg=[0.0], start_xy=[2, 2], n_poly=[1], edges=edges)
    env.reset()
    xy, _, rew, done = env.step([[2]])  # RIGHT into the wall: refused, distance unchanged => -1
    assert tuple(xy[0, 0]) == (2, 2) and rew[0, 0] == -1
    assert env.export("LAST_LOS")[0, 0] == 1
    assert env.export("LAST_LAMBDA")[0, 0] == 4e6 * 0.5 / (4 * 100.0) ** 2
    for a in (1, 1, 1, 1, 1):  # climb above the wall (y = 7 > 6)
        xy, *_ = env.step([[a]])
    assert tuple(xy[0, 0]) == (2, 7)
    xy, *_ = env.step([[2]])
    assert tuple(xy[0, 0]) == (3, 7)
    # from (3,7) the ray to (6,2) clips the wall's top-right region? check against the predicate directly
    los = env.export("LAST_LOS")[0, 0]
    seg = [3.5, 7.5, 6.5, 2.5]
    want = int(any(O.segment_hits(seg, e) for e in edges[0, :4]))
    assert los == want
