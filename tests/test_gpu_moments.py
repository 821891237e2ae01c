"""K4b — batch running moments (warp-shuffle Chan merges) and the standardise pass, vs the oracle's
sequential reference recurrence.  Chan merging differs from the sequential pass only by rounding, so
the tolerance is 1e-10 relative on the moments (well inside the 1e-5 of the north star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 1000, 4097, 262144 + 3, 5_000_001])
def test_moments_match_sequential_reference(n):
    import torch
    from oracle import oracle as O
    from rad_team_b200.math_utils import RunningMoments

    rng = np.random.default_rng(n)
    x = rng.poisson(200.0, n).astype(np.float32) + (rng.random(n) < 0.01) * 5e4
    x = x.astype(np.float32)
    m = RunningMoments()
    xd = torch.as_tensor(x, device="cuda")
    m.update(xd)
    got = m.state.cpu().numpy()
    want = O.moments_update(np.zeros(3), x)
    assert got[0] == want[0] == n
    np.testing.assert_allclose(got[1:], want[1:], rtol=1e-10, atol=1e-9)
    # deterministic: the same buffer gives the same bits
    m2 = RunningMoments(); m2.update(xd)
    assert np.array_equal(m2.state.cpu().numpy(), got)
    # a second batch folds into the running state
    y = (rng.random(max(n // 2, 1)) * 1e3).astype(np.float32)
    m.update(torch.as_tensor(y, device="cuda"))
    want = O.moments_update(want, y)
    got = m.state.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-9)
    # standardise: (x - mean) / max(std, 1), fp32 out
    z = m.standardize(xd).cpu().numpy()
    np.testing.assert_allclose(z, O.moments_standardize(got, x), rtol=1e-5, atol=1e-6)


def test_rank_ordered_merge_matches_oracle_and_is_order_sensitive_only_in_rounding():
    import torch
    from oracle import oracle as O
    from rad_team_b200.math_utils import RunningMoments

    rng = np.random.default_rng(0)
    parts = [rng.poisson(lam, 10000 + 17 * i).astype(np.float32) for i, lam in enumerate((5, 50, 500, 5000))]
    triples = []
    for p in parts:
        m = RunningMoments(); m.update(torch.as_tensor(p, device="cuda"))
        triples.append(m.state.clone())
    gathered = torch.stack(triples).contiguous()
    merged = RunningMoments(); merged.merge_gathered(gathered)
    got = merged.state.cpu().numpy()
    want = np.zeros(3)
    for t in triples:
        want = O.moments_merge(want, t.cpu().numpy())
    assert np.array_equal(got, want)  # same Chan formula, same order => same bits
    seq = O.moments_update(np.zeros(3), np.concatenate(parts))
    np.testing.assert_allclose(got, seq, rtol=1e-10)
    # an empty rank contributes nothing
    g2 = torch.cat([gathered[:2], torch.zeros(1, 3, dtype=torch.float64, device="cuda"), gathered[2:]]).contiguous()
    m3 = RunningMoments(); m3.merge_gathered(g2)
    assert np.array_equal(m3.state.cpu().numpy(), got)


@pytest.mark.parametrize("off_in,off_out", [(1, 1), (2, 2), (3, 3), (1, 0), (0, 2), (3, 1)])
def test_moments_accept_any_contiguous_float_view(off_in, off_out):
    """A row readings[t] of a [T, E, A] buffer with E*A % 4 != 0, or x[1:], is not 16-byte aligned: same results."""
    import torch
    from oracle import oracle as O
    from rad_team_b200.math_utils import RunningMoments

    rng = np.random.default_rng(off_in * 7 + off_out)
    n = 10007
    x = rng.poisson(120.0, n + 4).astype(np.float32)
    xd = torch.as_tensor(x, device="cuda")
    view = xd[off_in:off_in + n]
    assert view.data_ptr() % 16 == 4 * off_in
    m = RunningMoments()
    m.update(view)
    got = m.state.cpu().numpy()
    want = O.moments_update(np.zeros(3), x[off_in:off_in + n])
    assert got[0] == n
    np.testing.assert_allclose(got[1:], want[1:], rtol=1e-10, atol=1e-9)
    out = torch.full((n + 4,), -7.0, dtype=torch.float32, device="cuda")
    z = m.standardize(view, out=out[off_out:off_out + n]).cpu().numpy()
    np.testing.assert_allclose(z, O.moments_standardize(got, x[off_in:off_in + n]), rtol=1e-5, atol=1e-6)
    o = out.cpu().numpy()
    assert (o[:off_out] == -7.0).all() and (o[off_out + n:] == -7.0).all()  # nothing written outside the view
    # tiny buffers: head / tail only
    for k in (1, 2, 3, 5):
        mk = RunningMoments()
        mk.update(xd[off_in:off_in + k])
        np.testing.assert_allclose(mk.state.cpu().numpy(), O.moments_update(np.zeros(3), x[off_in:off_in + k]),
                                   rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("backend", ["peer", "nccl"])
def test_rollout_exchange_keeps_a_separate_global_state(backend):
    """One rank, three rollouts: global <- chan(global, this rollout), the per-rollout accumulator is cleared, every
    reading is counted once (the round-1 scheme re-merged the running state and double-counted at world > 1)."""
    import torch
    from oracle import oracle as O
    from rad_team_b200.distributed import RolloutStatsExchange
    from rad_team_b200.math_utils import RunningMoments

    rng = np.random.default_rng(3)
    ex = RolloutStatsExchange(backend=backend)
    assert ex.backend == backend and ex.world == 1
    local, glob = RunningMoments(), RunningMoments()
    want = np.zeros(3)
    total = 0.0
    for r in range(3):
        x = rng.poisson(50.0 * (r + 1), 5000 + r).astype(np.float32)
        local.update(torch.as_tensor(x, device="cuda"))
        delta = local.state.cpu().numpy().copy()
        ep = torch.tensor([-3.0 * r, 120.0, float(r)], dtype=torch.float64, device="cuda")
        merged = ex.exchange(local, glob, ep)
        assert np.array_equal(merged.cpu().numpy(), ep.cpu().numpy())
        assert np.array_equal(ex.gathered.cpu().numpy()[0], np.concatenate([delta, ep.cpu().numpy()]))
        assert np.array_equal(local.state.cpu().numpy(), np.zeros(3))
        want = O.moments_merge(want, O.moments_merge(np.zeros(3), delta))
        total += x.size
        assert np.array_equal(glob.state.cpu().numpy(), want)
    assert glob.state[0].item() == total and ex.errors() == 0
    with pytest.raises(ValueError):
        ex.exchange(glob, glob, None)  # one handle cannot be both accumulator and running state
    ex.close()


def test_rollout_exchange_is_cuda_graph_capturable():
    import torch
    from rad_team_b200.distributed import RolloutStatsExchange
    from rad_team_b200.math_utils import RunningMoments

    ex = RolloutStatsExchange(backend="peer")
    local, glob = RunningMoments(), RunningMoments()
    x = torch.arange(4096, dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        local.update(x)
        ex.exchange(local, glob, None)  # warm
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            local.update(x)
            ex.exchange(local, glob, None)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert glob.state[0].item() == 4096 * 4  # the warm-up call + 3 replays (capture itself executes nothing)
