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
