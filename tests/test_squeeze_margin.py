"""Error budget of the PTRS squeeze (rad_team_b200/csrc/rt_device.cuh: ptrs_squeeze), checked on the CPU.

The device decides most acceptance tests of the Poisson sampler from an approximate evaluation (single-precision
lg2.approx logarithms, three Stirling terms) and a margin; the decision must never differ from the exact test's.
This file restates the device formula in NumPy with the logarithm replaced by an ADVERSARY: the true value moved
by the full hardware bound the budget assumes (2^-21 (1 + |ln x|), sign chosen against us) — and checks on a
few hundred thousand random slow-path cases that a decision is only ever taken when it is the exact one, and how
rarely the exact test is still needed.  The device code itself is pinned against the oracle on the GPU
(tests/test_gpu_env_parity.py::test_ptrs_acceptance_threshold_adversarial and every count comparison there)."""
import math

import numpy as np

C = 2.0 ** -20
HW = 2.0 ** -21  # |flog(x) - ln x| <= HW * (1 + |ln x|): float rounding of x plus lg2.approx.f32's documented error


def squeeze(V, us, a, b, invalpha, lam, loglam, k, adv):
    """The device formula; `adv` in {-1, +1}^4 pushes each approximate logarithm to the edge of its bound."""
    x = k + 1.0
    flog = lambda v, s: math.log(v) + s * HW * (1 + abs(math.log(v)))
    usf = np.float32(us)
    wf = float(np.float32(a) / (usf * usf) + np.float32(b))
    Lv, Li, Lw = flog(V, adv[0]), flog(invalpha, adv[1]), flog(wf, adv[2])
    if k < 7:
        G = math.log(math.factorial(k))
        gerr, gmag = 0.0, G
    else:
        Lq = flog(float(np.float32(x) / np.float32(lam)), adv[3])
        Lx = loglam + Lq
        ix = float(np.float32(1.0) / np.float32(x))
        G = (x - 0.5) * Lx - x + 0.9189385332046727 + (ix / 12.0 - ix ** 3 / 360.0)
        gerr = C * x * (1.0 + abs(Lq))
        gmag = x * abs(Lx) + x
    kl = k * loglam
    diff = (Lv + Li - Lw) - ((kl - lam) - G)
    logs = abs(Lv) + abs(Li) + abs(Lw) + 3.0
    mags = lam + abs(kl) + gmag + logs
    margin = 4.0 * (C * logs + gerr + 2.0 ** -23) + 2.0 ** -40 * mags
    if not margin < 1.0:
        return 0, diff, margin
    return (1 if diff < -margin else -1 if diff > margin else 0), diff, margin


def test_squeeze_never_contradicts_the_exact_test():
    rng = np.random.default_rng(2024)
    n_slow = n_decided = 0
    worst = 0.0
    for lam in np.concatenate([10 ** rng.uniform(1.0, 4.5, 400), [10.0, 10.5, 1e5, 1e6]]):
        lam = float(lam)
        slam, loglam = math.sqrt(lam), math.log(lam)
        b = 0.931 + 2.53 * slam
        a = -0.059 + 0.02483 * b
        invalpha = 1.1239 + 1.1328 / (b - 3.4)
        vr = 0.9277 - 3.6224 / (b - 2.0)
        for _ in range(600):
            U, V = rng.random() - 0.5, rng.random()
            us = 0.5 - abs(U)
            k = math.floor((2 * a / us + b) * U + lam + 0.43)
            if (us >= 0.07 and V <= vr) or k < 0 or (us < 0.013 and V > us) or V == 0.0:
                continue
            n_slow += 1
            exact = (math.log(V) + math.log(invalpha) - math.log(a / (us * us) + b)) - \
                    (-lam + k * loglam - math.lgamma(k + 1.0))
            for adv in ((1, 1, -1, -1), (-1, -1, 1, 1), (1, -1, 1, -1), (-1, 1, -1, 1)):
                dec, diff, margin = squeeze(V, us, a, b, invalpha, lam, loglam, k, adv)
                if diff is not None and margin < 1.0:
                    worst = max(worst, abs(diff - exact) / margin)
                if dec > 0:
                    assert exact < 0.0, (lam, U, V, k, exact, diff, margin)
                elif dec < 0:
                    assert exact > 0.0, (lam, U, V, k, exact, diff, margin)
            n_decided += dec != 0
    assert n_slow > 20000
    assert worst < 0.5            # the adversarial error uses less than half of the margin
    assert n_decided > 0.9 * n_slow  # over lambda in [10, 3e4]; > 99 % below 1e3 (the margin grows with lambda)
