"""CPU check of the argument behind the float32-exact fast division of k_jac_project (aar_jacobian.cuh: div_xy + DivGuard).

The kernel replaces float(double(X / Z)) (multicam_mapper.cpp:644-648) by float(q) for a double q within a few ulps of the
correctly rounded quotient Q, and keeps the result only if the guard accepts q: the low 29 mantissa bits of q are more than
8 ulps away from a float32 rounding boundary (0x10000000) and |float(q)| is in [2^-125, FLT_MAX].  The claim tested here, with
the guard restated in numpy integer arithmetic exactly as the device computes it: for every accepted q, every double within
3 ulps of q (so in particular Q, for which |q - Q| <= 2.6 ulps is the analytic bound and 1 ulp the measured maximum,
profiles/r1_divcheck.txt) rounds to the same float32.
"""
import numpy as np


def guard_accepts(q):
    """DivGuard::see + ok(), per quotient: near code on the low word, range code on the float32 bits."""
    bits = q.view(np.uint64)
    lo = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    near = (lo * np.uint64(8) + np.uint64(0x80000040)) & np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        f = q.astype(np.float32)
    fb = f.view(np.uint32).astype(np.uint64)
    rng = (fb * np.uint64(2) - np.uint64(0x02000000)) & np.uint64(0xFFFFFFFF)
    return (near > np.uint64(16 << 3)) & (rng <= np.uint64(0xFCFFFFFE))


def neighbours(q, k):
    """The double k ulps away from q (same sign, finite, away from zero crossings for the magnitudes used here)."""
    b = q.view(np.int64)
    return np.where(q >= 0, b + k, b - k).astype(np.int64).view(np.float64)


def check(q):
    ok = guard_accepts(q)
    with np.errstate(over="ignore"):
        f0 = q.astype(np.float32)
        for k in range(-3, 4):
            fk = neighbours(q, k).astype(np.float32)
            bad = ok & (fk.view(np.uint32) != f0.view(np.uint32))
            assert not bad.any(), (q[bad][:4], k)
    return ok


def test_accepted_quotients_round_like_their_neighbours():
    rng = np.random.default_rng(7)
    # pixel-like magnitudes, random mantissas over many binades (incl. the edges of the float32 range), both signs
    q = np.concatenate([
        rng.uniform(-2000, 2000, 2_000_000),
        np.ldexp(rng.uniform(1, 2, 2_000_000), rng.integers(-160, 160, 2_000_000)) * rng.choice([-1.0, 1.0], 2_000_000),
    ])
    ok = check(q)
    assert ok[:2_000_000].mean() > 0.999999 - 1e-5          # the guard rejects ~6e-8 of ordinary quotients


def test_quotients_at_float32_rounding_boundaries_are_rejected_or_safe():
    rng = np.random.default_rng(8)
    base = rng.uniform(1, 2000, 1_000_000).view(np.uint64)
    for j in range(-20, 21):
        q = ((base & ~np.uint64(0x1FFFFFFF)) | np.uint64(0x10000000 + j)).view(np.float64)
        ok = check(q)
        if abs(j) <= 8:
            assert not ok.any()                            # inside the margin: always sent to the IEEE divisions
    # values whose float32 image is zero, subnormal, below 2^-125, infinite or NaN never pass
    special = np.array([0.0, -0.0, 1e-50, 1.0e-38, 2.0e-38, 2.3e-38, 3.5e38, 1e300, np.inf, -np.inf, np.nan])
    assert not guard_accepts(special).any()
    assert guard_accepts(np.array([2.4e-38, 3.0e38, -3.0e38, 1.0, -1234.56789])).all()
