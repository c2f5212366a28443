"""CPU checks of the exactness arguments the CUDA kernels rely on, independent of any GPU run.

np.longdouble (x87, 64-bit mantissa on this platform) holds x * nbin exactly for a 53-bit x and nbin < 2^10, so it serves
as exact arithmetic here; the reference's own operation sequence is evaluated in float64 as numpy does.
"""
import numpy as np
import pytest

LD_OK = np.finfo(np.longdouble).nmant >= 63


def _reference_bins(x, nbin):
    """_pulse_phase_fast / _profile_fast on an already computed phase x: ph - floor(ph), floor(ph * nbin), A4 clamp"""
    fr = x - np.floor(x)
    return np.minimum(np.floor(fr * nbin).astype(np.int64), nbin - 1)


def _adversarial_phases(rng, nbin, lo, hi, n):
    """phases in [lo, hi): random ones plus values a few ulps around every kind of bin edge"""
    x = rng.uniform(lo, hi, n)
    k = np.floor(rng.uniform(lo, hi - 1, n))
    j = rng.integers(0, nbin + 1, n)
    edge = k + j / nbin
    for d in (-3, -2, -1, 0, 1, 2, 3):
        y = edge.copy()
        for _ in range(abs(d)):
            y = np.nextafter(y, np.inf if d > 0 else -np.inf)
        x = np.concatenate([x, y])
    return x[(x >= lo) & (x < hi)]


@pytest.mark.skipif(not LD_OK, reason="needs an 80-bit long double as exact arithmetic")
@pytest.mark.parametrize("nbin", [3, 7, 10, 20, 24, 48, 50, 100, 777, 1000, 1023])
def test_generic_short_path_identity(nbin):
    """x >= 512: the reference's bin equals floor(x * nbin) - floor(x) * nbin in exact arithmetic, and the product rounded
    DOWN to fp64 has the same floor (what hx_fold.cu's GF path computes with DMUL.RD + two magic adds + IMAD)."""
    rng = np.random.default_rng(nbin)
    for lo, hi in ((512.0, 1024.0), (1e4, 1e5), (3e6, 4e6), (1e9, 2e9), (2.0**40, 2.0**40 + 1e6)):
        x = _adversarial_phases(rng, nbin, lo, hi, 20000)
        ref = _reference_bins(x, nbin)
        p = x.astype(np.longdouble) * np.longdouble(nbin)  # exact
        assert np.all(p.astype(np.float64).astype(np.longdouble) - p <= np.spacing(p.astype(np.float64)))  # sanity
        exact = (np.floor(p) - np.floor(x).astype(np.longdouble) * nbin).astype(np.int64)
        assert np.array_equal(exact, ref), (nbin, lo)
        # fp64 product rounded down: the largest double <= p
        pr = p.astype(np.float64)
        pr = np.where(pr.astype(np.longdouble) > p, np.nextafter(pr, -np.inf), pr)
        assert np.all(pr.astype(np.longdouble) <= p)
        assert np.array_equal(np.floor(pr), np.floor(p).astype(np.float64)), (nbin, lo)


def test_generic_short_path_needs_the_512_cycle_threshold():
    """below 512 cycles fr * nbin is not always exact: there the rounded product can reach the next integer (the reference
    then bins one higher than exact arithmetic), which is why those events replay the reference's operation sequence"""
    if not LD_OK:
        pytest.skip("needs an 80-bit long double")
    found = False
    rng = np.random.default_rng(5)
    for nbin in (3, 7, 10, 20, 100, 777, 1000):
        x = _adversarial_phases(rng, nbin, 0.0, 8.0, 200000)
        ref = _reference_bins(x, nbin)
        p = x.astype(np.longdouble) * np.longdouble(nbin)
        exact = (np.floor(p) - np.floor(x).astype(np.longdouble) * nbin).astype(np.int64)
        found |= bool(np.any(exact != ref))
    assert found


@pytest.mark.parametrize("nbin", [2, 16, 128, 256])
def test_power_of_two_fast_path_identity(nbin):
    """power-of-two nbin, x >= 0: scaling by nbin * 128 is exact, so the low bits of floor(x * nbin * 128) hold the
    reference's bin (bits [7, 7 + log2 nbin)) -- the magic-add / AND path of K1 and K2."""
    rng = np.random.default_rng(nbin)
    for lo, hi in ((0.0, 4.0), (1.0, 1e3), (1e6, 2e6), (1e9, 1e9 + 1e5)):
        x = _adversarial_phases(rng, nbin, lo, hi, 20000)
        ref = _reference_bins(x, nbin)
        scaled = x * (nbin * 128.0)  # exact (power of two)
        fast = (np.floor(scaled).astype(np.int64) >> 7) & (nbin - 1)
        assert np.array_equal(fast, ref), (nbin, lo)


def test_packed_counter_increment_encoding():
    """nbin >= 128 packed counters: word = bin mod nbin/2, increment (lo & topbit) | 1 puts upper-half bins in the high
    field; unpacking low - high / high recovers both bins for any event sequence shorter than the flush interval"""
    nbin, log2 = 128, 7
    shift = 6 + log2
    topbit = 1 << shift
    rng = np.random.default_rng(0)
    bins = rng.integers(0, nbin, 8000)  # < 2^shift - 1 events
    words = np.zeros(nbin // 2, dtype=np.int64)
    for b in bins:
        lo = int(b) << 7  # the scaled phase's low word: bin at bits [7, 14)
        words[(lo >> 7) & (nbin // 2 - 1)] += (lo & topbit) | 1
    up = words >> shift
    low = (words & (topbit - 1)) - up
    ref = np.bincount(bins, minlength=nbin)
    assert np.array_equal(np.concatenate([low, up]), ref)


@pytest.mark.skipif(not LD_OK, reason="needs an 80-bit long double as exact arithmetic")
def test_phase_fraction_is_exact_outside_minus_half_to_zero():
    """`ph - floor(ph)` is computed without rounding for ph >= 0 and for ph <= -0.5 (K1's non-negative fast path, K2's
    |ph| >= 0.75 condition); inside (-0.5, 0) it can round, which is why those photons replay the reference's sequence
    (K1 generic mode, K2's fix-up kernel)."""
    rng = np.random.default_rng(1)
    mags = np.concatenate([rng.uniform(0.5, 4, 100000), 10 ** rng.uniform(0, 9, 100000)])
    for sign in (1.0, -1.0):
        ph = sign * mags
        fr = ph - np.floor(ph)
        exact = ph.astype(np.longdouble) - np.floor(ph).astype(np.longdouble)
        assert np.array_equal(fr.astype(np.longdouble), exact)
    ph = -rng.uniform(1e-9, 0.5, 200000)
    fr = ph - np.floor(ph)
    exact = ph.astype(np.longdouble) - np.floor(ph).astype(np.longdouble)
    assert np.any(fr.astype(np.longdouble) != exact)


def test_fdot_term_operand_orders_round_identically():
    """A3 / K1: the reference's ((0.5 * t) * t) * fd, the kernel's fl(t * t) * (0.5 * fd) and the pre-scaled
    fl(t * t) * (0.5 * fd * nbin * 128) / (nbin * 128) are the same double: scaling by powers of two commutes with rounding
    (no underflow in the ranges the host admits)."""
    rng = np.random.default_rng(3)
    t = np.concatenate([rng.uniform(0, 1e5, 200000), 10 ** rng.uniform(-3, 8.5, 200000)])
    fd = np.concatenate([rng.uniform(-1, 1, 200000) * 10 ** rng.uniform(-15, -4, 200000)] * 2)
    ref = ((0.5 * t) * t) * fd
    assert np.array_equal((t * t) * (0.5 * fd), ref)
    for nbin in (16, 128, 1024):
        s = nbin * 128.0
        assert np.array_equal(((t * t) * ((0.5 * fd) * s)) / s, ref)
        # and the scaled sum: fl(t*f*s) + fl(m2*s) rounds like (fl(t*f) + fl(m2)) * s
        f = rng.uniform(0.01, 1000, t.size)
        assert np.array_equal((t * (f * s)) + ((t * t) * ((0.5 * fd) * s)), ((t * f) + ref) * s)


def test_harmonic_domain_shift_and_sum_identity():
    """K3: a circular shift of a profile rotates its k-th harmonic by k * theta_shift, so the Z^2_n of the shifted sum of
    the time slices is obtained from per-slice harmonic coefficients (what qffa_step2_kernel accumulates) -- equal to the
    direct roll-and-add of hendrics/efsearch.py:350-363 to rounding."""
    rng = np.random.default_rng(4)
    nprof, nbin, n = 32, 64, 3
    prof = rng.poisson(5.0, (nprof, nbin)).astype(np.float64)
    shifts = rng.integers(0, nbin, nprof)
    splat = np.zeros(nbin)
    for k in range(nprof):
        splat += np.roll(prof[k], shifts[k])
    j = np.arange(nbin)
    direct = 0.0
    viaco = 0.0
    for h in range(1, n + 1):
        c, s = np.cos(2 * np.pi * h * j / nbin), np.sin(2 * np.pi * h * j / nbin)
        direct += (splat @ c) ** 2 + (splat @ s) ** 2
        Cr, Sr = prof @ c, prof @ s                      # per-slice coefficients
        th = 2 * np.pi * h * shifts / nbin
        C = np.sum(Cr * np.cos(th) - Sr * np.sin(th))    # roll by +s: bin b <- b - s, phase advances by theta_s
        S = np.sum(Sr * np.cos(th) + Cr * np.sin(th))
        viaco += C ** 2 + S ** 2
    np.testing.assert_allclose(viaco, direct, rtol=1e-12)
