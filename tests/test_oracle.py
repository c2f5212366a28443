"""CPU: the oracle against the reference's doctest known answers, and python oracle == C oracle."""
import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import hx_oracle as O


def test_known_answers_one_plus_sin():
    # hendrics/efsearch.py:387-392, hendrics/ffa.py:84-91,126-131,179-186,200-206
    phase = 2 * np.pi * np.arange(0, 1, 0.01)
    norm = np.sin(phase) + 1
    assert np.isclose(O.z_n_fast(phase, norm, n=4), 50)
    assert np.isclose(O.z_n_fast(phase, norm, n=2), 50)
    assert np.isclose(O.z_n_fast_cached(norm, n=4), 50)
    assert np.isclose(O.z_n_fast_cached(norm, n=2), 50)
    ks, zs = O.z_n_fast_cached_all(norm, nmax=4)
    assert ks[1] == 2 and np.isclose(zs[1], 50) and np.isclose(zs[3], 50)
    h, m = O.h_test(norm, nmax=4)
    assert np.isclose(h, 50) and m == 1


def test_histogram_equals_numpy_on_in_range_data():
    # hendrics/base.py:682-690
    rng = np.random.default_rng(0)
    x = rng.uniform(0.0, 1.0, 100)
    y = rng.uniform(2.0, 3.0, 100)
    H, _, _ = np.histogram2d(x, y, bins=(5, 5), range=[(0.0, 1.0), (2.0, 3.0)])
    Hn = O.histogram2d(x, y, bins=(5, 5), range=[[0.0, 1.0], [2.0, 3.0]])
    assert np.all(H == Hn)


def test_find_peaks_in_image_toy():
    # hendrics/base.py:1129-1138
    rng = np.random.default_rng(1)
    image = rng.normal(0, 0.01, (10, 10))
    image[5, 5], image[7, 2], image[8, 1] = 2, 1, 3
    assert np.allclose(O.find_peaks_in_image(image, n=2), [(8, 1), (5, 5)])


def test_profile_fast_equals_upstream_formulation():
    rng = np.random.default_rng(2)
    for nbin in (16, 128, 10):
        ph = rng.uniform(0, 1, 5000)
        a = O.profile_fast(ph, nbin)
        b = O.profile_fast_upstream(ph, nbin)
        assert np.array_equal(a, b)
        assert np.array_equal(a, np.histogram(ph, bins=nbin, range=(0, 1))[0])


def test_segment_selection_rules():
    ts = np.array([10.0, 11.0, 12.5, 13.0, 20.0])
    t, masks = O.select_segments(ts)
    assert t[0] == 0 and len(masks) == 1 and masks[0].sum() == 4  # A1 re-zero, A2 last event dropped
    t, masks = O.select_segments(np.arange(0, 100.0, 0.5), segment_size=30.0)
    assert len(masks) == 4 or len(masks) == 3  # trailing short segment is skipped (< 0.2 * segment)
    assert all(m.sum() == 60 for m in masks[:3])


@pytest.mark.parametrize("nbin,nharm,fdots", [(128, 2, [0.0]), (16, 0, [0.0]), (32, 3, [-1e-6, 2e-6]), (10, 1, [1e-6])])
def test_c_oracle_equals_python_oracle(nbin, nharm, fdots):
    ts = O.simulate_events(6000, 500.0, seed=3)
    freqs = 9.85 + 0.1 * np.arange(40) / 40
    out = O.folding_search_core(ts, freqs, fdots=fdots, nbin=nbin, nharm=nharm, return_profiles=True)
    stats_py, prof_py = out[-2], out[-1]
    tsel, masks = O.select_segments(ts)
    s_c, p_c = C.fold_search(tsel[masks[0]], freqs, fdots, nbin, nharm, want_profiles=True)
    assert np.array_equal(p_c.reshape(prof_py.shape), prof_py)
    np.testing.assert_allclose(s_c.reshape(np.shape(stats_py)), stats_py, rtol=1e-12)
    # the parallel python variant used by the CPU baseline is the same computation
    out2 = O.folding_search_core(ts, freqs, fdots=fdots, nbin=nbin, nharm=nharm, parallel=True)
    np.testing.assert_allclose(out2[-1], stats_py, rtol=1e-12)


def test_pulse_is_recovered():
    # behavioural fixture of hendrics/tests/test_efsearch.py:218-242, 287-316
    ts = O.simulate_events(10000, 25.25, seed=5) + 8e7
    for func, kw in ((O.epoch_folding_search, dict(nbin=64)), (O.z_n_search, dict(nbin=64, nharm=2))):
        f, s, step, length = O.folding_search(ts, 8e7, 9.85, 9.95, func=func, oversample=2, **kw)
        assert np.isclose(f[np.argmax(s)], 1 / 0.101, atol=1 / 25.25)


def test_exact_bins_flag_fp64_rounding():
    """Long baseline: fp64 phases move a few photons across bin edges w.r.t. exact arithmetic
    (SURVEY finding 0.4) -- the oracle must follow fp64, not exact, arithmetic."""
    rng = np.random.default_rng(7)
    ts = np.sort(rng.uniform(0, 1e8, 20000))
    f, fd, nbin = 700.123456789, 1.2345e-9, 128
    ph = np.empty_like(ts)
    O.pulse_phase_fast(ts, f, fd, ph)
    b64 = np.floor(ph * nbin).astype(np.int64)
    bex = O.exact_bins(ts, f, fd, nbin)
    flips = int(np.sum(b64 != bex))
    assert 0 < flips < 400


def test_topk_flat_is_deterministic():
    x = np.array([1.0, 5.0, 5.0, 2.0, 5.0])
    idx, val = O.topk_flat(x, 3)
    assert list(idx) == [1, 2, 4] and list(val) == [5.0, 5.0, 5.0]


# ---- slow (`fold_events`) branch: weights / exposure correction --------------------------------------
def test_phase_exposure_known_answers():
    # an integer number of periods exposes every bin equally
    np.testing.assert_allclose(O.phase_exposure(0.0, 7.0, 1, 8), np.ones(8), rtol=1e-15)
    # 2.5 periods from phase 0: the first half of the profile is seen 3 times, the second half twice
    e = O.phase_exposure(0.0, 2.5, 1, 4)
    np.testing.assert_allclose(e, [1, 1, 2 / 3, 2 / 3], rtol=1e-15)
    # a gap: [0, 1.25] + [1.75, 3]: bins 0 and 3 seen 3 times, bins 1 and 2 twice
    e = O.phase_exposure(0.0, 3.0, 1, 4, gti=np.array([[0.0, 1.25], [1.75, 3.0]]))
    np.testing.assert_allclose(e, [1, 2 / 3, 2 / 3, 1], rtol=1e-15)


def test_fold_events_unit_weights_equals_fast_profile():
    ts = O.simulate_events(5000, 100.0, seed=8)
    prof = np.zeros(32, dtype=np.int64)
    for f, fd in ((9.9, 0.0), (9.91, 1e-5), (3.3, -2e-5)):
        O._fold_one(ts, f, fd, 32, prof)
        bins, p, perr = O.fold_events(ts, f, fd, nbin=32)
        assert np.array_equal(p, prof)
        np.testing.assert_allclose(bins, (np.arange(32) + 0.5) / 32)
        np.testing.assert_allclose(perr, np.sqrt(prof))
        w = np.full(ts.size, 2.0)
        assert np.array_equal(O.fold_events(ts, f, fd, nbin=32, weights=w)[1], 2 * prof)


def test_slow_branch_reduces_to_fast_branch():
    ts = O.simulate_events(4000, 80.0, seed=4)
    f = np.arange(9.85, 9.95, 0.01)
    a = O.z_n_search(ts, f, nharm=2, nbin=64, fdots=[0, 1e-5])
    b = O.z_n_search(ts, f, nharm=2, nbin=64, fdots=[0, 1e-5], weights=np.ones(ts.size))
    np.testing.assert_allclose(a[2], b[2], rtol=1e-13)
    a = O.epoch_folding_search(ts, f, nbin=16, segment_size=30.0)
    b = O.epoch_folding_search(ts, f, nbin=16, segment_size=30.0, weights=np.ones(ts.size))
    np.testing.assert_allclose(a[1], b[1], rtol=1e-13)
    with pytest.raises(ValueError, match="specify the GTIs"):
        O.epoch_folding_search(ts, f, expocorr=True)


def test_orbital_profiles_do_not_depend_on_the_sin_flavour():
    """Row f4: folding_orbital_search goes through two ``np.sin`` per event (hendrics/efsearch.py:295-296) and numpy's
    fp64 sin is a SIMD kernel whose last bit depends on the CPU (it differs from the correctly rounded value for ~0.1 %
    of arguments on AVX-512).  That does not reach the profiles: X*sin is far smaller than t, the rounded difference
    rarely changes, and a last-bit change of a time almost never crosses a bin edge.  So bit-exact profile parity is a
    meaningful bar for the orbital kernel too (tests/test_orbital_gpu.py), whatever sin the host library uses."""
    f0, X, Porb, T0 = 1.2345, 3.0, 5000.0, 1234.0
    te = O.simulate_events(40001, 20000.0, f0=f0, amp=0.8, seed=16)
    ts = np.sort(te + X * np.sin(2 * np.pi * (te - T0) / Porb))
    T0s = np.concatenate([[T0], np.random.default_rng(1).uniform(0, Porb, 30)])
    s_np, p_np = O.orbital_stats(ts, f0, 0.0, X, Porb, T0s, fun=O.epoch_folding_search, return_profiles=True, nbin=16)
    s_cr, p_cr = O.orbital_stats(ts, f0, 0.0, X, Porb, T0s, fun=O.epoch_folding_search, return_profiles=True, nbin=16,
                                 sin=O.sin_correctly_rounded)
    assert np.array_equal(p_np, p_cr) and np.array_equal(s_np, s_cr)
    a = np.random.default_rng(0).uniform(-50, 50, 200000)
    assert np.mean(np.sin(a) != O.sin_correctly_rounded(a)) < 0.01
    # the demodulated times themselves: a last-bit difference for a small fraction of the events at most
    d_np, d_cr = O.orbital_demodulate(ts, T0, X, Porb), O.orbital_demodulate(ts, T0, X, Porb, sin=O.sin_correctly_rounded)
    assert np.mean(d_np != d_cr) < 0.01 and np.max(np.abs(d_np - d_cr)) <= np.spacing(ts.max())
