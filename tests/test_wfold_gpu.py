"""K1w (weighted / exposure-corrected fold) against the oracle's restatement of upstream's
`fold_events` branch.  Tolerances: profiles bit-exact when the weights are integers (light-curve counts; every
partial sum is exact), else rtol 1e-12 (different summation order); statistics rtol 1e-9 (north_star)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import hx_oracle as O

pytestmark = pytest.mark.gpu

RTOL_STAT = 1e-9  # BASELINE.json north_star tolerance for chi^2 / Z^2_n
PULSE_F = 1 / 0.101


def _oracle_profiles(ts, freqs, fdots, nbin, weights, gti=None):
    out = np.zeros((len(fdots), len(freqs), nbin))
    for i, fd in enumerate(fdots):
        for j, f in enumerate(freqs):
            out[i, j] = O.fold_events(ts, f, fd, nbin=nbin, weights=weights, gti=gti, expocorr=gti is not None)[1]
    return out


def _oracle_stats(prof, nharm):
    fn = (lambda p: O.ef_profile_stat(p)) if nharm == 0 else (lambda p: O.z_n_binned_events(p, nharm))
    return np.array([[fn(p) for p in row] for row in prof])


@pytest.mark.parametrize("nbin,nharm", [(16, 0), (64, 2), (128, 4), (10, 0), (12, 1), (100, 2)])
def test_integer_weights_bit_exact(nbin, nharm):
    from hendrics_b200.search import fold_profiles_weighted

    rng = np.random.default_rng(nbin)
    ts = O.simulate_events(30001, 300.0, seed=nbin)
    w = rng.poisson(3.0, ts.size).astype(np.float64)
    freqs = 9.85 + 0.002 * np.arange(37)
    fdots = np.array([0.0, 2e-6, -1e-6])
    stats, prof = fold_profiles_weighted(ts, freqs, fdots, nbin=nbin, nharm=nharm, weights=w)
    ref_p = _oracle_profiles(ts, freqs, fdots, nbin, w)
    assert np.array_equal(prof, ref_p)
    np.testing.assert_allclose(stats, _oracle_stats(ref_p, nharm), rtol=RTOL_STAT)
    assert prof.sum(axis=-1).min() == w.sum() == prof.sum(axis=-1).max()


def test_unit_weight_array_equals_unweighted_kernel():
    from hendrics_b200.search import fold_profiles, fold_profiles_weighted

    ts = O.simulate_events(20000, 100.0, seed=5)
    freqs = 9.8 + 0.001 * np.arange(200)
    fdots = np.array([0.0, 1e-5])
    s0, p0 = fold_profiles(ts, freqs, fdots, nbin=32, nharm=2)
    s1, p1 = fold_profiles_weighted(ts, freqs, fdots, nbin=32, nharm=2, weights=np.ones(ts.size))
    s2, p2 = fold_profiles_weighted(ts, freqs, fdots, nbin=32, nharm=2, weights=1)
    assert np.array_equal(p0, p1) and np.array_equal(p1, p2)
    np.testing.assert_allclose(s1, s0, rtol=1e-12)
    np.testing.assert_allclose(s2, s0, rtol=1e-12)


@pytest.mark.parametrize("nbin", [32, 20])
def test_float_weights(nbin):
    from hendrics_b200.search import fold_profiles_weighted

    rng = np.random.default_rng(7)
    ts = O.simulate_events(12345, 50.0, seed=7)
    w = rng.uniform(0.1, 2.0, ts.size)
    freqs = 9.85 + 0.004 * np.arange(20)
    for nharm in (0, 3):
        stats, prof = fold_profiles_weighted(ts, freqs, [0.0, 1e-5], nbin=nbin, nharm=nharm, weights=w)
        ref_p = _oracle_profiles(ts, freqs, [0.0, 1e-5], nbin, w)
        np.testing.assert_allclose(prof, ref_p, rtol=1e-12)
        np.testing.assert_allclose(stats, _oracle_stats(ref_p, nharm), rtol=RTOL_STAT)


@pytest.mark.parametrize("nbin,fdots", [(16, [0.0]), (64, [0.0, 3e-6]), (24, [1e-6])])
def test_exposure_correction(nbin, fdots):
    from hendrics_b200.search import fold_profiles_weighted

    ts = O.simulate_events(40000, 400.0, seed=11)
    gti = np.array([[0.0, 57.3], [80.2, 199.9], [200.0, 233.0], [301.5, 420.0], [500.0, 600.0]])
    keep = np.zeros(ts.size, bool)
    for g0, g1 in gti:
        keep |= (ts >= g0) & (ts < g1)
    ts = ts[keep]
    freqs = 9.85 + 0.003 * np.arange(33)
    for nharm in (0, 2):
        stats, prof = fold_profiles_weighted(ts, freqs, fdots, nbin=nbin, nharm=nharm, gti=gti)
        ref_p = _oracle_profiles(ts, freqs, fdots, nbin, 1, gti=gti)
        np.testing.assert_allclose(prof, ref_p, rtol=1e-12)
        np.testing.assert_allclose(stats, _oracle_stats(ref_p, nharm), rtol=RTOL_STAT)


def test_exposure_normalisation_matches_long_double_oracle():
    """The per-bin exposure itself (profile of unit weights / corrected profile) vs np.longdouble upstream code."""
    from hendrics_b200.search import fold_profiles_weighted

    ts = np.sort(np.random.default_rng(3).uniform(0.0, 97.0, 5000))
    gti = np.array([[0.0, 10.13], [12.5, 50.0], [50.0, 51.0], [70.7, 97.0]])
    f, fd, nbin = 3.21987, 1.5e-4, 16
    _, corrected = fold_profiles_weighted(ts, [f], [fd], nbin=nbin, gti=gti)
    _, raw = fold_profiles_weighted(ts, [f], [fd], nbin=nbin)
    dt = abs(ts[1] - ts[0])
    sp, ep = O.pulse_phase(np.array([ts[0], ts[-1] + dt]), f, fd, to_1=False)
    expo = O.phase_exposure(sp, ep, 1, nbin, gti=O.pulse_phase(gti, f, fd, to_1=False))
    np.testing.assert_allclose(raw[0, 0] / corrected[0, 0], expo, rtol=1e-14)


def test_public_api_weights_expocorr_segments():
    from hendrics_b200 import epoch_folding_search, z_n_search

    rng = np.random.default_rng(21)
    ts = O.simulate_events(30000, 300.0, seed=21) + 1000.0
    w = rng.poisson(2.0, ts.size)
    freqs = np.arange(9.85, 9.95, 0.002)
    gti = np.array([[1000.0, 1120.0], [1130.0, 1300.0]]) - 1000.0
    cases = [
        dict(weights=w),
        dict(weights=w, fdots=[0.0, 1e-6]),
        dict(expocorr=True, gti=gti),
        dict(expocorr=True, gti=gti, weights=w.astype(float), segment_size=100.0),
        dict(weights=w, segment_size=77.0),
    ]
    for kw in cases:
        for mine, ref, extra in ((epoch_folding_search, O.epoch_folding_search, dict(nbin=32)),
                                 (z_n_search, O.z_n_search, dict(nbin=64, nharm=2))):
            a = mine(ts, freqs, **extra, **kw)
            b = ref(ts, freqs, **extra, **kw)
            assert len(a) == len(b)
            for x, y in zip(a[:-1], b[:-1]):
                assert np.array_equal(x, y)
            np.testing.assert_allclose(a[-1], b[-1], rtol=RTOL_STAT)
    with pytest.raises(ValueError, match="specify the GTIs"):
        z_n_search(ts, freqs, expocorr=True)


def test_event_splits_and_trial_batches(monkeypatch):
    """few trials x many events (events split over CTAs) and a tiny scratch cap (several trial batches)."""
    from hendrics_b200.search import fold_profiles_weighted

    rng = np.random.default_rng(9)
    ts = O.simulate_events(300001, 3000.0, seed=9)
    w = rng.integers(0, 5, ts.size).astype(np.float64)
    freqs = 9.9 + 1e-4 * np.arange(5)
    stats, prof = fold_profiles_weighted(ts, freqs, [0.0], nbin=16, nharm=0, weights=w)
    ref_p = _oracle_profiles(ts, freqs, [0.0], 16, w)
    assert np.array_equal(prof, ref_p)
    monkeypatch.setenv("HX_FOLD_SCRATCH_CAP_KB", "64")
    ts2 = ts[:5001]
    freqs2 = 9.8 + 1e-3 * np.arange(700)
    stats2, prof2 = fold_profiles_weighted(ts2, freqs2, [0.0, 1e-6], nbin=16, nharm=1, weights=w[:5001])
    ref2 = _oracle_profiles(ts2, freqs2, [0.0, 1e-6], 16, w[:5001])
    assert np.array_equal(prof2, ref2)
    np.testing.assert_allclose(stats2, _oracle_stats(ref2, 1), rtol=RTOL_STAT)


def test_c_abi_host_entry():
    from hendrics_b200 import _cabi

    rng = np.random.default_rng(2)
    ts = O.simulate_events(7777, 60.0, seed=2)
    w = rng.poisson(4.0, ts.size).astype(np.float64)
    freqs = _cabi.as_f64(9.8 + 0.01 * np.arange(21))
    fdots = _cabi.as_f64([0.0, 1e-5])
    gti = _cabi.as_f64([[0.0, 30.0], [31.0, 61.0]])
    nbin = 16
    stats = np.empty((2, 21))
    prof = np.empty((2, 21, nbin))
    dt = abs(ts[1] - ts[0])
    rc = _cabi.lib().hx_wfold_search(_cabi.dptr(ts), _cabi.dptr(w), 1.0, ts.size, _cabi.dptr(freqs), 21,
                                     _cabi.dptr(fdots), 2, nbin, 2, _cabi.dptr(gti), 2, ts[0], ts[-1] + dt,
                                     _cabi.dptr(stats), _cabi.dptr(prof), 0)
    _cabi.check(rc)
    ref_p = _oracle_profiles(ts, freqs, fdots, nbin, w, gti=gti)
    np.testing.assert_allclose(prof, ref_p, rtol=1e-12)
    np.testing.assert_allclose(stats, _oracle_stats(ref_p, 2), rtol=RTOL_STAT)
    # argument errors are reported, not computed around
    rc = _cabi.lib().hx_wfold_search(_cabi.dptr(ts), None, 1.0, ts.size, _cabi.dptr(freqs), 21, _cabi.dptr(fdots), 2,
                                     1, 2, None, 0, 0.0, 1.0, _cabi.dptr(stats), None, 0)
    assert rc == _cabi.HX_ERR_ARG


def test_cli_from_light_curve_and_expocorr(tmp_path):
    """hendrics/tests/test_efsearch.py:266-285, 318-344 (search from a light curve) and `--expocorr`."""
    from hendrics_b200.efsearch import main_efsearch, main_zsearch
    from hendrics_b200.io import EventList, Lightcurve, load_folding, save_events, save_lcurve

    ts = O.simulate_events(10000, 25.25, seed=42) + 8e7
    dt = 1 / 64 / PULSE_F * 3.0001  # bins much shorter than the pulse period
    edges = np.arange(8e7, 8e7 + 25.25 + dt, dt)
    counts, _ = np.histogram(ts, bins=edges)
    lc = Lightcurve(time=(edges[:-1] + edges[1:]) / 2, counts=counts, gti=np.array([[8e7, 8e7 + 25.25]]), dt=dt,
                    mjdref=56000)
    lcfile = str(tmp_path / "lcurve_lc.p")
    save_lcurve(lc, lcfile)
    out = main_efsearch([lcfile, "-f", "9.85", "-F", "9.95", "-n", "64", "--fit-candidates"])
    assert out[0].endswith("lcurve_EF_9.85-9.95Hz.p")
    ef = load_folding(out[0])
    assert np.isclose(ef.peaks[0], PULSE_F, atol=1 / 25.25)
    rf, rs, _, _ = O.folding_search(lc.time, 8e7, 9.85, 9.95, func=O.epoch_folding_search, oversample=2, nbin=64,
                                    weights=lc.counts)
    assert np.array_equal(ef.freq, rf)
    np.testing.assert_allclose(ef.stat, rs, rtol=RTOL_STAT)
    out = main_zsearch([lcfile, "-f", "9.85", "-F", "9.95", "-n", "64", "--fit-candidates", "--fit-frequency",
                        str(PULSE_F)])
    ef = load_folding(out[0])
    assert out[0].endswith("lcurve_Z22_9.85-9.95Hz.p") and ef.N == 2
    assert abs(ef.freq[np.argmax(ef.stat)] - PULSE_F) < 1 / 25.25

    gti = np.array([[8e7, 8e7 + 10.0], [8e7 + 12.0, 8e7 + 25.25]])
    keep = ((ts >= gti[0, 0]) & (ts < gti[0, 1])) | ((ts >= gti[1, 0]) & (ts < gti[1, 1]))
    evfile = str(tmp_path / "events_ev.p")
    save_events(EventList(time=ts[keep], gti=gti, mjdref=56000), evfile)
    out = main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "-n", "64", "--expocorr"])
    ef = load_folding(out[0])
    rf, rs, _, _ = O.folding_search(ts[keep], 8e7, 9.85, 9.95, func=O.z_n_search, oversample=2, nbin=64, nharm=2,
                                    expocorr=True, gti=gti - 8e7)
    np.testing.assert_allclose(ef.stat, rs, rtol=RTOL_STAT)
    assert abs(ef.freq[np.argmax(ef.stat)] - PULSE_F) < 1 / 25.25


@pytest.mark.parametrize("nbin", [16, 50, 128])
def test_cell_types_agree(nbin):
    """integer weights through the u32 cells == the same through the fp64 cells == the oracle; weights whose sum
    does not fit in 32 bits fall back to the fp64 cells; the pre-scaled bin path == the generic one."""
    from hendrics_b200 import _cabi
    from hendrics_b200.search import fold_profiles_weighted

    rng = np.random.default_rng(nbin)
    ts = O.simulate_events(20011, 150.0, seed=nbin + 1)
    w = rng.poisson(2.5, ts.size).astype(np.float64)
    freqs = 9.85 + 0.002 * np.arange(70)
    fdots = [0.0, 1e-6]
    ref_p = _oracle_profiles(ts, freqs, fdots, nbin, w)
    lib = _cabi.lib()
    try:
        for force, want_cell in ((0, 1), (1, 1), (2, 0), (3, 0)):
            lib.hx_wfold_set_tuning(0, 0, 0, force)
            _, prof = fold_profiles_weighted(ts, freqs, fdots, nbin=nbin, nharm=1, weights=w)
            assert _cabi.wfold_last_kernel()[1] == want_cell
            assert np.array_equal(prof, ref_p)
        lib.hx_wfold_set_tuning(0, 0, 0, 0)
        big = w.copy()
        big[:3] = 2.0**31
        _, prof = fold_profiles_weighted(ts, freqs[:5], [0.0], nbin=nbin, nharm=1, weights=big)
        assert _cabi.wfold_last_kernel()[1] == 0
        assert np.array_equal(prof, _oracle_profiles(ts, freqs[:5], [0.0], nbin, big))
        _, prof = fold_profiles_weighted(ts, freqs[:5], [0.0], nbin=nbin, nharm=1, weights=1)
        assert _cabi.wfold_last_kernel()[1] == 2
        for R, U in ((1, 2), (1, 4), (2, 2), (2, 4), (4, 0)):
            lib.hx_wfold_set_tuning(R, U, 3, 0)
            _, p1 = fold_profiles_weighted(ts, freqs, fdots, nbin=nbin, nharm=1, weights=w)
            _, p2 = fold_profiles_weighted(ts, freqs, fdots, nbin=nbin, nharm=1, weights=w + 0.5)
            assert np.array_equal(p1, ref_p)
            np.testing.assert_allclose(p2, _oracle_profiles(ts, freqs, fdots, nbin, w + 0.5), rtol=1e-12)
    finally:
        lib.hx_wfold_set_tuning(0, 0, 0, 0)


def test_negative_phases_take_the_generic_bin_path():
    """times before the reference epoch (negative phases): ph - floor(ph) rounds, the fast path is not allowed."""
    from hendrics_b200.search import fold_profiles_weighted

    rng = np.random.default_rng(5)
    ts = O.simulate_events(9000, 90.0, seed=5) - 45.0
    w = rng.uniform(0.2, 3.0, ts.size)
    freqs = 9.85 + 0.003 * np.arange(40)
    for nbin in (32, 30):
        stats, prof = fold_profiles_weighted(ts, freqs, [0.0, -2e-5], nbin=nbin, nharm=2, weights=w)
        ref_p = _oracle_profiles(ts, freqs, [0.0, -2e-5], nbin, w)
        np.testing.assert_allclose(prof, ref_p, rtol=1e-12)
        np.testing.assert_allclose(stats, _oracle_stats(ref_p, 2), rtol=RTOL_STAT)
        _, pi = fold_profiles_weighted(ts, freqs, [0.0, -2e-5], nbin=nbin, nharm=2, weights=np.ones(ts.size))
        assert np.array_equal(pi, _oracle_profiles(ts, freqs, [0.0, -2e-5], nbin, 1))


@pytest.mark.parametrize("kind", ["counts", "float"])
def test_sparse_light_curve_drops_empty_bins(kind):
    """A light curve at fine time resolution is mostly zeros: they are dropped before the fold (adding a zero weight
    changes nothing) and the profiles stay identical to the oracle's, which folds every bin like upstream."""
    from hendrics_b200 import _cabi
    from hendrics_b200.search import fold_profiles_weighted

    ph = O.simulate_events(20000, 2000.0, seed=17)
    edges = np.arange(0.0, 2000.0 + 1e-9, 2e-3)  # 10^6 bins, 2 % occupied
    counts = np.histogram(ph, bins=edges)[0].astype(np.float64)
    t = (edges[:-1] + edges[1:]) / 2
    w = counts if kind == "counts" else counts * 0.37
    freqs = 9.85 + 0.002 * np.arange(40)
    lib = _cabi.lib()
    try:
        stats, prof = fold_profiles_weighted(t, freqs, [0.0, 1e-7], nbin=32, nharm=2, weights=w)
        ms, cell, n_used = _cabi.wfold_last_kernel()
        assert n_used == np.count_nonzero(w) < 0.03 * t.size and cell == (1 if kind == "counts" else 0)
        lib.hx_wfold_set_tuning(0, 0, 0, 4)  # keep the zeros
        stats_all, prof_all = fold_profiles_weighted(t, freqs, [0.0, 1e-7], nbin=32, nharm=2, weights=w)
        assert _cabi.wfold_last_kernel()[2] == t.size
    finally:
        lib.hx_wfold_set_tuning(0, 0, 0, 0)
    if kind == "counts":
        assert np.array_equal(prof, prof_all) and np.array_equal(stats, stats_all)
    else:  # the split of the event list over CTAs differs: fp64 partial sums combine in another order
        np.testing.assert_allclose(prof, prof_all, rtol=1e-13)
        np.testing.assert_allclose(stats, stats_all, rtol=1e-11)
    ref_p = _oracle_profiles(t, freqs[:6], [0.0, 1e-7], 32, w)
    if kind == "counts":
        assert np.array_equal(prof[:, :6], ref_p)
    else:
        np.testing.assert_allclose(prof[:, :6], ref_p, rtol=1e-12)
    np.testing.assert_allclose(stats[:, :6], _oracle_stats(ref_p, 2), rtol=RTOL_STAT)
    # all-zero weights: empty profiles, like upstream (0/0 statistics)
    _, p0 = fold_profiles_weighted(t[:1000], freqs[:3], [0.0], nbin=16, nharm=0, weights=np.zeros(1000))
    assert not p0.any()
