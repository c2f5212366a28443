"""Parity at BASELINE.json's full sizes.  The oracle cannot fold 10^12 event*trials in seconds, so the full-size runs
are checked (a) exactly, on a random subset of their trials that the C oracle does fold at the full event count, and
(b) through size-independent properties: conserved counts, additivity of profiles over a split of the event list,
independence of a trial's result from the grid it is launched in, and the chi^2 / Z^2 identities of a flat profile."""
import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import hx_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _selected(ts):
    """A1 / A2 of the upstream search: re-zero to the first event, drop t == t[-1]."""
    t = (ts - ts[0]).astype(np.float64)
    return t[t < t[-1]]


def test_c2_full_grid_subset_against_oracle():
    """C2: epoch folding nbin=16, 1e6 events x 1e6 trial frequencies, the whole grid in one call."""
    from hendrics_b200 import epoch_folding_search
    from hendrics_b200.search import fold_profiles

    ts = O.simulate_events(1_000_000, 1e5, seed=2)
    freqs = 7.4 + 5e-6 * np.arange(1_000_000)
    f, stats = epoch_folding_search(ts, freqs, nbin=16)
    assert stats.shape == (1_000_000,) and np.array_equal(f, freqs) and np.all(np.isfinite(stats))
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(freqs.size, 1500, replace=False))
    t = _selected(ts)
    ref = C.fold_search(t, freqs[pick], [0.0], 16, 0)
    np.testing.assert_allclose(stats[pick], ref[0], rtol=RTOL)
    # a trial does not depend on the grid it was launched in, nor on the kernel variant (fused vs scratch + profiles)
    s_sub, p_sub = fold_profiles(t, freqs[pick], [0.0], nbin=16, nharm=0)
    np.testing.assert_allclose(s_sub[0], stats[pick], rtol=1e-13)
    assert np.all(p_sub.sum(axis=-1) == t.size)  # every selected event is in exactly one bin of every trial
    # additivity over a split of the event list (bit-exact integer counts)
    k = 611_113
    _, p_a = fold_profiles(t[:k], freqs[pick[:200]], [0.0], nbin=16)
    _, p_b = fold_profiles(t[k:], freqs[pick[:200]], [0.0], nbin=16)
    assert np.array_equal(p_a + p_b, p_sub[:, :200])
    # the pulsation injected at 1/0.101 Hz is the global maximum of the 10^6-trial periodogram
    assert abs(f[np.argmax(stats)] - 1 / 0.101) < 2e-5
    # ALL 10^6 trials: the profiles of the whole grid (scratch variant), bit-exact against the oracle on the subset,
    # every one of them holding every selected event exactly once, and the chi^2 recomputed from them on the host equal
    # to the statistics of the fused-epilogue run above -- so every statistic of the grid is tied to integer counts
    s_all, p_all = fold_profiles(t, freqs, [0.0], nbin=16, nharm=0)
    p_all = p_all[0]
    _, ref_p = C.fold_search(t, freqs[pick], [0.0], 16, 0, want_profiles=True)
    assert np.array_equal(p_all[pick], ref_p[0])
    assert np.all(p_all.sum(axis=1) == t.size)
    mean = t.size / 16.0
    chi2 = ((p_all - mean) ** 2 / mean).sum(axis=1)
    np.testing.assert_allclose(stats, chi2, rtol=1e-12)
    np.testing.assert_allclose(s_all[0], stats, rtol=1e-13)


def test_c3_slice_full_event_count_against_oracle():
    """C3: Z^2_4 nbin=128, 1e7 events, a (64 fdot x 256 f) block of the grid; oracle on a subset of its trials."""
    from hendrics_b200 import z_n_search

    ts = O.simulate_events(10_000_000, 1e5, seed=3)
    freqs = 9.9 + 5e-6 * np.arange(256)
    fdots = -1.6e-9 + 5e-11 * np.arange(64)
    fg, fdg, stats = z_n_search(ts, freqs, nharm=4, nbin=128, fdots=fdots)
    assert stats.shape == (64, 256) and np.array_equal(fg[0], freqs) and np.array_equal(fdg[:, 0], fdots)
    t = _selected(ts)
    jf = np.array([0, 17, 101, 200, 255])
    for i in (0, 31, 32, 63):
        ref = C.fold_search(t, freqs[jf], fdots[i:i + 1], 128, 4)
        np.testing.assert_allclose(stats[i, jf], ref[0], rtol=RTOL)
    # fdot == 0 row (index 32) equals the f-only search: different kernel instantiation, same numbers
    assert fdots[32] == 0.0 or abs(fdots[32]) < 1e-20
    _, s1 = z_n_search(ts, freqs, nharm=4, nbin=128)
    if fdots[32] == 0.0:
        np.testing.assert_allclose(stats[32], s1, rtol=1e-13)


def test_c4_full_event_count_conserved_counts_and_oracle_rows():
    """C4: --fast on 1e8 events, nbin=128, nprof=2048: 2-D profiles of a few centre frequencies."""
    import torch

    from hendrics_b200.efsearch import _qffa_device_call, _shift_grids

    rng = np.random.default_rng(4)
    n_ev = 100_000_000
    tc = rng.uniform(-5e4, 5e4, n_ev)
    tc.sort()
    nbin, nprof, n = 128, 2048, 2
    L, Q = _shift_grids(nbin, 2, 8, True)
    cents = [(9.87, 0.0, 0.0), (9.9, 1e-10, 0.0), (9.93, -2e-10, 0.0)]
    td = torch.from_numpy(tc).cuda()
    stats, prof = _qffa_device_call(td, tc[0], tc[-1], cents, nbin, nprof, L, Q, n, want_profiles=True)
    prof = np.asarray(prof).astype(np.int64).reshape(len(cents), nprof, nbin)
    # hist2d excludes the right edge in time (base.py:674): the events at t == t[-1] are dropped, all others counted
    n_valid = int(np.count_nonzero(tc < tc[-1]))
    assert np.all(prof.sum(axis=(1, 2)) == n_valid)
    # time slices are equally long, the source is steady: the slice totals are the same for every centre frequency
    assert np.array_equal(prof[0].sum(axis=1), prof[1].sum(axis=1)) and np.array_equal(prof[0].sum(axis=1), prof[2].sum(axis=1))
    # bit-exact against the oracle on the first 5 million events' slices (the oracle's hist2d is serial)
    m = 5_000_000
    j_hi = int((tc[m - 1] - tc[0]) * (1.0 / ((tc[-1] - tc[0]) / nprof))) - 1  # slices completely inside tc[:m]
    sub = np.ascontiguousarray(tc[:m])
    for c, (f, fd, fdd) in enumerate(cents[:2]):
        ph = O.fast_phase_fdot(sub, f, fd) if fd else O.fast_phase(sub, f)
        i = (ph * nbin).astype(np.int64)
        j = ((sub - tc[0]) * (1.0 / ((tc[-1] - tc[0]) / nprof))).astype(np.int64)
        ref = np.zeros((nprof, nbin), dtype=np.int64)
        np.add.at(ref, (j, i), 1)
        assert np.array_equal(prof[c, :j_hi], ref[:j_hi])
    # statistics: finite, and equal to the oracle's shift-and-sum of the GPU's own (bit-exact) profiles
    ref_stats = C.fast_step(prof[0].astype(np.uint64), L, Q, n)
    np.testing.assert_allclose(np.asarray(stats).reshape(len(cents), -1)[0], ref_stats.ravel(), rtol=RTOL)


def test_flat_profile_identities_at_scale():
    """Events at exact multiples of 1/(f nbin) fill all bins equally: chi^2 == 0 and Z^2_n == 0 whatever the size."""
    from hendrics_b200.search import fold_profiles

    nbin, f = 64, 8.0
    n = 6_400_000
    t = (np.arange(n, dtype=np.float64) + 0.5) / (f * nbin)  # phase = (k + 0.5)/nbin mod 1, exact in fp64
    for nharm in (0, 3):
        s, p = fold_profiles(t, [f], [0.0], nbin=nbin, nharm=nharm)
        assert np.all(p == n // nbin)
        assert abs(s[0, 0]) < 1e-12 * n


def test_c1_exact_baseline_shape_everything_against_oracle():
    """C1 at its exact BASELINE shape (Z^2_2, 1e5 events x 1e4 trial frequencies, nbin = 128, no fdot): the oracle folds
    all 1e9 event*trials in about a second, so EVERYTHING is compared -- all 1e4 statistics within 1e-9, all 1e4
    profiles bit for bit, and the top-10 candidate list (value descending, ties by index) identical."""
    from hendrics_b200 import z_n_search
    from hendrics_b200.search import fold_profiles

    ts = O.simulate_events(100_000, 1e5, seed=1) + 8e7  # NuSTAR-like absolute times, removed by the search (A1)
    freqs = np.arange(9.876, 9.926 + 1e-8 * 5e-6, 5e-6)  # folding_search's grid for -f 9.876 -F 9.926, oversample 2
    assert abs(freqs.size - 10_000) <= 1
    f, stats = z_n_search(ts, freqs, nharm=2, nbin=128)
    t = _selected(ts)
    ref_s, ref_p = C.fold_search(t, freqs, [0.0], 128, 2, want_profiles=True)
    assert np.array_equal(f, freqs)
    np.testing.assert_allclose(stats, ref_s[0], rtol=RTOL)
    s2, prof = fold_profiles(t, freqs, [0.0], nbin=128, nharm=2)
    assert np.array_equal(prof, ref_p)  # 1e4 x 128 counts, bit-exact
    np.testing.assert_allclose(s2[0], stats, rtol=1e-13)  # fused epilogue == scratch + fold_stat_kernel
    assert np.array_equal(O.topk_flat(stats, 10)[0], O.topk_flat(ref_s[0], 10)[0])
    # the public candidate finder (EFPeriodogram.find_peaks -> search_best_peaks) on both statistic arrays
    from hendrics_b200.io import search_best_peaks

    srt = np.sort(ref_s[0])
    thr = 0.5 * (srt[-50] + srt[-51])  # between two values: a 1e-13 difference cannot move a run boundary
    bx, bs = search_best_peaks(freqs, stats, thr)
    rx, rs = search_best_peaks(freqs, ref_s[0], thr)
    assert np.array_equal(bx, rx)
    np.testing.assert_allclose(bs, rs, rtol=RTOL)
    assert abs(freqs[np.argmax(stats)] - 1 / 0.101) < 1e-5


def test_c3_bench_shape_profiles_bit_exact_through_packed_drains():
    """C3 at the exact shape bench.py times per GPU at N = 8 (Z^2_4, nbin = 128, 1e7 events x 2048 f x 64 fdots): the
    kernel variant is {R: 2, packed counters, shared fdot rows, fdot-term tile} and every CTA drains its packed
    counters ~94 times.  Profiles of an 8 x 8 sub-grid (64 trials spread over the grid) are compared bit for bit with
    the oracle at the full event count, their statistics within 1e-9, plus the conserved-count property on all
    131 072 profiles."""
    from hendrics_b200 import _cabi
    from hendrics_b200.search import fold_profiles

    ts = O.simulate_events(10_000_000, 1e5, seed=3)
    t = _selected(ts)
    freqs = 7.4 + 5e-6 * np.arange(2048)
    fdots = -1.6e-9 + 5e-11 * np.arange(64)
    stats, prof = fold_profiles(t, freqs, fdots, nbin=128, nharm=4)
    shape = _cabi.fold_last_kernel_shape()
    assert (shape["R"], shape["mode"], shape["rowfd"], shape["m2t"]) == (2, 2, 1, 1), shape
    assert prof.shape == (64, 2048, 128)
    assert np.all(prof.sum(axis=-1) == t.size)  # every event in exactly one bin of every one of the 131 072 trials
    jf = np.array([0, 31, 32, 63, 64, 1000, 2016, 2047])  # lane / warp / CTA boundaries of the trial layout
    ifd = np.array([0, 1, 17, 31, 32, 33, 62, 63])
    ref_s, ref_p = C.fold_search(t, freqs[jf], fdots[ifd], 128, 4, want_profiles=True)
    got_p = prof[np.ix_(ifd, jf)]
    assert np.array_equal(got_p, ref_p)
    np.testing.assert_allclose(stats[np.ix_(ifd, jf)], ref_s, rtol=RTOL)


def _poisson_times(n, T, seed, f0=1 / 0.101, amp=0.25):
    """n sorted event times on [0, T) from a rate ~ 1 + amp cos(2 pi f0 t), sorted BY CONSTRUCTION (cumulative sum of
    exponential gaps, then thinning): no 1e8-element sort."""
    rng = np.random.default_rng(seed)
    m = int(n * (1 + amp) * 1.02) + 1000
    t = np.cumsum(rng.exponential(T / m, m))
    keep = rng.uniform(0, 1 + amp, m) < 1 + amp * np.cos(2 * np.pi * f0 * t)
    t = t[keep][:n]
    assert t.size == n
    return t


def test_c5_full_event_count_long_baseline():
    """C5 at its BASELINE size: 1e8 events over ~1e8 s (1e9 pulse cycles), Z^2_2, nbin = 128, trial frequencies at the
    step 1/(2T) around the pulse.  The whole 4096-frequency block runs on the GPU; a subset of its trials is folded by
    the oracle at the full event count: statistics within 1e-9, profiles bit for bit (the fp64 phase has only ~23
    fractional bits left at 1e9 cycles -- any deviation from the reference's rounding sequence would move counts)."""
    from hendrics_b200 import z_n_search
    from hendrics_b200.search import fold_profiles

    T = 1e8
    ts = _poisson_times(100_000_000, T, seed=5)
    freqs = 1 / 0.101 + 5e-9 * (np.arange(4096) - 2048)
    f, stats = z_n_search(ts, freqs, nharm=2, nbin=128)
    assert stats.shape == (4096,) and np.all(np.isfinite(stats))
    t = _selected(ts)
    pick = np.array([0, 1, 63, 64, 2047, 2048, 2049, 4095])
    ref_s, ref_p = C.fold_search(t, freqs[pick], [0.0], 128, 2, want_profiles=True)
    np.testing.assert_allclose(stats[pick], ref_s[0], rtol=RTOL)
    s2, prof = fold_profiles(t, freqs[pick], [0.0], nbin=128, nharm=2)
    assert np.array_equal(prof, ref_p)
    np.testing.assert_allclose(s2[0], ref_s[0], rtol=RTOL)
    # the injected pulsation is recovered at the resolution of the grid
    assert abs(f[np.argmax(stats)] - 1 / 0.101) <= 2 * 5e-9
