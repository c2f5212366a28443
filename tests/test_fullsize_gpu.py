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
