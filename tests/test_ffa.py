"""CPU: the FFA oracle (oracle/hx_oracle.py) against golden vectors produced by the REAL reference code
(hendrics/ffa.py, hendrics/efsearch.py:search_with_ffa; tests/golden/make_golden.py section 8)."""
import glob
import os

import numpy as np
import pytest

from oracle import hx_oracle as ho

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SINGLE = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "ffa_single_*.npz")))


def load(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as d:
        return {k: d[k] for k in d.files}


def same(a, b, rtol):
    """equal NaN pattern (all-zero tables) and values within rtol"""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    np.testing.assert_allclose(a[ok], b[ok], rtol=rtol, atol=0)


def test_ffa_index_helpers_known_answers():
    # doctests of hendrics/ffa.py:244-256
    assert ho.ffa_start_value(0, 0) == 0 and ho.ffa_start_value(4, 0) == 4 and ho.ffa_start_value(5, 2) == 2
    assert ho.ffa_start_value(8, 1) == 8 and ho.ffa_start_value(7, 1) == 5 and ho.ffa_start_value(10, 2) == 9
    # the table in the module docstring of hendrics/ffa.py:9-25: step 1, prof 3 = 1 + 3'' ; step 2, prof 7 = 3 + 7''''
    assert (ho.ffa_start_value(3, 1), ho.ffa_shift(3, 1)) == (1, 2)
    assert (ho.ffa_start_value(7, 2), ho.ffa_shift(7, 2)) == (3, 4)


@pytest.mark.parametrize("name", SINGLE)
def test_ffa_single_oracle_vs_reference(name):
    g = load(name)
    per, st = ho.ffa(g["counts"], int(g["bin_period"][0]))
    assert np.array_equal(per, g["periods"])
    same(st, g["stats"], 1e-12)


def test_ffa_search_oracle_vs_reference():
    g = load("ffa_search")
    dt, p0, p1 = g["params"]
    per, st = ho.ffa_search(g["counts"], dt, p0, p1)
    assert np.array_equal(per, g["periods"])
    same(st, g["stats"], 1e-12)


@pytest.mark.parametrize("name", ["ffa_events", "ffa_events_nbin24"])
def test_search_with_ffa_oracle_vs_reference(name):
    g = load(name)
    f0, f1, nbin, n = g["params"]
    fr, st, _, length = ho.search_with_ffa(g["times"], f0, f1, nbin=int(nbin), n=int(n))
    assert np.array_equal(fr, g["freqs"]) and length == g["length"][0]
    same(st, g["stats"], 1e-12)


def test_ffa_recovers_the_pulse():
    # behaviour the reference tests for (hendrics/tests/test_ffa.py:10-31): the best period is the injected one
    g = load("ffa_events")
    best = g["freqs"][np.nanargmax(g["stats"])]
    assert abs(best - 1 / 0.101) < 5e-3


def test_ffa_tables_reproduce_the_arange_quirk():
    # np.arange(0, 1, 1/49) has 50 elements: harmonic 2 of the reference reads a skewed table for such periods
    assert len(np.arange(0, 1, 1 / 49)) == 50
    tab = ho.ffa_tables(49, 2)
    j = np.arange(49)
    np.testing.assert_allclose(tab[0, 0], np.cos(2 * np.pi * j / 49), atol=1e-15)
    assert not np.allclose(tab[0, 1], np.cos(4 * np.pi * j / 49))  # the quirk
    tab = ho.ffa_tables(37, 2)
    np.testing.assert_allclose(tab[0, 1], np.cos(4 * np.pi * np.arange(37) / 37), atol=1e-14)


def test_host_profile_helpers_known_answers():
    """doctests of hendrics/ffa.py:126-131, 179-186, 200-206 on the host mirror (no GPU involved), and agreement with the
    oracle's restatement on a random profile"""
    from hendrics_b200.efsearch import h_test as h_test_reexported
    from hendrics_b200.ffa import h_test, z_n_fast_cached, z_n_fast_cached_all

    phase = 2 * np.pi * np.arange(0, 1, 0.01)
    norm = np.sin(phase) + 1
    assert np.isclose(z_n_fast_cached(norm, n=4), 50) and np.isclose(z_n_fast_cached(norm, n=2), 50)
    ks, zs = z_n_fast_cached_all(norm, nmax=4)
    assert ks[1] == 2 and np.isclose(zs[1], 50) and np.isclose(zs[3], 50)
    h, m = h_test(norm, nmax=4)
    assert np.isclose(h, 50) and m == 1 and h_test_reexported is h_test
    prof = np.random.default_rng(3).poisson(30, 64).astype(float)
    np.testing.assert_allclose(z_n_fast_cached(prof, 3), ho.z_n_fast_cached(prof, 3), rtol=1e-13)
    np.testing.assert_allclose(z_n_fast_cached_all(prof, 5)[1], ho.z_n_fast_cached_all(prof, 5)[1], rtol=1e-13)
    assert h_test(prof, nmax=5)[1] == ho.h_test(prof, nmax=5)[1]


def test_host_tables_and_period_grid_match_the_oracle():
    """host-side pieces of hendrics_b200.ffa that need no GPU: the harmonic tables handed to hx_ffa (including the periods
    where np.arange overshoots) and the trial-period grid of hendrics/ffa.py:314"""
    from hendrics_b200.ffa import _ffa_tables, _period_fractions

    for bp in (2, 16, 37, 49, 98, 103, 128, 250):
        for n in (1, 2, 4):
            assert np.array_equal(_ffa_tables(bp, n), ho.ffa_tables(bp, n)), (bp, n)
    for ntab in (2, 8, 1024):
        ref = np.array([5 + n / (ntab - 1) for n in range(ntab)])
        assert np.array_equal(5 + _period_fractions(ntab), ref)
        assert not _period_fractions(ntab).flags.writeable  # cached: must not be modified in place


def test_ffa_oracle_properties():
    """size-independent properties of the fast folding algorithm (the same ones tests/test_ffa_gpu.py uses at scale):
    table 0 is the plain fold at P bins, every table keeps all the counts, the last table of bin period P is the
    zero-drift table of P + 1 folded from the same data, and the search grid is increasing within a bin period"""
    rng = np.random.default_rng(12)
    P, n_raw = 24, 24 * 200 + 7
    counts = rng.poisson(3.0, n_raw).astype(np.float64)
    ntab = int(2 ** np.ceil(np.log2(n_raw // P + 1)))
    arr = np.zeros(ntab * P)
    arr[:n_raw] = counts
    arr = arr.reshape(ntab, P)
    for step in range(int(np.log2(ntab))):
        arr = ho.ffa_step(arr, step, ntab)
    assert np.all(arr.sum(axis=1) == counts.sum())          # every table holds all the counts
    prof = np.zeros(P)
    np.add.at(prof, np.arange(n_raw) % P, counts)
    assert np.array_equal(arr[0], prof)                      # no drift: the plain fold
    # maximal drift: one bin per table row, i.e. row r is rolled by -r -> folding at P + 1 bins, seen modulo P
    prof1 = np.zeros(P)
    idx = np.arange(ntab * P)
    row, col = idx // P, idx % P
    padded = np.zeros(ntab * P)
    padded[:n_raw] = counts
    np.add.at(prof1, (col - row) % P, padded)
    assert np.array_equal(arr[ntab - 1], prof1)
    per, st = ho.ffa(counts, P)
    assert per[0] == P and per[-1] == P + 1 and np.all(np.diff(per) > 0)
    np.testing.assert_allclose(st[0], ho.z_n_fast_cached(prof, 2), rtol=1e-12)
    # linearity of the tables in the light curve
    c2 = rng.poisson(1.0, n_raw).astype(np.float64)
    a1 = np.zeros(ntab * P); a1[:n_raw] = c2; a1 = a1.reshape(ntab, P)
    a12 = np.zeros(ntab * P); a12[:n_raw] = counts + c2; a12 = a12.reshape(ntab, P)
    for step in range(int(np.log2(ntab))):
        a1 = ho.ffa_step(a1, step, ntab)
        a12 = ho.ffa_step(a12, step, ntab)
    assert np.array_equal(a12, arr + a1)


def _blocked_ffa(counts, P, ks):
    """numpy restatement of csrc/hx_ffa.cu's index arithmetic: passes of k fused steps on gathered working sets, the shift
    offset hs = (h * 2^sigma) mod P doubled from step to step, output tables folded in pairs (2m, 2m+1 share both inputs,
    shifts one bin apart)."""
    n_raw = len(counts)
    ntables = int(2 ** np.ceil(np.log2(n_raw // P + 1)))
    A = np.zeros(ntables * P, dtype=np.int64)
    A[:n_raw] = counts
    A = A.reshape(ntables, P)
    assert sum(ks) == int(np.log2(ntables))
    s0 = 0
    for k in ks:
        out = np.zeros_like(A)
        for c in range(ntables >> k):
            h = c & ((1 << s0) - 1)
            G = (c >> s0) << (s0 + k)
            B = np.stack([A[G + (q << s0) + h] for q in range(1 << k)])
            hs = h % P
            for sg in range(k):
                B2 = np.zeros_like(B)
                for r in range(0, 1 << k, 2):
                    mu = r & ((2 << sg) - 1)
                    i1 = (r - mu) + (mu >> 1)
                    i2 = i1 + (1 << sg)
                    sh = (hs + (mu >> 1)) % P
                    B2[r] = B[i1] + np.roll(B[i2], -sh)
                    B2[r + 1] = B[i1] + np.roll(B[i2], -(sh + 1))
                B = B2
                hs = (2 * hs) % P
            out[c << k:(c + 1) << k] = B
        A = out
        s0 += k
    return A


@pytest.mark.parametrize("P,n_raw,ks", [(7, 7 * 60 + 3, (2, 2, 2)), (16, 16 * 200 + 5, (3, 3, 2)), (5, 5 * 31 + 1, (1, 4)),
                                        (5, 5 * 31 + 1, (5,)), (9, 9 * 100, (4, 3)), (3, 3 * 500, (1,) * 9), (129, 129 * 60, (3, 3))])
def test_blocked_pairwise_butterfly_equals_the_reference_steps(P, n_raw, ks):
    rng = np.random.default_rng(P)
    counts = rng.poisson(3.0, n_raw)
    ntables = int(2 ** np.ceil(np.log2(n_raw // P + 1)))
    arr = np.zeros(ntables * P)
    arr[:n_raw] = counts
    arr = arr.reshape(ntables, P)
    for step in range(int(np.log2(ntables))):
        arr = ho.ffa_step(arr, step, ntables)
    assert np.array_equal(_blocked_ffa(counts, P, ks), arr.astype(np.int64))
