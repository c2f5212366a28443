"""GPU: the CUDA fast-folding path (csrc/hx_ffa.cu through hendrics_b200.ffa / efsearch.search_with_ffa) against golden
vectors of the REAL reference and against the oracle; blocked passes vs the per-step fallback."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SINGLE = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "ffa_single_*.npz")))
RTOL = 1e-9  # north_star tolerance for fp64 statistics (observed ~1e-14: only the summation order differs)


def load(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as d:
        return {k: d[k] for k in d.files}


def same(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    np.testing.assert_allclose(a[ok], b[ok], rtol=rtol, atol=0)


@pytest.fixture(autouse=True)
def _reset_tuning():
    from hendrics_b200 import _cabi

    yield
    _cabi.lib().hx_ffa_set_tuning(0)


@pytest.mark.parametrize("name", SINGLE)
@pytest.mark.parametrize("k", [0, 1, 2, 3, 7, -1])
def test_ffa_single_cuda_vs_reference(name, k):
    from hendrics_b200 import _cabi
    from hendrics_b200.ffa import ffa

    g = load(name)
    _cabi.lib().hx_ffa_set_tuning(k)
    per, st = ffa(g["counts"], int(g["bin_period"][0]))
    assert np.array_equal(per, g["periods"])
    same(st, g["stats"])
    ms, passes, steps, kk = _cabi.ffa_last_kernel()
    assert steps == int(np.log2(len(per)))
    if k == -1:
        assert passes == steps and kk == 0
    elif k > 0:
        assert passes == -(-steps // min(k, kk))


def test_ffa_search_cuda_vs_reference():
    from hendrics_b200.ffa import ffa_search

    g = load("ffa_search")
    dt, p0, p1 = g["params"]
    per, st = ffa_search(g["counts"], dt, p0, p1)
    assert np.array_equal(per, g["periods"])
    same(st, g["stats"])


@pytest.mark.parametrize("name", ["ffa_events", "ffa_events_nbin24"])
def test_search_with_ffa_cuda_vs_reference(name):
    from hendrics_b200.efsearch import search_with_ffa

    g = load(name)
    f0, f1, nbin, n = g["params"]
    fr, st, none, length = search_with_ffa(g["times"], f0, f1, nbin=int(nbin), n=int(n))
    assert none is None and length == g["length"][0]
    assert np.array_equal(fr, g["freqs"])
    same(st, g["stats"])


def test_hist1d_and_rebin_bit_exact():
    import torch

    from hendrics_b200 import _cabi
    from oracle import hx_oracle as ho

    rng = np.random.default_rng(5)
    x = np.sort(rng.uniform(-3.0, 1003.0, 200_001))  # some events outside the range on both sides
    x[1000] = 1000.0  # exactly on the (excluded) right edge
    x[7] = 0.0        # exactly on the left edge
    bins = 12_345
    H = ho.histogram(x, bins=bins, range=[0.0, 1000.0])
    xd = torch.from_numpy(x).cuda()
    Hd = torch.empty(bins, dtype=torch.int32, device="cuda")
    _cabi.check(_cabi.lib().hx_hist1d_dev(xd.data_ptr(), xd.numel(), 0.0, 0.0, 1000.0, bins, Hd.data_ptr(), 0))
    torch.cuda.synchronize()
    assert np.array_equal(Hd.cpu().numpy().astype(np.int64), H.astype(np.int64))
    for r in (2, 3, 8):
        out = torch.empty(bins // r, dtype=torch.int32, device="cuda")
        _cabi.check(_cabi.lib().hx_rebin_u32_dev(Hd.data_ptr(), bins, r, out.data_ptr(), 0))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().astype(np.float64), ho._quick_rebin(H, r))


@pytest.mark.parametrize("bp,n_raw", [(16, 16 * 5000 + 3), (131, 131 * 1000), (1000, 50_000), (20000, 20000 * 9 + 17), (5, 11), (7, 7), (9, 5)])
def test_ffa_cuda_vs_oracle_shapes(bp, n_raw):
    """ragged lengths, a period that only fits the per-step fallback (20000 bins), two tables, and the degenerate
    single-table case (hendrics/ffa.py:341-342)"""
    from hendrics_b200.ffa import ffa
    from oracle import hx_oracle as ho

    rng = np.random.default_rng(bp)
    counts = rng.poisson(1.5, n_raw).astype(np.float64)
    per, st = ffa(counts, bp)
    per_o, st_o = ho.ffa(counts, bp)
    assert np.array_equal(per, per_o)
    same(st, st_o)


def test_ffa_host_pointer_entry_and_errors():
    import ctypes

    from hendrics_b200 import _cabi
    from hendrics_b200.ffa import _ffa_tables, ffa
    from oracle import hx_oracle as ho

    rng = np.random.default_rng(1)
    counts = rng.poisson(2.0, 4000).astype(np.uint32)
    bp, ntab = 32, 128
    tab = np.ascontiguousarray(_ffa_tables(bp, 2))
    st = np.empty(ntab)
    _cabi.check(_cabi.lib().hx_ffa(counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), counts.size, bp, 2, _cabi.dptr(tab),
                                   _cabi.dptr(st), ntab, 0))
    same(st, ho.ffa(counts.astype(np.float64), bp)[1])
    # ntables not a power of two / too small for the light curve
    for bad in (100, 64):
        rc = _cabi.lib().hx_ffa(counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), counts.size, bp, 2, _cabi.dptr(tab),
                                _cabi.dptr(st), bad, 0)
        assert rc == _cabi.HX_ERR_ARG
    with pytest.raises(ValueError):
        ffa(np.array([1.0, 2.5, 3.0] * 100), 16)  # fractional counts
    with pytest.raises(ValueError):
        ffa(np.array([1.0, -2.0, 3.0] * 100), 16)


def test_ffa_large_light_curve_conserves_counts_and_matches_fold():
    """size-independent properties at a realistic size (2^17 tables x 128 bins): the zero-drift table of the FFA is the
    plain fold of the light curve at the integer period, and every table holds all the counts (checked through Z^2_n of
    a flat-vs-pulsed light curve: the injected period gives the largest statistic)."""
    from hendrics_b200.ffa import ffa
    from oracle import hx_oracle as ho

    rng = np.random.default_rng(9)
    P, ntab = 128, 1 << 17
    n_raw = P * (ntab - 1000) + 77
    t = np.arange(n_raw)
    true_period = P + 300.25 / (ntab - 1)
    counts = rng.poisson(0.5 * (1 + 0.6 * np.cos(2 * np.pi * t / true_period))).astype(np.float64)
    per, st = ffa(counts, P)
    assert per.size == ntab
    assert abs(per[np.nanargmax(st)] - true_period) < 2.0 / (ntab - 1)
    # table 0 (no drift) == fold at exactly P bins: Z^2_2 of the folded profile
    prof = np.zeros(P)
    np.add.at(prof, t % P, counts)
    np.testing.assert_allclose(st[0], ho.z_n_fast_cached(prof, 2), rtol=RTOL)


def test_ffa_randomized_periods_lengths_and_pass_splits():
    """random bin periods (aligned / unaligned, tiny, > 128), ragged light-curve lengths, random steps-per-pass overrides,
    nharm 1..4 (tensor-core epilogue for <= 3, plain epilogue for 4): always the oracle's numbers"""
    from hendrics_b200 import _cabi
    from hendrics_b200.ffa import ffa
    from oracle import hx_oracle as ho

    rng = np.random.default_rng(int(os.environ.get("HX_TEST_SEED", 2024)))
    lib = _cabi.lib()
    for trial in range(24):
        P = int(rng.choice([2, 3, 5, 8, 17, 31, 32, 33, 64, 100, 127, 128, 129, 130, 131, 200, 257, 300]))
        ntab_target = int(2 ** rng.integers(1, 11))
        n_raw = int(P * ntab_target - rng.integers(0, P))
        counts = rng.poisson(rng.choice([0.05, 1.0, 20.0]), max(n_raw, P + 1)).astype(np.float64)
        if counts.sum() == 0:  # every final table holds ALL the counts: an empty light curve divides by zero (the
            counts[0] = 1.0    # reference's numba code raises ZeroDivisionError there, hx_ffa returns NaN)
        z_n_n = int(rng.integers(1, 5))
        lib.hx_ffa_set_tuning(int(rng.choice([0, 1, 2, 3, 4, 5, 6, 8, -1])))
        per, st = ffa(counts, P, z_n_n=z_n_n)
        per_o, st_o = ho.ffa(counts, P, z_n_n=z_n_n)
        assert np.array_equal(per, per_o), (trial, P)
        # rtol 1e-9; atol 1e-9 for statistics that cancel to ~0 (e.g. two equal bins at P = 2: c_1 = k - k)
        assert np.array_equal(np.isnan(st), np.isnan(st_o)), (trial, P)
        ok = ~np.isnan(st_o)
        np.testing.assert_allclose(st[ok], st_o[ok], rtol=RTOL, atol=1e-9, err_msg=f"trial {trial} P={P} n={z_n_n}")
