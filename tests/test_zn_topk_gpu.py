"""K4 (unbinned Z^2_n) and top-k against the oracle."""
import numpy as np
import pytest

from oracle import hx_oracle as O

pytestmark = pytest.mark.gpu


def _unbinned_oracle(ts, freqs, fdots, n):
    out = np.zeros((len(fdots), len(freqs)))
    for i, fd in enumerate(fdots):
        for j, f in enumerate(freqs):
            ph = ts * f + 0.5 * ts * ts * fd
            ph -= np.floor(ph)
            out[i, j] = O.z_n_events(ph, n)
    return out


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_unbinned_zn(n):
    from hendrics_b200 import z_n_search_unbinned

    ts = O.simulate_events(20000, 1000.0, seed=21)
    freqs = 9.85 + 0.1 * np.arange(150) / 150
    fdots = np.array([0.0, 2e-6])
    ref = _unbinned_oracle(ts, freqs, fdots, n)
    fg, fdg, got = z_n_search_unbinned(ts, freqs, nharm=n, fdots=fdots)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-9)
    f1, got1 = z_n_search_unbinned(ts, freqs, nharm=n)
    np.testing.assert_allclose(got1, ref[0], rtol=1e-9)
    # binned statistic converges to the unbinned one (1 - sinc^2 correction, SURVEY finding 0.3)
    from hendrics_b200 import z_n_search
    _, zb = z_n_search(ts, freqs, nharm=n, nbin=1024)
    k = np.argmax(ref[0])
    assert abs(zb[k] - ref[0][k]) / ref[0][k] < 1e-2


def test_unbinned_known_answer():
    """hendrics/efsearch.py:387-392: phases drawn so that the profile is 1 + sin -> Z^2 = N/2 * ... ;
    here: all photons at one phase give Z^2_n = 2 n N exactly."""
    from hendrics_b200 import z_n_search_unbinned

    ts = np.arange(1000, dtype=np.float64)  # integer times, f = 1 -> phase 0 for every photon
    f, z = z_n_search_unbinned(ts, np.array([1.0]), nharm=3)
    assert np.isclose(z[0], 2 * 3 * 1000)


def test_unbinned_zn_event_splits_are_reproducible_and_negative_phases():
    """Few trials -> the event list is split across CTAs and the partial sums are added in split order: two runs give
    the same bits, and the result equals the oracle.  Negative fdots / times make the scaled phase negative (quadrant
    bits of a two's-complement rint)."""
    from hendrics_b200 import z_n_search_unbinned

    ts = O.simulate_events(300001, 3000.0, seed=5) - 1500.0  # odd count, times on both sides of 0
    freqs = np.array([9.9, 9.9009, 9.95])
    fdots = np.array([-3e-6, 0.0, 1e-6])
    ref = _unbinned_oracle(ts, freqs, fdots, 3)
    a = z_n_search_unbinned(ts, freqs, nharm=3, fdots=fdots)[2]
    b = z_n_search_unbinned(ts, freqs, nharm=3, fdots=fdots)[2]
    assert np.array_equal(a, b)
    np.testing.assert_allclose(a, ref, rtol=1e-9)


def test_unbinned_zn_long_baseline_uses_the_reference_s_rounded_phase():
    """T*f ~ 1e9 cycles: the fp64 phase carries ~1e-7 cycles of rounding noise, which the kernel must REPRODUCE (same
    products, same sum) to stay within 1e-9 of the reference -- an exactly reduced phase would differ by ~1e-6."""
    from hendrics_b200 import z_n_search_unbinned

    ts = O.simulate_events(50000, 1e8, seed=9)
    freqs = 9.9009 + 5e-9 * np.arange(64)
    ref = _unbinned_oracle(ts, freqs, np.array([0.0]), 2)
    _, got = z_n_search_unbinned(ts, freqs, nharm=2)
    np.testing.assert_allclose(got, ref[0], rtol=1e-9)


def test_unbinned_zn_argument_checks():
    from hendrics_b200 import _cabi, z_n_search_unbinned

    ts = np.arange(100.0)
    with pytest.raises(_cabi.HxError, match="nharm"):
        z_n_search_unbinned(ts, np.array([1.0]), nharm=9)
    with pytest.raises(_cabi.HxError):
        z_n_search_unbinned(np.array([0.0, np.nan, 2.0]), np.array([1.0]), nharm=2)
    with pytest.raises(_cabi.HxError, match="range"):
        z_n_search_unbinned(ts * 1e13, np.array([10.0]), nharm=2)
    ms, ops = None, None
    z_n_search_unbinned(ts, np.array([1.0, 2.0]), nharm=2, fdots=[0.0, 1e-9])
    ms, ops = _cabi.zn_last_kernel()
    assert ms > 0 and ops == 8 + 16 + 4 + 4


def test_topk_matches_oracle():
    import ctypes

    from hendrics_b200 import _cabi

    rng = np.random.default_rng(5)
    x = rng.chisquare(4, size=200001)
    x[[7, 99999, 150000]] = [300.0, 300.0, 299.0]  # exact tie: lower index first
    x[12345] = np.nan
    for k in (1, 5, 10, 64):
        idx = np.empty(k, dtype=np.int64)
        val = np.empty(k)
        n_out = ctypes.c_int64(0)
        rc = _cabi.lib().hx_topk(_cabi.dptr(x), x.size, k, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                 _cabi.dptr(val), ctypes.byref(n_out), 0)
        _cabi.check(rc)
        assert n_out.value == k
        xx = np.where(np.isnan(x), -np.inf, x)
        ref_idx, ref_val = O.topk_flat(xx, k)
        assert np.array_equal(idx, ref_idx)
        assert np.array_equal(val, ref_val)
    assert idx[0] == 7 and idx[1] == 99999 and idx[2] == 150000
