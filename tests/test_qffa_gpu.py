"""K2/K3 parity: `--fast` 2-D profiles bit-exact, statistics rel 1e-9 vs the oracle."""
import ctypes

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import hx_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _centred(n, T, seed):
    ts = O.simulate_events(n, T, seed=seed)
    return ts - (ts.max() + ts.min()) / 2


def _host_qffa(ts, cents, nbin, nprof, L, Q, n):
    from hendrics_b200 import _cabi

    ts = _cabi.as_f64(ts)
    cents = np.asarray(cents, dtype=np.float64).reshape(-1, 3)
    mf, mfd, mfdd = (np.ascontiguousarray(cents[:, i]) for i in range(3))
    Lf, Qf = _cabi.as_f64(L.ravel()), _cabi.as_f64(Q.ravel())
    stats = np.empty((cents.shape[0], Lf.size))
    prof = np.empty((cents.shape[0], nprof, nbin), dtype=np.uint32)
    rc = _cabi.lib().hx_qffa_search(
        _cabi.dptr(ts), ts.size, _cabi.dptr(mf), _cabi.dptr(mfd), _cabi.dptr(mfdd), cents.shape[0], nbin, nprof,
        _cabi.dptr(Lf), _cabi.dptr(Qf), Lf.size, n, _cabi.dptr(stats),
        prof.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), 0)
    _cabi.check(rc)
    return stats.reshape((cents.shape[0],) + L.shape), prof


@pytest.mark.parametrize("nbin,nprof,n,mode", [(16, 128, 1, 0), (32, 64, 2, 1), (128, 2048, 2, 0), (64, 100, 3, 2),
                                               (10, 37, 1, 1)])
def test_qffa_step_matches_oracle(nbin, nprof, n, mode):
    ts = _centred(40000, 2000.0, seed=nbin)
    cents = []
    for i in range(11):  # more than one fused group of centres
        f = 9.88 + 0.004 * i
        fd = 0.0 if mode == 0 else 1e-7 * (i - 3)
        fdd = 0.0 if mode < 2 else 1e-11 * (i + 1)
        cents.append((f, fd, fdd))
    lin = np.linspace(-nbin * 2, nbin * 2, 16)
    L, Q = np.meshgrid(lin, lin, indexing="ij")
    stats, prof = _host_qffa(ts, cents, nbin, nprof, L, Q, n)
    for ic, (f, fd, fdd) in enumerate(cents):
        ref_p = C.qffa_profiles(ts, f, fd, fdd, nbin, nprof)
        assert np.array_equal(prof[ic].astype(np.uint64), ref_p), f"centre {ic}: 2-D profile not bit-exact"
        ref_s = C.fast_step(ref_p, L, Q, n)
        np.testing.assert_allclose(stats[ic], ref_s, rtol=RTOL)
    # the python oracle agrees with the C oracle on the first centre (numba path of the reference)
    f, fd, fdd = cents[0]
    py = O.search_with_qffa_step(ts, f, fd, fdd, nbin=nbin, nprof=nprof, npfact=2, oversample=8, n=n,
                                 return_profiles=True)
    assert np.array_equal(py[3], prof[0].astype(np.uint64))
    np.testing.assert_allclose(stats[0], py[2], rtol=RTOL)


def test_qffa_last_event_dropped_and_gaps():
    """Right edge excluded (base.py:674) and a long data gap (rows outside the smem window)."""
    rng = np.random.default_rng(3)
    ts = np.sort(np.concatenate([rng.uniform(0, 100, 9000), rng.uniform(5000, 5100, 9000)]))
    ts = ts - (ts.max() + ts.min()) / 2
    lin = np.linspace(-32, 32, 8)
    L, Q = np.meshgrid(lin, lin, indexing="ij")
    stats, prof = _host_qffa(ts, [(3.3, 0, 0)], 16, 128, L, Q, 2)
    ref_p = C.qffa_profiles(ts, 3.3, 0, 0, 16, 128)
    assert np.array_equal(prof[0].astype(np.uint64), ref_p)
    assert prof[0].sum() == ts.size - 1  # t == t[-1] is not counted
    np.testing.assert_allclose(stats[0], C.fast_step(ref_p, L, Q, 2), rtol=RTOL)


def test_search_with_qffa_api():
    from hendrics_b200.efsearch import search_with_qffa, search_with_qffa_step

    ts = O.simulate_events(30000, 1000.0, seed=9) + 8e7
    np.random.seed(0)
    ref = O.search_with_qffa(ts, 9.85, 9.95, nbin=16, n=2, oversample=8)
    got = search_with_qffa(ts, 9.85, 9.95, nbin=16, n=2, oversample=8, silent=True)
    assert len(got) == 6
    for a, b in zip(got[:2], ref[:2]):
        assert np.array_equal(a, b)
    np.testing.assert_allclose(got[2], ref[2], rtol=RTOL)
    assert got[3] == ref[3] and got[4] == ref[4] and got[5] == ref[5]
    # best candidate close to the injected pulse
    i = np.unravel_index(np.argmax(got[2]), got[2].shape)
    assert abs(got[0][i] - 1 / 0.101) < 2 / 1000.0
    # no fdot search -> 4-tuple of 1-D arrays
    ref4 = O.search_with_qffa(ts, 9.85, 9.95, nbin=16, n=1, search_fdot=False)
    got4 = search_with_qffa(ts, 9.85, 9.95, nbin=16, n=1, search_fdot=False)
    assert len(got4) == 4 and got4[0].ndim == 1
    assert np.array_equal(got4[0], ref4[0])
    np.testing.assert_allclose(got4[1], ref4[1], rtol=RTOL)
    # single step API
    tc = (ts - (ts.max() + ts.min()) / 2).astype(np.float64)
    r = O.search_with_qffa_step(tc, 9.9, nbin=16, nprof=128, n=2, return_profiles=True)
    g = search_with_qffa_step(tc, 9.9, nbin=16, nprof=128, n=2, return_profiles=True)
    assert np.array_equal(g[0], r[0]) and np.array_equal(g[1], r[1])
    assert np.array_equal(g[3].astype(np.uint64), r[3])
    np.testing.assert_allclose(g[2], r[2], rtol=RTOL)


def test_lane_centre_kernel_equals_lane_event_kernel():
    """The two K2 implementations (lane = centre, default; lane = event, fallback) give identical 2-D
    profiles; also with many centres (several warp slots, ragged last slot) and a non-zero fdot/fddot."""
    from hendrics_b200 import _cabi

    ts = _centred(60000, 3000.0, seed=77)
    lin = np.linspace(-64, 64, 8)
    L, Q = np.meshgrid(lin, lin, indexing="ij")
    lib = _cabi.lib()
    for nbin, nprof, mode in ((32, 256, 0), (128, 512, 1), (16, 128, 2)):
        cents = [(9.7 + 0.003 * i, 0.0 if mode == 0 else 2e-8 * (i % 5 - 2), 0.0 if mode < 2 else 3e-12) for i in range(75)]
        try:
            lib.hx_qffa_set_tuning(0)
            s_lane, p_lane = _host_qffa(ts, cents, nbin, nprof, L, Q, 2)
            lib.hx_qffa_set_tuning(1008)
            s_ev, p_ev = _host_qffa(ts, cents, nbin, nprof, L, Q, 2)
        finally:
            lib.hx_qffa_set_tuning(0)
        assert np.array_equal(p_lane, p_ev), (nbin, nprof, mode)
        np.testing.assert_allclose(s_lane, s_ev, rtol=1e-11)
        ref_p = C.qffa_profiles(ts, *cents[37], nbin, nprof)
        assert np.array_equal(p_lane[37].astype(np.uint64), ref_p)


def test_qffa_unsorted_times_fall_back_and_stay_exact():
    """Unsorted input: the reference still histograms every photon (only its range uses t[0], t[-1]);
    the lane = centre path detects it on the device and the lane = event kernel takes over."""
    rng = np.random.default_rng(8)
    ts = _centred(20000, 1000.0, seed=5)
    perm = rng.permutation(ts.size - 2) + 1
    tu = np.concatenate([ts[:1], ts[perm], ts[-1:]])  # keep first/last so that the range is unchanged
    lin = np.linspace(-32, 32, 8)
    L, Q = np.meshgrid(lin, lin, indexing="ij")
    stats, prof = _host_qffa(tu, [(9.9, 0, 0), (9.91, 1e-7, 0)], 16, 128, L, Q, 2)
    for ic, c in enumerate([(9.9, 0, 0), (9.91, 1e-7, 0)]):
        ref_p = C.qffa_profiles(tu, *c, 16, 128)
        assert np.array_equal(prof[ic].astype(np.uint64), ref_p)
        np.testing.assert_allclose(stats[ic], C.fast_step(ref_p, L, Q, 2), rtol=RTOL)


def test_qffa_photons_near_t0_are_fixed_up():
    """Photons with |t| < 1/f_min take the reference's rounding (phase - floor(phase) is inexact there):
    put many photons inside that window and compare with the oracle, several centres."""
    rng = np.random.default_rng(9)
    ts = np.sort(np.concatenate([rng.uniform(-500, 500, 20000), rng.uniform(-0.4, 0.4, 3000)]))
    cents = [(2.5 + 0.01 * i, -1e-7 * i, 0.0) for i in range(40)]
    lin = np.linspace(-32, 32, 4)
    L, Q = np.meshgrid(lin, lin, indexing="ij")
    stats, prof = _host_qffa(ts, cents, 64, 128, L, Q, 2)
    for ic in (0, 13, 39):
        ref_p = C.qffa_profiles(ts, *cents[ic], 64, 128)
        assert np.array_equal(prof[ic].astype(np.uint64), ref_p), ic


def test_qffa_profile_batches_are_transparent(monkeypatch):
    """Centre batching (the 2 GiB profile-scratch bound, forced down to 1 MiB here: 75 centres in 3 batches) changes
    nothing: identical 2-D profiles and statistics."""
    ts = _centred(40000, 2000.0, seed=91)
    lin = np.linspace(-64, 64, 8)
    L, Q = np.meshgrid(lin, lin, indexing="ij")
    cents = [(9.7 + 0.003 * i, 1e-8 * (i % 3 - 1), 0.0) for i in range(75)]
    monkeypatch.delenv("HX_QFFA_BATCH_MIB", raising=False)
    s_one, p_one = _host_qffa(ts, cents, 32, 256, L, Q, 2)
    monkeypatch.setenv("HX_QFFA_BATCH_MIB", "1")
    s_many, p_many = _host_qffa(ts, cents, 32, 256, L, Q, 2)
    assert np.array_equal(p_one, p_many)
    assert np.array_equal(s_one, s_many)


@pytest.mark.parametrize("nbin,nprof,n,mode,n_ev", [(128, 2048, 2, 0, 300_000), (128, 8, 4, 1, 400_000), (64, 512, 1, 2, 200_000),
                                                    (16, 64, 3, 0, 50_001), (32, 4, 2, 1, 1_000_000), (128, 256, 5, 0, 60_000)])
def test_fused_coefficient_drain_equals_the_profile_path(nbin, nprof, n, mode, n_ev):
    """Without a `profiles` output K2 reduces every finished time slice in the SM to its harmonic coefficients (fp64
    accumulators, packed counters emptied several times per slice, slices shared between CTAs through fp64 REDs); with
    one, the counts go to the global scratch and K3a computes the same coefficients from it.  Same statistics either
    way (the summation order differs: 1e-12), and both within 1e-9 of the oracle's shift-and-sum on the bit-exact
    profiles.  nprof = 4 / 8 with 1e6 / 4e5 events puts several CTAs on every slice; n = 5 exceeds the fused path's
    harmonics and takes the scratch path on both sides."""
    from hendrics_b200.efsearch import _qffa_device_call, _shift_grids, _to_device

    ts = _centred(n_ev, 5000.0, seed=nbin + nprof)
    cents = [(9.9 + 0.0007 * k, (3e-8 if mode >= 1 else 0.0) * (k % 3), (2e-12 if mode >= 2 else 0.0)) for k in range(70)]
    L, Q = _shift_grids(nbin, 2, 4, True)
    td = _to_device(ts)
    fused = _qffa_device_call(td, ts[0], ts[-1], cents, nbin, nprof, L, Q, n)
    scratch, prof = _qffa_device_call(td, ts[0], ts[-1], cents, nbin, nprof, L, Q, n, want_profiles=True)
    np.testing.assert_allclose(fused, scratch, rtol=1e-12, atol=1e-12 * np.abs(scratch).max())
    again = _qffa_device_call(td, ts[0], ts[-1], cents, nbin, nprof, L, Q, n)
    np.testing.assert_allclose(again, fused, rtol=1e-13, atol=1e-13 * np.abs(fused).max())
    for k in (0, 33, 69):
        ref_p = C.qffa_profiles(ts, *cents[k], nbin, nprof)
        assert np.array_equal(prof[k].astype(np.uint64), ref_p)
        np.testing.assert_allclose(fused[k], C.fast_step(ref_p, L, Q, n), rtol=RTOL)


def test_fused_drain_with_whole_slices_per_cta_is_bit_reproducible():
    """The C4 shape in small: 1280 centres (5 CTA columns) and 2048 time slices -- many more slices than CTAs along the
    event axis, so every CTA folds whole slices and each slice's coefficients are summed by one thread in program order:
    two runs give the same bits (and the profile path the same numbers to rounding)."""
    from hendrics_b200.efsearch import _qffa_device_call, _shift_grids, _to_device

    ts = _centred(250_001, 5000.0, seed=77)
    ts = ts[~((ts > 300.0) & (ts < 900.0))]  # a data gap: ~250 empty time slices in a row, CTAs with nothing to fold
    cents = [(9.85 + 0.00008 * k, 0.0, 0.0) for k in range(1280)]
    L, Q = _shift_grids(128, 2, 8, True)
    td = _to_device(ts)
    a = _qffa_device_call(td, ts[0], ts[-1], cents, 128, 2048, L, Q, 2)
    b = _qffa_device_call(td, ts[0], ts[-1], cents, 128, 2048, L, Q, 2)
    assert np.array_equal(a, b)
    scratch, prof = _qffa_device_call(td, ts[0], ts[-1], cents[:64], 128, 2048, L, Q, 2, want_profiles=True)
    np.testing.assert_allclose(a[:64], scratch, rtol=1e-12, atol=1e-12 * np.abs(scratch).max())
    ref_p = C.qffa_profiles(ts, *cents[17], 128, 2048)
    assert np.array_equal(prof[17].astype(np.uint64), ref_p)
    np.testing.assert_allclose(a[17], C.fast_step(ref_p, L, Q, 2), rtol=RTOL)


def test_streamed_upload_from_host_memory(monkeypatch):
    """hx_qffa_search on a host array of >= 2^22 events uploads it in chunks while K2 already folds the chunks that have
    arrived (the kernel waits on a progress word written by the copy stream).  Same bits as the plain upload, from
    pageable and from page-locked memory (the two issue orders), and an unsorted list -- detected on the device only after
    the optimistic run -- still gives the right answer through the fallback."""
    import torch

    from hendrics_b200.efsearch import _shift_grids
    from hendrics_b200.fastfold import _qffa_host_call

    rng = np.random.default_rng(5)
    ts = np.sort(rng.uniform(-2500.0, 2500.0, 4_500_001))
    cents = [(9.85 + 0.0001 * k, 0.0, 0.0) for k in range(96)]
    L, Q = _shift_grids(64, 2, 4, True)
    monkeypatch.setenv("HX_QFFA_STREAM", "0")
    plain = _qffa_host_call(ts, cents, 64, 1024, L, Q, 2)
    monkeypatch.setenv("HX_QFFA_STREAM", "1")
    pageable = _qffa_host_call(ts, cents, 64, 1024, L, Q, 2)
    pinned = torch.from_numpy(ts).pin_memory().numpy()
    locked = _qffa_host_call(pinned, cents, 64, 1024, L, Q, 2)
    np.testing.assert_allclose(pageable, plain, rtol=1e-13, atol=1e-13 * np.abs(plain).max())
    np.testing.assert_allclose(locked, plain, rtol=1e-13, atol=1e-13 * np.abs(plain).max())
    ref_p = C.qffa_profiles(ts, *cents[40], 64, 1024)
    np.testing.assert_allclose(pageable[40], C.fast_step(ref_p, L, Q, 2), rtol=RTOL)
    # unsorted: two distant blocks swapped
    bad = ts.copy()
    bad[1000:2000], bad[3_000_000:3_001_000] = ts[3_000_000:3_001_000].copy(), ts[1000:2000].copy()
    got = _qffa_host_call(bad, cents[:8], 64, 1024, L, Q, 2)
    ref_b = C.qffa_profiles(bad, *cents[3], 64, 1024)
    np.testing.assert_allclose(got[3], C.fast_step(ref_b, L, Q, 2), rtol=RTOL)
