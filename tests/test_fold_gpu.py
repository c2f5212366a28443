"""K1 parity: CUDA fold search vs the CPU oracle (profiles bit-exact, statistics rel 1e-9)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import hx_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-9  # north_star tolerance for Z^2_n / chi^2 in fp64


def _events(n, T, seed):
    return O.simulate_events(n, T, seed=seed)


def _host_call(ts, freqs, fdots, nbin, nharm, flags=0, want_profiles=True):
    from hendrics_b200 import _cabi

    ts = _cabi.as_f64(ts)
    freqs = _cabi.as_f64(np.atleast_1d(freqs))
    fdots = _cabi.as_f64(np.atleast_1d(fdots))
    stats = np.empty((fdots.size, freqs.size))
    prof = np.empty((fdots.size, freqs.size, nbin), dtype=np.uint32) if want_profiles else None
    rc = _cabi.lib().hx_fold_search(
        _cabi.dptr(ts), ts.size, _cabi.dptr(freqs), freqs.size, _cabi.dptr(fdots), fdots.size, nbin, nharm, flags,
        _cabi.dptr(stats), prof.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)) if want_profiles else None, 0)
    _cabi.check(rc)
    return stats, prof


def _assert_stats(got, ref):
    good = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), good)
    np.testing.assert_allclose(got[good], ref[good], rtol=RTOL, atol=0)


CASES = [
    # n_ev, T, nbin, nharm, n_f, fdots
    (20000, 1000.0, 128, 2, 300, [0.0]),
    (20000, 1000.0, 16, 0, 500, [0.0]),
    (33333, 1000.0, 64, 4, 257, [-2e-6, 0.0, 3e-6]),
    (5001, 300.0, 32, 1, 70, [1e-5]),
    (20000, 1000.0, 8, 0, 100, [0.0, 1e-6]),
    (4097, 1000.0, 256, 3, 65, [0.0]),
    (3000, 1000.0, 1024, 2, 33, [0.0, 2e-6]),
]


@pytest.mark.parametrize("n_ev,T,nbin,nharm,n_f,fdots", CASES)
@pytest.mark.parametrize("flags", [0, 2])
def test_fold_matches_oracle(n_ev, T, nbin, nharm, n_f, fdots, flags):
    ts = _events(n_ev, T, seed=n_ev % 97)
    freqs = 9.85 + 0.1 * np.arange(n_f) / n_f
    ref_s, ref_p = C.fold_search(ts, freqs, fdots, nbin, nharm, want_profiles=True)
    want_prof = flags == 2
    got_s, got_p = _host_call(ts, freqs, fdots, nbin, nharm, flags=flags, want_profiles=want_prof)
    if want_prof:
        assert np.array_equal(got_p.astype(np.int64), ref_p), "profile counts must be bit-exact"
    _assert_stats(got_s, ref_s)


@pytest.mark.parametrize("nbin,nharm", [(10, 0), (100, 2), (48, 3), (7, 1)])
def test_fold_non_power_of_two_nbin(nbin, nharm):
    ts = _events(15000, 800.0, seed=3)
    freqs = 9.8 + 0.2 * np.arange(130) / 130
    fdots = [0.0, 1e-6]
    ref_s, ref_p = C.fold_search(ts, freqs, fdots, nbin, nharm, want_profiles=True)
    got_s, got_p = _host_call(ts, freqs, fdots, nbin, nharm)
    assert np.array_equal(got_p.astype(np.int64), ref_p)
    _assert_stats(got_s, ref_s)


def test_forced_generic_equals_fast_path():
    ts = _events(20000, 1000.0, seed=5)
    freqs = 9.85 + 0.1 * np.arange(200) / 200
    fdots = [0.0, 4e-6]
    a_s, a_p = _host_call(ts, freqs, fdots, 64, 2, flags=0)
    b_s, b_p = _host_call(ts, freqs, fdots, 64, 2, flags=1)
    assert np.array_equal(a_p, b_p)
    np.testing.assert_allclose(a_s, b_s, rtol=1e-12)


def test_negative_fdot_negative_phase():
    """Negative phases (generic path is selected by the range check)."""
    ts = _events(8000, 500.0, seed=7)
    freqs = np.array([0.001, 0.01, 0.5])
    fdots = [-1e-3, -1e-5]
    ref_s, ref_p = C.fold_search(ts, freqs, fdots, 32, 2, want_profiles=True)
    got_s, got_p = _host_call(ts, freqs, fdots, 32, 2)
    assert np.array_equal(got_p.astype(np.int64), ref_p)
    _assert_stats(got_s, ref_s)


def test_long_baseline_large_phase():
    """T = 1e8 s, f ~ 10 Hz: |phase| ~ 1e9 cycles; bins must still be bit-exact."""
    rng = np.random.default_rng(11)
    ts = np.sort(rng.uniform(0, 1e8, 30000))
    freqs = 9.9 + 5e-9 * np.arange(96)
    ref_s, ref_p = C.fold_search(ts, freqs, [0.0, 1e-16], 128, 2, want_profiles=True)
    got_s, got_p = _host_call(ts, freqs, [0.0, 1e-16], 128, 2)
    assert np.array_equal(got_p.astype(np.int64), ref_p)
    _assert_stats(got_s, ref_s)


def test_event_tail_sizes():
    """Ragged event counts around the tile size (512) and odd counts (16-byte bulk-copy tail)."""
    freqs = 9.85 + 0.1 * np.arange(40) / 40
    base = _events(3000, 100.0, seed=13)
    for n in [1, 2, 3, 511, 512, 513, 1023, 1025, 2049]:
        ts = base[:n]
        for fdots in ([0.0], [1e-5]):
            ref_s, ref_p = C.fold_search(ts, freqs, fdots, 16, 1, want_profiles=True)
            got_s, got_p = _host_call(ts, freqs, fdots, 16, 1)
            assert np.array_equal(got_p.astype(np.int64), ref_p), (n, fdots)
            _assert_stats(got_s, ref_s)


def test_python_api_matches_oracle_search():
    from hendrics_b200 import epoch_folding_search, z_n_search

    ts = _events(20000, 1000.0, seed=1) + 8e7  # NuSTAR-like MET offset, removed by re-zeroing
    freqs = np.arange(9.85, 9.95, 0.0005)
    f0, s0 = O.z_n_search(ts, freqs, nharm=2, nbin=128)
    f1, s1 = z_n_search(ts, freqs, nharm=2, nbin=128)
    assert np.array_equal(f0, f1)
    np.testing.assert_allclose(s1, s0, rtol=RTOL)
    fd = np.array([-1e-6, 0, 2e-6])
    r0 = O.epoch_folding_search(ts, freqs, nbin=16, fdots=fd)
    r1 = epoch_folding_search(ts, freqs, nbin=16, fdots=fd)
    assert len(r1) == 3 and r1[2].shape == (3, freqs.size)
    np.testing.assert_allclose(r1[2], r0[2], rtol=RTOL)
    assert np.array_equal(r1[0], r0[0]) and np.array_equal(r1[1], r0[1])
    # same top-k candidate list
    assert np.array_equal(O.topk_flat(r1[2], 10)[0], O.topk_flat(r0[2], 10)[0])


def test_segments():
    from hendrics_b200 import z_n_search

    ts = _events(30000, 1000.0, seed=2)
    freqs = np.arange(9.85, 9.95, 0.002)
    f0, s0 = O.z_n_search(ts, freqs, nharm=2, nbin=32, segment_size=300.0)
    f1, s1 = z_n_search(ts, freqs, nharm=2, nbin=32, segment_size=300.0)
    np.testing.assert_allclose(s1, s0, rtol=RTOL)


def test_many_trials_fused_path():
    """Enough trials that the fused-epilogue (no scratch) path is taken, incl. a ragged last block."""
    ts = _events(6000, 500.0, seed=4)
    freqs = 5.0 + 10.0 * np.arange(40013) / 40013
    fdots = [0.0, 1e-6, 2e-6]
    ref = C.fold_search(ts, freqs, fdots, 16, 2)
    got, _ = _host_call(ts, freqs, fdots, 16, 2, want_profiles=False)
    _assert_stats(got, ref)
    ref = C.fold_search(ts, freqs[:20011], [0.0], 128, 0)
    got, _ = _host_call(ts, freqs[:20011], [0.0], 128, 0, want_profiles=False)
    _assert_stats(got, ref)


def test_microbench_runs():
    from hendrics_b200 import _cabi

    for kind in range(4):
        assert _cabi.microbench(kind, reps=2) > 1e9


@pytest.mark.parametrize("nbin", [128, 256, 512])
@pytest.mark.parametrize("fdots", [[0.0], [0.0, 3e-6]])
def test_packed_counters_equal_plain_and_oracle(nbin, fdots):
    """nbin >= 128 uses packed (two bins per word) counters; they must give the same bits as the
    plain layout and as the oracle, including across the periodic drains."""
    from hendrics_b200 import _cabi

    ts = _events(40000, 1000.0, seed=nbin)
    freqs = 9.85 + 0.1 * np.arange(200) / 200
    ref_s, ref_p = C.fold_search(ts, freqs, fdots, nbin, 2, want_profiles=True)
    lib = _cabi.lib()
    try:
        lib.hx_fold_set_packing(1)
        p_s, p_p = _host_call(ts, freqs, fdots, nbin, 2)
        lib.hx_fold_set_packing(0)
        u_s, u_p = _host_call(ts, freqs, fdots, nbin, 2)
    finally:
        lib.hx_fold_set_packing(-1)
    assert np.array_equal(p_p.astype(np.int64), ref_p)
    assert np.array_equal(u_p.astype(np.int64), ref_p)
    _assert_stats(p_s, ref_s)
    _assert_stats(u_s, ref_s)
    a_s, _ = _host_call(ts, freqs, fdots, nbin, 2, want_profiles=False)  # automatic choice
    _assert_stats(a_s, ref_s)


def test_all_events_in_one_bin_no_counter_overflow():
    """70000 photons in a single phase bin: more than a 16-bit (or packed 13-bit) field holds."""
    ts = np.arange(70000, dtype=np.float64)
    freqs = np.array([1.0, 1.0 + 1.0 / 128 / 70000])
    for nbin in (16, 128, 256):
        ref_s, ref_p = C.fold_search(ts, freqs, [0.0], nbin, 1, want_profiles=True)
        got_s, got_p = _host_call(ts, freqs, [0.0], nbin, 1)
        assert np.array_equal(got_p.astype(np.int64), ref_p)
        assert got_p[0, 0, 0] == 70000
        _assert_stats(got_s, ref_s)
    # upper half bin of a packed pair
    ts2 = ts + 0.75
    ref_s, ref_p = C.fold_search(ts2, freqs[:1], [0.0], 128, 1, want_profiles=True)
    got_s, got_p = _host_call(ts2, freqs[:1], [0.0], 128, 1)
    assert np.array_equal(got_p.astype(np.int64), ref_p) and got_p[0, 0, 96] == 70000


def test_exact_phase_mode_matches_rational_arithmetic():
    """Extension (not the parity path): HX_FOLD_EXACT_PHASE bins equal exact rational arithmetic on the
    fp64 inputs, while the reference-parity fp64 phase flips a few photons on a 1e8 s baseline."""
    from hendrics_b200 import fold_profiles

    rng = np.random.default_rng(7)
    ts = np.sort(rng.uniform(0, 1e8, 4000))
    nbin = 128
    for f, fd in ((700.123456789, 0.0), (700.123456789, 1.2345e-9), (9.9009900990099, -3e-12)):
        exact = np.bincount(O.exact_bins(ts, f, fd, nbin), minlength=nbin)
        _, pe = fold_profiles(ts, [f], [fd], nbin=nbin, nharm=2, exact_phase=True)
        assert np.array_equal(pe[0, 0], exact), (f, fd)
        _, pp = fold_profiles(ts, [f], [fd], nbin=nbin, nharm=2)
        ref_s, ref_p = C.fold_search(ts, [f], [fd], nbin, 2, want_profiles=True)
        assert np.array_equal(pp[0, 0], ref_p[0, 0])  # parity mode still follows the reference bit for bit
    # with |t f| ~ 7e10 cycles the two differ
    _, pp = fold_profiles(ts, [700.123456789], [1.2345e-9], nbin=nbin, nharm=2)
    _, pe = fold_profiles(ts, [700.123456789], [1.2345e-9], nbin=nbin, nharm=2, exact_phase=True)
    assert pp.sum() == pe.sum() == ts.size


def test_gpu_topk_in_find_peaks():
    from hendrics_b200 import base

    rng = np.random.default_rng(1)
    img = rng.chisquare(4, size=(64, 20000))
    img[3, 17] = img[40, 19999] = 500.0
    old = base.GPU_TOPK_MIN_SIZE
    try:
        base.GPU_TOPK_MIN_SIZE = 1000
        got = base.find_peaks_in_image(img, n=7)
    finally:
        base.GPU_TOPK_MIN_SIZE = old
    assert np.array_equal(got, O.find_peaks_in_image(img, n=7)) or np.array_equal(
        [tuple(g) for g in got], [np.unravel_index(i, img.shape) for i in O.topk_flat(img, 7)[0]])
    assert tuple(got[0]) == (3, 17) and tuple(got[1]) == (40, 19999)


def test_upstream_semantics_switches_follow_the_oracle():
    """A1/A2 (the two unverifiable readings of stingray's _folding_search) can be flipped in the product
    exactly like in the oracle."""
    from hendrics_b200 import search, z_n_search

    ts = _events(9000, 400.0, seed=23) + 50.0
    freqs = np.arange(9.85, 9.95, 0.001)
    saved = dict(search.UPSTREAM_SEMANTICS)
    try:
        for rz in (True, False):
            for dl in (True, False):
                search.UPSTREAM_SEMANTICS.update(rezero_to_first_event=rz, drop_last_event=dl)
                f1, s1 = z_n_search(ts, freqs, nharm=2, nbin=32)
                f0, s0 = O.z_n_search(ts, freqs, nharm=2, nbin=32, rezero_to_first_event=rz, drop_last_event=dl)
                np.testing.assert_allclose(s1, s0, rtol=RTOL, err_msg=f"rezero={rz} drop_last={dl}")
    finally:
        search.UPSTREAM_SEMANTICS.update(saved)


def test_trial_batching_of_the_scratch_path(monkeypatch):
    """Large trial grids on the scratch path (packed counters, profiles, event splits) are processed in
    batches bounded by a scratch cap; force a tiny cap so that several batches (and a ragged last one) run."""
    monkeypatch.setenv("HX_FOLD_SCRATCH_CAP_KB", "256")
    ts = _events(12000, 600.0, seed=31)
    freqs = 9.0 + 2.0 * np.arange(1500) / 1500
    fdots = [0.0, 1e-6, 2e-6]
    for nbin in (128, 16):
        ref_s, ref_p = C.fold_search(ts, freqs, fdots, nbin, 2, want_profiles=True)
        got_s, got_p = _host_call(ts, freqs, fdots, nbin, 2)
        assert np.array_equal(got_p.astype(np.int64), ref_p)
        _assert_stats(got_s, ref_s)
        got_s2, _ = _host_call(ts, freqs, fdots, nbin, 2, want_profiles=False)
        _assert_stats(got_s2, ref_s)


def test_randomized_configs_against_oracle():
    """Fuzz: random sizes, nbin (power of two or not), harmonics, fdot signs, time offsets and flags."""
    rng = np.random.default_rng(int(os.environ.get("HX_TEST_SEED", 2026)))  # other seeds: HX_TEST_SEED=n pytest -k randomized
    for it in range(40):
        n_ev = int(rng.integers(1, 6000))
        T = float(10 ** rng.uniform(0.5, 4))
        ts = np.sort(rng.uniform(0, T, n_ev)) + (rng.choice([0.0, 3.0, 1e3]))
        nbin = int(rng.choice([2, 3, 8, 16, 20, 32, 64, 100, 128, 256, 512, 777, 1024]))
        nharm = int(rng.choice([0, 1, 2, 5]))
        n_f = int(rng.integers(1, 150))
        freqs = rng.uniform(0.01, 50) + rng.uniform(1e-6, 1e-2) * np.arange(n_f)
        n_fd = int(rng.choice([1, 1, 2, 5]))
        fdots = [0.0] if n_fd == 1 and rng.random() < 0.5 else list(rng.uniform(-1, 1, n_fd) * 10 ** rng.uniform(-9, -4))
        flags = int(rng.choice([0, 0, 1, 2, 3]))
        ref_s, ref_p = C.fold_search(ts, freqs, fdots, nbin, nharm, want_profiles=True)
        want_prof = bool(rng.random() < 0.6)
        got_s, got_p = _host_call(ts, freqs, fdots, nbin, nharm, flags=flags, want_profiles=want_prof)
        msg = f"it={it} n_ev={n_ev} T={T:.3g} nbin={nbin} nharm={nharm} n_f={n_f} fdots={fdots} flags={flags}"
        if want_prof:
            assert np.array_equal(got_p.astype(np.int64), ref_p), msg
        good = np.isfinite(ref_s)
        assert np.array_equal(np.isfinite(got_s), good), msg
        np.testing.assert_allclose(got_s[good], ref_s[good], rtol=RTOL, atol=1e-9, err_msg=msg)


def test_error_paths():
    from hendrics_b200 import _cabi

    ts = _events(100, 10.0, seed=1)
    f = np.array([1.0])
    with pytest.raises(_cabi.HxError, match="bad sizes"):
        _host_call(ts, f, [0.0], 1, 0)
    with pytest.raises(_cabi.HxError, match="bad sizes"):
        _host_call(ts, f, [0.0], 4096, 0)
    with pytest.raises(_cabi.HxRangeError, match="exceeds the fp64 folding range"):
        _host_call(ts * 1e12, np.array([1e6]), [0.0], 16, 0)
    with pytest.raises(_cabi.HxError, match="non-finite"):
        _host_call(np.array([0.0, np.nan, 1.0]), f, [0.0], 16, 0)
    # the library stays usable after an error
    s, _ = _host_call(ts, f, [0.0], 16, 0)
    assert np.isfinite(s).all()


def test_concurrent_calls_from_host_threads():
    """The library serialises its entry points: searches issued from several host threads give the single-thread
    results (the reference makes no re-entrancy promise; this one does)."""
    import threading

    from hendrics_b200 import epoch_folding_search, z_n_search

    ts = _events(30000, 500.0, 13)
    freqs = np.arange(9.8, 10.0, 0.0005)
    want_z = z_n_search(ts, freqs, nharm=2, nbin=64)[1]
    want_e = epoch_folding_search(ts, freqs, nbin=16, fdots=[0.0, 1e-6])[2]
    out, errs = {}, []

    def work(k):
        try:
            for _ in range(5):
                if k % 2:
                    out[k] = z_n_search(ts, freqs, nharm=2, nbin=64)[1]
                else:
                    out[k] = epoch_folding_search(ts, freqs, nbin=16, fdots=[0.0, 1e-6])[2]
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for k in range(4):
        assert np.array_equal(out[k], want_z if k % 2 else want_e)


@pytest.mark.parametrize("nbin,fdots,f_lo", [(20, [0.0], 9.8), (100, [0.0, -3e-6, 2e-6], 9.8), (1000, [0.0], 0.7), (24, [-1e-7], 150.0)])
def test_generic_short_path_equals_reference_sequence_and_oracle(nbin, fdots, f_lo):
    """Non power-of-two nbin: the short exact path (phase >= 512 cycles: floor(x*nbin) - floor(x)*nbin from a rounded-down
    product) gives the reference's bins -- compared with the operation-by-operation replay (hx_fold_set_generic_fast(0)) and
    with the oracle; events below and above the 512-cycle threshold, unsorted events (mixed stages), negative fdot."""
    from hendrics_b200 import _cabi

    rng = np.random.default_rng(nbin)
    ts = np.sort(rng.uniform(0.0, 3000.0, 40000))
    ts[:50] = rng.uniform(0.0, 512.5 / f_lo, 50)  # around the threshold t = 512 / f
    ts[5000:5010] = 512.0 / (f_lo + 0.01 * np.arange(10))  # phases of ~512 cycles exactly
    rng.shuffle(ts[10000:20000])  # unsorted part: stages mixing early and late events
    freqs = f_lo + 0.2 * np.arange(70) / 70
    lib = _cabi.lib()
    try:
        lib.hx_fold_set_generic_fast(1)
        s_fast, p_fast = _host_call(ts, freqs, fdots, nbin, 2)
        lib.hx_fold_set_generic_fast(0)
        s_ref, p_ref = _host_call(ts, freqs, fdots, nbin, 2)
    finally:
        lib.hx_fold_set_generic_fast(1)
    assert np.array_equal(p_fast, p_ref)
    assert np.array_equal(s_fast, s_ref)
    o_s, o_p = C.fold_search(ts, freqs, fdots, nbin, 2, want_profiles=True)
    assert np.array_equal(p_fast.astype(np.int64), o_p)
    _assert_stats(s_fast, o_s)
