"""K1win (`hx_fold_windows_dev`): all windows of `dyn_folding_search` (hendrics/efsearch.py:1196-1243) / all segments
of `--segment-size` (stingray `_folding_search`) in one launch, against the oracle window by window
and against the one-K1-launch-per-window path."""
import numpy as np
import pytest

from oracle import hx_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _oracle_window(sel, freqs, nharm, nbin):
    """What dyn_folding_search stores for one window: func(times_filt, trial_freqs)[1], zeros when func raises."""
    try:
        if nharm:
            return O.z_n_search(sel, freqs, nharm=nharm, nbin=nbin)[1]
        return O.epoch_folding_search(sel, freqs, nbin=nbin)[1]
    except Exception:
        return np.zeros_like(freqs)


@pytest.mark.parametrize("nharm,nbin", [(2, 128), (0, 20), (3, 16)])
def test_independent_windows_match_the_oracle_window_by_window(nharm, nbin):
    from hendrics_b200.search import windowed_searches

    ts = O.simulate_events(30000, 400.0, seed=21) + 8e7
    # a hole: window 3 is empty; window 4 has events in its first 10 % plus ONE event near its end: its single
    # segment is as long as that last event is late, the half-open cut drops it, and what is left spans < 0.2 of the
    # segment, so upstream skips it (0/0 = NaN); window 5 has events only in its first 10 % (a short but valid search)
    ts = ts[~((ts >= 8e7 + 150.0) & (ts < 8e7 + 200.0))]
    ts = ts[~((ts >= 8e7 + 205.0) & (ts < 8e7 + 250.0))]
    ts = ts[~((ts >= 8e7 + 255.0) & (ts < 8e7 + 300.0))]
    ts = np.sort(np.append(ts, 8e7 + 249.0))
    ts = np.append(ts, ts[-1])  # a duplicated last event: both copies are dropped by the half-open segment (A2)
    starts = 8e7 + 50.0 * np.arange(8)
    stops = starts + 50.0
    freqs = np.arange(9.85, 9.95, 0.002)
    stats, status = windowed_searches(nharm, ts, freqs, starts, stops, nbin=nbin)
    assert stats.shape == (8, 1, freqs.size)
    assert status[3] == 2 and np.all(stats[3] == 0)
    assert status[4] == 0 and np.all(np.isnan(stats[4]))
    assert status[5] == 1 and np.all(np.isfinite(stats[5]))
    for w in range(8):
        sel = ts[(ts >= starts[w]) & (ts < stops[w])]
        with np.errstate(all="ignore"):
            ref = _oracle_window(sel, freqs, nharm, nbin)
        np.testing.assert_allclose(stats[w, 0], ref, rtol=RTOL, equal_nan=True, err_msg=f"window {w}")


def test_dyn_folding_search_uses_one_launch_and_equals_the_per_window_path(monkeypatch):
    from hendrics_b200 import _cabi, z_n_search
    from hendrics_b200 import search as S
    from hendrics_b200.efsearch import dyn_folding_search
    from hendrics_b200.io import EventList

    ts = O.simulate_events(200_000, 12800.0, seed=22) + 8e7
    ev = EventList(time=ts, gti=np.array([[8e7, 8e7 + 12800.0]]))
    n0 = _cabi.launch_count()
    t1, f1, s1 = dyn_folding_search(ev, 9.85, 9.95, func=z_n_search, oversample=2, time_step=128, nharm=2)
    launches = _cabi.launch_count() - n0
    assert s1.shape == (100, f1.size) and launches <= 8, launches  # plan + fold (+ range checks), not 100 x K1
    monkeypatch.setattr(S, "WINDOWED_MAX_UNITS_PER_WINDOW", 0.0)  # force the reference-like loop of K1 launches
    t2, f2, s2 = dyn_folding_search(ev, 9.85, 9.95, func=z_n_search, oversample=2, time_step=128, nharm=2)
    assert np.array_equal(t1, t2) and np.array_equal(f1, f2)
    np.testing.assert_allclose(s1, s2, rtol=1e-12)
    sel = ts[(ts >= 8e7 + 128.0 * 7) & (ts < 8e7 + 128.0 * 8)]
    np.testing.assert_allclose(s1[7], O.z_n_search(sel, f1, nharm=2, nbin=128)[1], rtol=RTOL)


@pytest.mark.parametrize("nbin,fdots", [(32, 0), (20, [0.0, 2e-7, -1e-7])])
def test_segment_size_sum_matches_oracle_and_loop(monkeypatch, nbin, fdots):
    """`--segment-size`: many short segments summed in order and divided by the number of folded segments."""
    from hendrics_b200 import _cabi, epoch_folding_search, z_n_search
    from hendrics_b200 import search as S

    ts = O.simulate_events(60000, 3000.0, seed=23) + 8e7
    ts = ts[~((ts >= 8e7 + 1000.0) & (ts < 8e7 + 1100.0))]  # segments without events are skipped
    freqs = np.arange(9.85, 9.95, 0.001)
    n0 = _cabi.launch_count()
    got = z_n_search(ts, freqs, nharm=2, nbin=nbin, segment_size=37.0, fdots=fdots)
    assert _cabi.launch_count() - n0 <= 8
    ref = O.z_n_search(ts, freqs, nharm=2, nbin=nbin, segment_size=37.0, fdots=fdots)
    np.testing.assert_allclose(got[-1], ref[-1], rtol=RTOL)
    got_e = epoch_folding_search(ts, freqs, nbin=nbin, segment_size=37.0, fdots=fdots)
    ref_e = O.epoch_folding_search(ts, freqs, nbin=nbin, segment_size=37.0, fdots=fdots)
    np.testing.assert_allclose(got_e[-1], ref_e[-1], rtol=RTOL)
    monkeypatch.setattr(S, "WINDOWED_MAX_UNITS_PER_WINDOW", 0.0)
    loop = z_n_search(ts, freqs, nharm=2, nbin=nbin, segment_size=37.0, fdots=fdots)
    np.testing.assert_allclose(got[-1], loop[-1], rtol=1e-12)
