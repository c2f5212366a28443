"""Orbital-demodulation fold (hendrics/efsearch.py:263-310) against the oracle's restatement of the T0 loop.
sin() is CUDA's on the device (<= 2 ulp) and numpy's SIMD kernel in the oracle (<= 1 ulp, CPU dependent): the two may
differ in the last bit of sin, but X*sin is ~1e4 times smaller than t, so the rounded difference t - X*sin changes in
its last bit for ~1e-3 of the events only, and such a change crosses a bin edge with probability ~ulp(t)*f*nbin ~ 1e-10.
Measured (profiles/r02_orbital_diff.txt): all 4 x 71 profiles of these cases are IDENTICAL to the oracle's, with
numpy's sin and with a correctly rounded sin alike.  Bar: profile counts bit-exact, statistics within 1e-9, like K1."""
import os

import numpy as np
import pytest

from oracle import hx_oracle as O

pytestmark = pytest.mark.gpu


def _binary_events(n, T, f0, X, Porb, T0, seed):
    """pulsar in a circular orbit: emission times pulsed at f0, arrival = emission + X sin(2 pi (t - T0) / Porb)."""
    te = O.simulate_events(n, T, f0=f0, amp=0.8, seed=seed)
    return np.sort(te + X * np.sin(2 * np.pi * (te - T0) / Porb))


@pytest.mark.parametrize("nbin,nharm,fdot", [(16, 0, 0.0), (64, 2, 0.0), (32, 1, 2e-9), (20, 0, -1e-9)])
def test_orbital_stats_vs_oracle(nbin, nharm, fdot):
    from hendrics_b200.search import orbital_fold_stats

    f0, X, Porb, T0 = 1.2345, 3.0, 5000.0, 1234.0
    ts = _binary_events(40001, 20000.0, f0, X, Porb, T0, seed=nbin)
    T0s = np.concatenate([[T0], np.random.default_rng(1).uniform(0, Porb, 70)])
    stats, prof = orbital_fold_stats(ts, f0, fdot, X, Porb, T0s, nbin=nbin, nharm=nharm, return_profiles=True)
    fun = O.epoch_folding_search if nharm == 0 else O.z_n_search
    kw = dict(nbin=nbin) if nharm == 0 else dict(nbin=nbin, nharm=nharm)
    ref_s, ref_p = O.orbital_stats(ts, f0, fdot, X, Porb, T0s, fun=fun, return_profiles=True, **kw)
    assert prof.shape == ref_p.shape and np.array_equal(prof.sum(axis=1), ref_p.sum(axis=1))
    assert np.array_equal(prof, ref_p), np.abs(prof - ref_p).sum(axis=1)
    np.testing.assert_allclose(stats, ref_s, rtol=1e-9)
    # demodulating at the true epoch concentrates the pulsed power; a typical wrong epoch smears it
    assert stats[0] > 2 * np.median(stats)


def test_orbital_event_splits_and_c_abi():
    from hendrics_b200 import _cabi

    f0, X, Porb = 2.5, 1.5, 3000.0
    ts = _binary_events(300001, 30000.0, f0, X, Porb, 500.0, seed=3)
    T0s = _cabi.as_f64(np.linspace(0, Porb, 5))
    stats = np.empty(5)
    prof = np.empty((5, 16), dtype=np.uint32)
    rc = _cabi.lib().hx_orbital_search(_cabi.dptr(ts), ts.size, f0, 0.0, X, Porb, _cabi.dptr(T0s), 5, 16, 0,
                                       _cabi.dptr(stats), prof.ctypes.data_as(_cabi._c_u32p), 0)
    _cabi.check(rc)
    ref_s, ref_p = O.orbital_stats(ts, f0, 0.0, X, Porb, T0s, fun=O.epoch_folding_search, return_profiles=True, nbin=16)
    assert np.array_equal(prof.astype(np.int64), ref_p)
    np.testing.assert_allclose(stats, ref_s, rtol=1e-9)
    assert _cabi.lib().hx_orbital_search(_cabi.dptr(ts), ts.size, f0, 0.0, X, -1.0, _cabi.dptr(T0s), 5, 16, 0,
                                         _cabi.dptr(stats), None, 0) == _cabi.HX_ERR_ARG


def test_folding_orbital_search_csv(tmp_path, monkeypatch):
    """hendrics/tests/test_efsearch.py:607-628: every row of the parameter table ends up done, with its extrema."""
    import pandas as pd

    from hendrics_b200 import epoch_folding_search, z_n_search
    from hendrics_b200.efsearch import decide_binary_parameters, folding_orbital_search
    from hendrics_b200.io import EventList

    monkeypatch.chdir(tmp_path)
    ts = O.simulate_events(10000, 25.25, seed=42) + 8e7
    ev = EventList(time=ts, gti=np.array([[8e7, 8e7 + 25.25]]))
    np.random.seed(5)
    csv_file = decide_binary_parameters(137430, [0.03, 0.035], [2.0 * 86400, 2.5 * 86400], [0.0, 1.0],
                                        fdot_range=[0, 5e-10], reset=False, NMAX=10)
    table = pd.read_csv(csv_file)
    assert len(table) == 10
    folding_orbital_search(ev, csv_file, chunksize=10, outfile="out.csv")
    out = pd.read_csv("out.csv")
    assert len(out) == 10 and np.all(out["done"])
    assert np.all(out["max_stat"] >= out["min_stat"]) and np.all(out["best_T0"] >= 0)
    # the batched launch equals calling the search T0 by T0 (the reference's structure) on the same random T0s
    np.random.seed(7)
    folding_orbital_search(ev, csv_file, chunksize=10, outfile="a.csv", fun=z_n_search, nharm=2, nbin=32)
    np.random.seed(7)
    folding_orbital_search(ev, csv_file, chunksize=10, outfile="b.csv",
                           fun=lambda t, f, fdots=0, **kw: z_n_search(t, f, fdots=fdots, nharm=2, nbin=32))
    a, b = pd.read_csv("a.csv"), pd.read_csv("b.csv")
    np.testing.assert_allclose(a["max_stat"], b["max_stat"], rtol=1e-6)
    np.testing.assert_allclose(a["min_stat"], b["min_stat"], rtol=1e-6)
