"""End to end through the reference's CLI surface (behavioural checks of hendrics/tests/test_efsearch.py)."""
import os

import numpy as np
import pytest

from oracle import hx_oracle as O

pytestmark = pytest.mark.gpu

PULSE_F = 1 / 0.101


@pytest.fixture(scope="module")
def evfile(tmp_path_factory):
    from hendrics_b200.io import EventList, save_events

    d = tmp_path_factory.mktemp("cli")
    ts = O.simulate_events(10000, 25.25, seed=42) + 8e7
    ev = EventList(time=ts, gti=np.array([[8e7, 8e7 + 25.25]]), mjdref=56000,
                   energy=np.random.default_rng(0).uniform(3, 79, ts.size))
    fname = str(d / "events_ev.p")
    save_events(ev, fname)
    return fname


def test_zsearch_cli(evfile):
    from hendrics_b200.efsearch import main_zsearch
    from hendrics_b200.io import load_folding

    out = main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "-n", "64", "-N", "2", "--fit-candidates"])
    assert len(out) == 1 and out[0].endswith("events_Z22_9.85-9.95Hz.p") and os.path.exists(out[0])
    ef = load_folding(out[0])
    assert ef.kind == "Z2n" and ef.N == 2 and ef.nbin == 64 and ef.ncounts == 10000
    assert np.isclose(ef.peaks[0], PULSE_F, atol=1 / 25.25)
    assert len(ef.best_fits) >= 1 and abs(ef.best_fits[0]["mean"] - PULSE_F) < 1 / 25.25
    # identical to the oracle's folding_search on the same inputs
    ts = O.simulate_events(10000, 25.25, seed=42) + 8e7
    rf, rs, rstep, rlen = O.folding_search(ts, 8e7, 9.85, 9.95, func=O.z_n_search, oversample=2, nbin=64, nharm=2)
    assert np.array_equal(ef.freq, rf)
    np.testing.assert_allclose(ef.stat, rs, rtol=1e-9)


def test_efsearch_cli_fdot_and_energy(evfile):
    from hendrics_b200.efsearch import main_efsearch
    from hendrics_b200.io import load_folding

    out = main_efsearch([evfile, "-f", "9.85", "-F", "9.95", "-n", "32", "--emin", "3", "--emax", "79",
                         "--fdotmin", " -1e-3", "--fdotmax", "1e-3"])
    assert out[0].endswith("events_EF_3-79keV_9.85-9.95Hz.p")
    ef = load_folding(out[0])
    assert ef.kind == "EF" and len(ef.fdots) > 1 and ef.stat.shape == ef.freq.shape
    i = np.unravel_index(np.argmax(ef.stat), ef.stat.shape)
    assert np.isclose(ef.freq[i], PULSE_F, atol=1 / 25.25)


def test_zsearch_fast_cli(evfile):
    from hendrics_b200.efsearch import main_zsearch
    from hendrics_b200.io import load_folding

    np.random.seed(1)
    out = main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "-n", "64", "--fast"])
    assert out[0].endswith("events_Z22_9.85-9.95Hz_fast.p")
    ef = load_folding(out[0])
    assert len(ef.fdots) > 1
    assert os.path.exists(out[0] + "_best_cands.csv")
    # f1 outputs (hendrics/efsearch.py:1516-1533): the cuts through every detected candidate, read by HENplot
    from hendrics_b200.candidates import CandTable

    cands = CandTable.read(out[0] + "_best_cands.csv")
    root = out[0][:-2]
    side = sorted(f for f in os.listdir(os.path.dirname(out[0])) if os.path.basename(root) + "_cand_" in f)
    n_det = int(np.sum(~np.isnan(np.asarray(cands["pulse_amp"], dtype=float))))
    assert n_det >= 1 and len(side) == 2 * n_det
    cut = CandTable.read(os.path.join(os.path.dirname(out[0]), [f for f in side if "_cand_4_fdot" in f][0]))
    assert cut.colnames == ["freq", "stat"] and np.array_equal(cut["stat"], ef.stat[cands["f_idx"][0], :])
    i = np.unravel_index(np.argmax(ef.stat), ef.stat.shape)
    assert np.isclose(ef.freq[i], PULSE_F, atol=2 / 25.25)


def test_cli_errors(evfile):
    from hendrics_b200.efsearch import main_efsearch, main_zsearch

    with pytest.raises(ValueError, match="The fast option is only available for z searches"):
        main_efsearch([evfile, "-f", "9.85", "-F", "9.95", "--fast"])
    with pytest.warns(UserWarning, match="The number of bins is too small"):
        main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "-n", "2", "-N", "2", "--fast"])
    with pytest.raises(NotImplementedError):
        main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "--deorbit-par", "nonexistent.par"])


def test_transient_cli(evfile):
    from hendrics_b200.efsearch import main_zsearch

    np.random.seed(2)
    out = main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "-n", "16", "--transient"])
    assert out == []  # like the reference: `--transient` alone produces only the transient product
    npz = evfile.replace("_ev.p", "") + "_Z22_9.85-9.95Hz_transient.npz"
    assert os.path.exists(npz)
    d = np.load(npz)
    assert d["stats"].shape[0] == d["nave"].size and d["stats"].shape[2] == d["freqs"].size
    k = np.unravel_index(np.argmax(d["stats"][-1]), d["stats"][-1].shape)
    assert abs(d["freqs"][k[1]] - PULSE_F) < 2 / 25.25


def test_dyn_folding_search_and_candidates(evfile):
    from hendrics_b200 import z_n_search
    from hendrics_b200.efsearch import _analyze_qffa_results, dyn_folding_search, search_with_qffa
    from hendrics_b200.io import EFPeriodogram, load_events

    ev = load_events(evfile)
    times, freqs, stats = dyn_folding_search(ev, 9.85, 9.95, func=z_n_search, oversample=2, time_step=5.0, nharm=2, nbin=16)
    assert stats.shape == (times.size, freqs.size) and times.size == 5
    # each window equals a direct search on its events
    sel = ev.time[(ev.time >= 8e7) & (ev.time < 8e7 + 5.0)]
    f0, s0 = z_n_search(sel, freqs, nharm=2, nbin=16)
    np.testing.assert_allclose(stats[0], s0, rtol=1e-12)
    np.random.seed(3)
    r = search_with_qffa(ev.time, 9.85, 9.95, nbin=16, n=2, silent=True)
    ef = EFPeriodogram(r[0], r[2], "Z2n", 16, 2, fdots=r[1], M=1, oversample=8, ncounts=ev.time.size, filename=evfile,
                       pepoch=56000.0)
    rows = _analyze_qffa_results(ef)
    assert 1 <= len(rows) <= 5 and abs(rows[0]["f"] - PULSE_F) < 2 / 25.25
    assert rows[0]["power"] == r[2].max() and rows[0]["f_idx"] == np.unravel_index(np.argmax(r[2]), r[2].shape)[0]


def test_cli_runs_the_dynamical_search(evfile, tmp_path, monkeypatch):
    """efsearch.py:1912-1922: observations longer than --dynstep get the windowed search too; the reference draws
    Dyn.png from its arrays, here they are saved as Dyn.npz (hendrics/tests/test_efsearch.py:287-316 passes --dynstep 5)."""
    from hendrics_b200 import z_n_search
    from hendrics_b200.efsearch import dyn_folding_search, main_zsearch
    from hendrics_b200.io import load_events

    monkeypatch.chdir(tmp_path)
    out = main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "-n", "64", "--dynstep", "5"])
    assert os.path.exists(out[0]) and os.path.exists("Dyn.npz")
    d = np.load("Dyn.npz")
    ev = load_events(evfile)
    step = 1 / 2 / (ev.time[-1] - ev.gti[0, 0])
    t, f, s = dyn_folding_search(ev, 9.85, 9.95, step=step, func=z_n_search, oversample=2, time_step=5.0, nharm=2)
    assert d["stats"].shape == (5, f.size) and np.array_equal(d["times"], t) and np.array_equal(d["frequencies"], f)
    np.testing.assert_allclose(d["stats"], s, rtol=1e-12)
    # every window is the direct search of its events, from the oracle
    sel = ev.time[(ev.time >= 8e7 + 5.0) & (ev.time < 8e7 + 10.0)]
    np.testing.assert_allclose(d["stats"][1], O.z_n_search(sel, f, nharm=2, nbin=128)[1], rtol=1e-9)


def test_zsearch_ffa_cli(evfile):
    """hendrics/tests/test_efsearch.py:557-581 (`--ffa`): output name, experimental warning, pulse recovered"""
    from hendrics_b200.efsearch import main_zsearch
    from hendrics_b200.io import load_folding

    with pytest.warns(UserWarning, match="Fast Folding Algorithm functionality is experimental"):
        out = main_zsearch([evfile, "-f", "9.89", "-F", "9.92", "-n", "64", "--ffa", "--find-candidates"])
    assert len(out) == 1 and out[0].endswith("events_Z22_9.89-9.92Hz_ffa.p") and os.path.exists(out[0])
    ef = load_folding(out[0])
    assert ef.kind == "Z2n" and ef.nbin == 64
    best = ef.freq[np.nanargmax(ef.stat)]
    assert np.isclose(best, PULSE_F, atol=1 / 25.25)
    # same numbers as the oracle's restatement of search_with_ffa on the same events
    ts = O.simulate_events(10000, 25.25, seed=42) + 8e7
    fr, st, _, _ = O.search_with_ffa(ts, 9.89, 9.92, nbin=64, n=2)
    assert np.array_equal(ef.freq, fr)
    ok = ~np.isnan(st)
    assert np.array_equal(np.isnan(ef.stat), ~ok)
    np.testing.assert_allclose(ef.stat[ok], st[ok], rtol=1e-9)
