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
    i = np.unravel_index(np.argmax(ef.stat), ef.stat.shape)
    assert np.isclose(ef.freq[i], PULSE_F, atol=2 / 25.25)


def test_cli_errors(evfile):
    from hendrics_b200.efsearch import main_efsearch, main_zsearch

    with pytest.raises(ValueError, match="The fast option is only available for z searches"):
        main_efsearch([evfile, "-f", "9.85", "-F", "9.95", "--fast"])
    with pytest.warns(UserWarning, match="The number of bins is too small"):
        main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "-n", "2", "-N", "2", "--fast"])
    with pytest.raises(NotImplementedError):
        main_zsearch([evfile, "-f", "9.85", "-F", "9.95", "--ffa"])
