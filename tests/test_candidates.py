"""CPU: candidate post-processing (hendrics/efsearch.py:1246-1565), the stand-alone driver's host logic and the rebinding
shim.  No kernel is called: periodograms are synthetic arrays."""
import os
import types

import numpy as np
import pytest

from hendrics_b200.candidates import (
    CAND_COLUMNS,
    CandTable,
    _analyze_qffa_results,
    analyze_qffa_results,
    get_boundaries_from_level,
    get_xy_boundaries_from_level,
    trace_level_curves,
)
from hendrics_b200.io import EFPeriodogram, save_folding


def test_reference_doctest_known_answers():
    """The two known answers of the reference's doctests (hendrics/efsearch.py:1276-1283, 1311-1316): the half-maximum
    half-width of sinc^2 is 0.44."""
    x = np.linspace(-10, 10, 1000)
    y = np.sinc(x) ** 2 + np.sinc(x - 5) ** 2
    assert np.allclose(np.abs(get_boundaries_from_level(x, y, 0.5, 0)), 0.44, atol=0.1)
    X, Y = np.meshgrid(x, x)
    Z = np.sinc(np.sqrt(X**2 + Y**2)) ** 2 + np.sinc(np.sqrt((X - 5) ** 2 + Y**2)) ** 2
    vals = get_xy_boundaries_from_level(X, Y, Z, 0.5, 0, 0)
    assert np.allclose(np.abs(vals), 0.44, atol=0.1)
    # the neighbouring peak has its own contour, selected through (x0, y0)
    v5 = get_xy_boundaries_from_level(X, Y, Z, 0.5, 5, 0)
    assert np.allclose(v5, [5 - 0.44, 5 + 0.44, -0.44, 0.44], atol=0.1)
    # much sharper than the 0.1 of the doctest: linear interpolation on a 0.02 grid (the x extent feels the tail of
    # the neighbouring peak, the y extent is the pure sinc^2 half width)
    assert np.allclose(np.abs(vals[2:]), 0.4429, atol=2e-4) and np.allclose(np.abs(vals[:2]), 0.445, atol=2e-3)


def test_boundaries_1d_matches_the_reference_loop_literally():
    """get_boundaries_from_level against the reference's two while loops, transcribed here as the checker."""
    def loops(x, y, level, x0):
        max_idx = np.argmin(np.abs(x - x0))
        idx, lo, hi = max_idx, x0, x0
        while idx > 0 and y[idx] > level:
            lo = x[idx]
            idx -= 1
        idx = max_idx
        while idx < y.size and y[idx] > level:
            hi = x[idx]
            idx += 1
        return lo, hi

    rng = np.random.default_rng(0)
    for trial in range(200):
        n = rng.integers(1, 40)
        x = np.sort(rng.uniform(-5, 5, n))
        y = rng.uniform(0, 1, n)
        level = rng.uniform(0, 1)
        x0 = rng.uniform(-6, 6)
        assert get_boundaries_from_level(x, y, level, x0) == loops(x, y, level, x0), (trial, n)


def test_level_curves_of_a_circle_and_a_curvilinear_grid():
    u = np.linspace(-2, 2, 401)
    X, Y = np.meshgrid(u, u)
    curves, segs = trace_level_curves(X, Y, -(X**2 + Y**2), -1.0)
    assert len(curves) == 1
    r = np.hypot(curves[0][:, 0], curves[0][:, 1])
    assert np.all(np.abs(r - 1) < 1e-4)
    # closed curve: every crossing point is shared by exactly two segments
    assert segs[0].shape[0] == curves[0].shape[0]
    # two separate peaks -> two curves; a sheared (non-rectilinear) coordinate grid moves the box accordingly
    Z = np.exp(-((X + 1) ** 2 + Y**2) * 8) + np.exp(-((X - 1) ** 2 + Y**2) * 8)
    curves, _ = trace_level_curves(X, Y, Z, 0.5)
    assert len(curves) == 2
    box = get_xy_boundaries_from_level(X + 0.5 * Y, Y, Z, 0.5, -1, 0)
    rr = np.sqrt(np.log(2) / 8)
    assert np.allclose(box[2:], [-rr, rr], atol=1e-3) and np.allclose(box[:2], [-1 - rr * np.hypot(1, 0.5), -1 + rr * np.hypot(1, 0.5)], atol=2e-3)
    # saddle cells are resolved by the cell-centre value: a checkerboard cell with a high centre joins the high corners
    z = np.array([[1.0, 0.0], [0.0, 1.0]])
    c_hi, _ = trace_level_curves(np.array([[0.0, 1.0], [0.0, 1.0]]), np.array([[0.0, 0.0], [1.0, 1.0]]), z, 0.4)
    c_lo, _ = trace_level_curves(np.array([[0.0, 1.0], [0.0, 1.0]]), np.array([[0.0, 0.0], [1.0, 1.0]]), z, 0.6)
    assert len(c_hi) == 2 and len(c_lo) == 2
    with pytest.raises(ValueError):
        get_xy_boundaries_from_level(X, Y, Z, 5.0, 0, 0)


def _plane(nf=64, nfd=21, peak=400.0, seed=0):
    """A synthetic Z^2_2 plane in the layout search_with_qffa returns (first axis frequency, second fdot)."""
    f = np.linspace(9.85, 9.95, nf)
    fd = np.linspace(-1e-7, 1e-7, nfd)
    F, FD = np.meshgrid(f, fd, indexing="ij")
    rng = np.random.default_rng(seed)
    stat = rng.chisquare(4, F.shape)
    stat += peak * np.sinc((F - 9.9) / 0.01) ** 2 * np.sinc((FD - 2e-8) / 5e-8) ** 2
    return F, FD, stat


def test_candidate_table_side_files_and_meta(tmp_path, capsys):
    F, FD, stat = _plane()
    ef = EFPeriodogram(F, stat, "Z2n", 16, 2, fdots=FD, M=1, oversample=8, ncounts=10000, filename="nonexistent_ev.p",
                       pepoch=56000.25, mjdref=56000)
    fname = str(tmp_path / "obs_Z22_9.85-9.95Hz_fast.p")
    save_folding(ef, fname)
    ef2, table = analyze_qffa_results(fname)
    out = capsys.readouterr().out
    assert "Best candidate(s):" in out and "±" in out
    assert table.colnames == CAND_COLUMNS and 1 <= len(table) <= 5
    i, j = np.unravel_index(np.argmax(stat), stat.shape)
    assert table["power"][0] == stat.max() and table["f_idx"][0] == i and table["fdot_idx"][0] == j
    assert table["f"][0] == F[i, j] and table["fdot"][0] == FD[i, j] and table["mjd"][0] == 56000.25
    # the error box encloses the peak and is smaller than the plane
    assert -0.01 < table["f_err_n"][0] < 0 < table["f_err_p"][0] < 0.01
    assert -5e-8 < table["fdot_err_n"][0] < 0 < table["fdot_err_p"][0] < 5e-8
    assert table["pulse_amp"][0] > 0 and np.isnan(table["pulse_amp_ul_0.9"][0])
    # meta consumed by plot_folding (hendrics/plot.py:422-437)
    assert table.meta["nbin"] == 16 and table.meta["ndof"] == 1 and table.meta["label"] == "$Z^2_{2}$"
    assert table.meta["detlev"] > 20 and table.meta["filename"] is None
    for row in table:  # plot.py:638-648 iterates rows and indexes the plane with them
        assert ef2.stat[row["f_idx"], row["fdot_idx"]] == row["power"]
    # files: the candidate list and, for each detected candidate, the two cuts through it
    root = fname[:-2]
    best = root + f"_cand_4_fdot{FD[i, j]}.csv"
    cut = CandTable.read(best)
    assert cut.colnames == ["freq", "stat"] and np.array_equal(cut["stat"], stat[i, :]) and np.array_equal(cut["freq"], F[i, :])
    cut2 = CandTable.read(root + f"_cand_4_f{F[i, j]}.dat")
    assert cut2.colnames == ["fdot", "stat"] and np.array_equal(cut2["stat"], stat[:, j])
    lst = CandTable.read(fname + "_best_cands.csv")
    assert lst.colnames == CAND_COLUMNS and len(lst) == len(table)
    assert lst["power"][0] == float(f"{stat.max():.2f}") and lst["f_idx"][0] == i  # column formats applied on write
    assert open(fname + "_best_cands.csv").readline().strip() == ",".join(CAND_COLUMNS)
    n_detected = int(np.sum(~np.isnan(np.asarray(table["pulse_amp"], dtype=float))))
    assert len([f for f in os.listdir(tmp_path) if "_cand_" in f]) == 2 * n_detected


def test_no_detection_gives_one_upper_limit_row_and_no_side_files(tmp_path, capsys):
    F, FD, stat = _plane(peak=0.0, seed=3)
    ef = EFPeriodogram(F, stat, "Z2n", 16, 2, fdots=FD, M=1, oversample=8, ncounts=10000, filename="x", pepoch=56000.0)
    table = _analyze_qffa_results(ef, fname=str(tmp_path / "noise.p"))
    out = capsys.readouterr().out
    assert "No pulsations found" in out and "(90%)" in out
    assert len(table) == 5 and np.all(np.isnan(np.asarray(table["pulse_amp"], dtype=float)))
    assert np.all(np.asarray(table["pulse_amp_ul_0.9"], dtype=float) > 0)
    assert os.listdir(tmp_path) == []


def test_one_dimensional_periodogram_and_epoch_folding_thresholds(capsys):
    f = np.linspace(9.85, 9.95, 500)
    stat = 31 + 300 * np.sinc((f - 9.9) / 0.004) ** 2
    ef = EFPeriodogram(f, stat, "EF", 32, 1, M=1, oversample=2, ncounts=5000, filename="x", pepoch=1.0)
    table = _analyze_qffa_results(ef)
    capsys.readouterr()
    assert table["f_idx"][0] == np.argmax(stat) and table["fdot"][0] == 0 and table["fdot_err_n"][0] == 0
    assert table["f_err_n"][0] < 0 < table["f_err_p"][0]
    assert table.meta["ndof"] == 32 and table.meta["nbin"] == 32 and table.meta["label"] == r"$\chi^2_32$ Stat"
    with pytest.raises(ValueError, match="Did not understand stats shape"):
        _analyze_qffa_results(EFPeriodogram(f, stat.reshape(1, 1, -1), "EF", 32, 1, ncounts=1, filename="x", pepoch=1.0))


def test_cand_table_behaves_like_the_table_it_replaces(tmp_path):
    t = CandTable(names=["a", "b", "s"], dtype=[int, float, str], formats={"b": ".2f"})
    t.add_row([1, 2.125, "x y"])
    t.add_row([2, np.nan, "z"])
    assert len(t) == 2 and t["a"].tolist() == [1, 2] and t[1]["s"] == "z"
    assert len(t[np.array([False, True])]) == 1 and t["a", "b"].colnames == ["a", "b"]
    t["c"] = ["p", "q"]
    p = tmp_path / "t.csv"
    t.write(p)
    assert p.read_text().splitlines() == ["a,b,s,c", "1,2.12,x y,p", "2,nan,z,q"]
    with pytest.raises(OSError):
        t.write(p)
    t.write(tmp_path / "t.dat", format="ascii")
    assert (tmp_path / "t.dat").read_text().splitlines()[1] == '1 2.12 "x y" p'
    back = CandTable.read(tmp_path / "t.dat")
    assert back["s"].tolist() == ["x y", "z"] and back["a"].tolist() == [1, 2]
    with pytest.raises(ValueError):
        t.add_row([1])


def test_output_names_and_strategy_selection():
    from hendrics_b200 import epoch_folding_search, z_n_search
    from hendrics_b200.driver import _common_parser, _statistic, _strategy_name, _validate, output_root

    def root(argv, func=z_n_search):
        ns = _common_parser(["dir/obs_ev.p"] + argv)
        _validate(ns, func)
        return output_root("dir/obs_ev.p", ns, _statistic(func, ns)), _strategy_name(ns), ns

    assert root(["-f", "9.85", "-F", "9.95"])[:2] == ("dir/obs_Z22_9.85-9.95Hz", "exact")
    assert root(["-f", "9.85", "-F", "9.95", "-N", "3", "--fast"])[:2] == ("dir/obs_Z23_9.85-9.95Hz_fast", "fast")
    assert root(["-f", "1", "-F", "2", "--ffa"])[:2] == ("dir/obs_Z22_1-2Hz_ffa", "ffa")
    assert root(["-f", "1", "-F", "2", "--emin", "3"], epoch_folding_search)[0] == "dir/obs_EF_3-**keV_1-2Hz"
    assert root(["-f", "1", "-F", "2", "--emax", "79", "--mean-fdot", " -1.5e-10", "--fast"])[0] == \
        "dir/obs_Z22_**-79keV_1-2Hz_fast_fd-1.5e-10s-2"
    # negative numbers are accepted as values (hendrics/base.py check_negative_numbers_in_args)
    ns = root(["-f", "1", "-F", "2", "--fdotmin", "-1e-9", "--fdotmax", "1e-9"])[2]
    assert ns.fdotmin == -1e-9 and ns.segment_size == 1e32 and ns.nbin == 128 and ns.dynstep == 128 and ns.conflevel == 99
    # --fit-candidates implies --find-candidates unless a frequency is forced
    assert root(["-f", "1", "-F", "2", "--fit-candidates"])[2].find_candidates is True
    assert root(["-f", "1", "-F", "2", "--fit-candidates", "--fit-frequency", "1.5"])[2].find_candidates is False
    with pytest.raises(ValueError, match="The fast option is only available for z searches"):
        root(["-f", "1", "-F", "2", "--fast"], epoch_folding_search)
    with pytest.raises(NotImplementedError):
        root(["-f", "1", "-F", "2", "-p", "orbit.par"])


def test_trial_grid_rules():
    from hendrics_b200.efsearch import TrialGrid

    g = TrialGrid(1000.0, 9.85, 9.95)
    assert g.step == 1 / 2 / 1000.0 and g.fdotstep == 1 / 2 / 1000.0**2
    assert np.array_equal(g.freqs, np.arange(9.85, 9.95 + 1e-8 * g.step, g.step)) and g.fdots.tolist() == [0.0]
    assert TrialGrid(1000.0, 9.9, 9.9).freqs.tolist() == [9.9]  # fmin == fmax still holds one trial
    g = TrialGrid(1000.0, 1, 2, step=0.25, fdotmin=-1e-6, fdotmax=1e-6, fdotstep=5e-7)
    assert g.freqs.tolist() == [1, 1.25, 1.5, 1.75, 2.0] and len(g.fdots) == 5 and "5 frequencies and 5 fdots" in g.describe()


def test_rebinding_shim_patches_and_restores_a_module():
    import hendrics_b200.efsearch as E
    import hendrics_b200.rebind as rb
    from hendrics_b200 import epoch_folding_search, z_n_search

    sentinel = object()
    fake = types.ModuleType("fake_hendrics_efsearch")
    for name in rb.REBOUND:
        setattr(fake, name, sentinel)
    del fake.search_with_qffa_step  # an attribute the target lacks is added, and removed again on uninstall
    rb.install(fake)
    assert fake.z_n_search is z_n_search and fake.epoch_folding_search is epoch_folding_search
    assert fake.search_with_qffa is E.search_with_qffa and fake.search_with_ffa is E.search_with_ffa
    assert fake.search_with_qffa_step is E.search_with_qffa_step and rb.is_installed(fake)
    rb.install(fake)  # idempotent: the originals are not overwritten by our own functions
    rb.uninstall(fake)
    assert fake.z_n_search is sentinel and not hasattr(fake, "search_with_qffa_step") and not rb.is_installed(fake)
    try:
        import hendrics  # noqa: F401
    except Exception:
        with pytest.raises(ImportError):
            rb.install()


def test_orbital_extremes_follow_the_running_max_min():
    from hendrics_b200.orbital import _extremes

    T0 = np.array([10.0, 20.0, 30.0, 40.0])
    assert _extremes(T0, [1.0, 5.0, 5.0, np.nan]) == (5.0, 1.0, 20.0)  # first of equal maxima, NaN never wins
    assert _extremes(T0, [np.nan] * 4) == (0, 1e32, None)
    assert _extremes(T0, [0.0, -1.0, 0.0, 0.0]) == (0, -1.0, None)


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_boundaries_golden_from_the_real_reference():
    """tests/golden/boundaries.npz: 64 seeded cases run through the reference's own get_boundaries_from_level
    (hendrics/efsearch.py:1299-1332, executed by make_golden.py): identical limits."""
    g = np.load(os.path.join(GOLD, "boundaries.npz"))
    for x, y, (n, level, x0), ref in zip(g["x"], g["y"], g["params"], g["out"]):
        n = int(n)
        assert get_boundaries_from_level(x[:n], y[:n], level, x0) == tuple(ref)


def test_binary_parameter_table_golden_from_the_real_reference(tmp_path):
    """tests/golden/binary_parameters.npz: the reference's decide_binary_parameters (hendrics/efsearch.py:208-260) with
    numpy's global RNG seeded: the restructured generator draws the same numbers in the same order."""
    import pandas as pd

    from hendrics_b200.orbital import PARAMETER_COLUMNS, decide_binary_parameters

    np.random.seed(9)
    csv = decide_binary_parameters(137430, [0.03, 0.035], [2.0 * 86400, 2.5 * 86400], [0.0, 1.0], fdot_range=[0, 5e-10],
                                   NMAX=3, csv_file=str(tmp_path / "db.csv"), reset=True)
    df = pd.read_csv(csv)
    assert list(df.columns) == PARAMETER_COLUMNS and not df["done"].any()
    ref = np.load(os.path.join(GOLD, "binary_parameters.npz"))["table"]
    assert np.array_equal(df[["freq", "fdot", "X", "Porb"]].to_numpy(dtype=np.float64), ref)


@pytest.mark.skipif(not os.path.isdir("/root/reference/hendrics"), reason="the reference tree is only in the build container")
def test_the_reference_s_own_folding_search_runs_on_the_rebound_functions(monkeypatch):
    """INTEGRATION.md section 3, end to end on the host side: the REAL hendrics.efsearch (imported from /root/reference with
    its uninstallable third-party imports stubbed, as tests/golden/make_golden.py does), patched by rebind.install(),
    calls this package's z_n_search / epoch_folding_search with its own keywords and unpacks their return tuples.  There
    is no GPU here, so the one kernel-calling function is replaced by the oracle for the duration of the test: what is
    checked is the glue -- signatures, keyword names, tuple shapes -- not the kernels."""
    import importlib
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden

    finder = make_golden.StubFinder()
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, "/root/reference")
    try:
        hef = importlib.import_module("hendrics.efsearch")
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove("/root/reference")
    import hendrics_b200.rebind as rb
    import hendrics_b200.search as S
    from hendrics_b200.io import EventList
    from oracle import hx_oracle as O

    def oracle_core(nharm, times, frequencies, nbin=128, segment_size=np.inf, fdots=0, **_ignored):
        return O.folding_search_core(times, frequencies, fdots=fdots, nbin=nbin, nharm=nharm, segment_size=segment_size)

    monkeypatch.setattr(S, "_folding_search", oracle_core)
    rb.install(hef)
    try:
        ts = O.simulate_events(3000, 100.0, seed=4) + 8e7
        ev = EventList(time=ts, gti=np.array([[8e7, 8e7 + 100.0]]))
        got = hef.folding_search(ev, 9.85, 9.95, func=hef.z_n_search, oversample=2, nbin=32, nharm=2)
        ref = O.folding_search(ts, 8e7, 9.85, 9.95, func=O.z_n_search, oversample=2, nbin=32, nharm=2)
        assert len(got) == 4 and np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
        assert got[2] == ref[2] and got[3] == ref[3]
        got2 = hef.folding_search(ev, 9.85, 9.95, func=hef.epoch_folding_search, oversample=2, nbin=16, fdotmin=-1e-3,
                                  fdotmax=1e-3)
        ref2 = O.folding_search(ts, 8e7, 9.85, 9.95, func=O.epoch_folding_search, oversample=2, nbin=16, fdotmin=-1e-3,
                                fdotmax=1e-3)
        assert len(got2) == 6 and np.array_equal(got2[2], ref2[2]) and np.array_equal(got2[1], ref2[1])
    finally:
        rb.uninstall(hef)
        from unittest.mock import MagicMock

        # leave no stub (nor the stub-built hendrics) behind for later importorskip()s
        for name in [m for m, mod in list(sys.modules.items())
                     if m.split(".")[0] == "hendrics" or isinstance(mod, MagicMock)]:
            del sys.modules[name]
