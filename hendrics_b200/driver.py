"""The stand-alone ``HENzsearch`` / ``HENefsearch`` driver of this package.

The PRIMARY integration is the rebinding shim (``hendrics_b200.rebind``): with HENDRICS
installed, its own ``_common_main`` (hendrics/efsearch.py:1771-1983) runs unchanged on top of the B200 searches.  This
module is for machines without HENDRICS/stingray/astropy: it accepts the same command line (flag table of
hendrics/efsearch.py:1568-1768), writes files with the same names and the same ``EFPeriodogram`` content, and raises
the same errors, but it is organised as a small pipeline instead of one long function:

    parse -> validate -> per file: load -> (energy filter) -> (transient scan) -> strategy[exact | fast | ffa]
          -> (dynamical pass) -> periodogram -> candidates -> (peak fits) -> save -> (--fast: candidate table)

Figures (the reference's ``Dyn.png`` and transient GIF need matplotlib / imageio) are replaced by ``.npz`` files
holding the arrays they are drawn from.
"""

from __future__ import annotations

import argparse
import copy
import logging
import warnings
from typing import Callable, NamedTuple, Optional

import numpy as np

from .base import HENDRICS_STAR_VALUE, check_negative_numbers_in_args, hen_root, pf_upper_limit
from .io import HEN_FILE_EXTENSION, EFPeriodogram, get_file_type, save_folding

log = logging.getLogger("hendrics_b200")

# ---------------------------------------------------------------------------------------------
# command line: one row per flag -- (flags, argparse keywords).  Same flags, destinations, types and defaults as
# hendrics/efsearch.py:1568-1768.
# ---------------------------------------------------------------------------------------------
_FLAG = dict(default=False, action="store_true")
_CLI_TABLE = (
    (("files",), dict(nargs="+", help="event lists or light curves to search")),
    (("-f", "--fmin"), dict(type=float, required=True, help="lowest trial frequency (Hz)")),
    (("-F", "--fmax"), dict(type=float, required=True, help="highest trial frequency (Hz)")),
    (("--emin",), dict(type=float, default=None, help="lower energy (or PI) bound of the events used")),
    (("--emax",), dict(type=float, default=None, help="upper energy (or PI) bound of the events used")),
    (("--mean-fdot",), dict(type=float, default=0, help="central fdot of a --fast search")),
    (("--mean-fddot",), dict(type=float, default=0, help="central fddot of a --fast search")),
    (("--fdotmin",), dict(type=float, default=None, help="lowest trial fdot")),
    (("--fdotmax",), dict(type=float, default=None, help="highest trial fdot")),
    (("--dynstep",), dict(type=int, default=128, help="window length (s) of the dynamical search")),
    (("--npfact",), dict(type=int, default=2, help="size of the --fast parameter space around each centre")),
    (("--n-transient-intervals",), dict(type=int, default=None, help="time slices of the --transient scan")),
    (("-n", "--nbin"), dict(type=int, default=128, help="phase bins of the folded profile")),
    (("--segment-size",), dict(type=float, default=1e32,
                               help="fold segments of this length and average their statistics (default: whole data)")),
    (("--step",), dict(type=float, default=None, help="frequency step (default 1 / oversample / length)")),
    (("--oversample",), dict(type=float, default=None, help="frequency oversampling with respect to 1 / length")),
    (("--fast",), dict(help="quasi-fast folding (Z search only): pre-binned profiles shifted over an (f, fdot) "
                            "neighbourhood; ignores --expocorr, --fdotmin/max, --segment-size and --step", **_FLAG)),
    (("--ffa",), dict(help="the Fast Folding Algorithm on a binned light curve (slow pulsars; no fdot)", **_FLAG)),
    (("--transient",), dict(help="scan for transient pulsations; arrays go to <root>_transient.npz", **_FLAG)),
    (("--expocorr",), dict(help="exposure-corrected profiles (GTI aware; the weighted kernel)", **_FLAG)),
    (("--find-candidates",), dict(help="list the peaks above the --conflevel threshold", **_FLAG)),
    (("--conflevel",), dict(type=float, default=99, help="confidence level (percent) of the threshold")),
    (("--fit-candidates",), dict(help="fit a curve to every candidate peak", **_FLAG)),
    (("--curve",), dict(type=str, default="sinc", help="peak model: sinc or gaussian")),
    (("--fit-frequency",), dict(type=float, help="fit at this frequency instead of the detected peaks")),
    (("-N",), dict(type=int, default=2, help="harmonics of the Z^2_N statistic (Z search only)")),
    (("--force-memmap",), dict(help="keep large result planes in memory-mapped files", **_FLAG)),
    (("-p", "--deorbit-par"), dict(type=str, default=None, help="orbital parameter file (needs PINT: not supported)")),
    (("--loglevel",), dict(type=str, default="WARNING", help="INFO, WARNING, ERROR, CRITICAL or DEBUG")),
    (("--debug",), dict(help="same as --loglevel DEBUG", **_FLAG)),
)


def _common_parser(args=None):
    parser = argparse.ArgumentParser(description="Search for pulsars using the epoch folding or the Z_n^2 algorithm")
    for flags, kw in _CLI_TABLE:
        parser.add_argument(*flags, **kw)
    ns = parser.parse_args(check_negative_numbers_in_args(args))
    if ns.debug:
        ns.loglevel = "DEBUG"
    log.setLevel(ns.loglevel)
    return ns


# ---------------------------------------------------------------------------------------------
# what is searched, and how
# ---------------------------------------------------------------------------------------------
class Statistic(NamedTuple):
    kind: str          # EFPeriodogram.kind
    label: str         # goes into the file name
    nharm: int         # Z^2_n harmonics; 1 for epoch folding (what the reference passes as `n`)
    baseline: float    # expected value of the statistic under noise, subtracted by the peak fits
    func_kwargs: dict  # extra keywords of the search function


def _statistic(func, ns) -> Statistic:
    from .search import z_n_search

    if func is z_n_search:
        return Statistic("Z2n", f"Z2{ns.N}", ns.N, ns.N, {"nharm": ns.N})
    return Statistic("EF", "EF", 1, ns.nbin, {})


def _strategy_name(ns) -> str:
    return "fast" if ns.fast else "ffa" if ns.ffa else "exact"


# output name: root + one piece per rule that applies (hendrics/efsearch.py:1801-1817)
_NAME_RULES: tuple[tuple[Callable, Callable], ...] = (
    (lambda ns, st: True, lambda ns, st: f"_{st.label}"),
    (lambda ns, st: ns.emin is not None or ns.emax is not None,
     lambda ns, st: f"_{_g_or_star(ns.emin)}-{_g_or_star(ns.emax)}keV"),
    (lambda ns, st: ns.fmin is not None or ns.fmax is not None,
     lambda ns, st: f"_{_g_or_star(ns.fmin)}-{_g_or_star(ns.fmax)}Hz"),
    (lambda ns, st: ns.fast, lambda ns, st: "_fast"),
    (lambda ns, st: ns.ffa and not ns.fast, lambda ns, st: "_ffa"),
    (lambda ns, st: ns.mean_fdot is not None and not np.isclose(ns.mean_fdot * 1e10, 0),
     lambda ns, st: f"_fd{ns.mean_fdot * 1e10:g}e-10s-2"),
)


def _g_or_star(value):
    """``%g`` of a bound, ``**`` for an open one.  (The reference formats the ``**`` placeholder itself with ``:g``,
    hendrics/efsearch.py:1804-1811, which raises for a one-sided energy range; the intended name is produced here.)"""
    return HENDRICS_STAR_VALUE if value is None else f"{value:g}"


def output_root(fname, ns, stat) -> str:
    return hen_root(fname) + "".join(text(ns, stat) for applies, text in _NAME_RULES if applies(ns, stat))


class Plane(NamedTuple):
    """What a strategy hands back: the search plane and the numbers the periodogram records."""
    frequencies: np.ndarray
    fdots: object  # array, or 0 for a frequency-only search
    stats: np.ndarray
    step: float
    length: float
    ref_time: float

    @classmethod
    def from_results(cls, results, ref_time):
        if len(results) == 4:
            f, s, step, length = results
            return cls(f, 0, s, step, length, ref_time)
        f, fd, s, step, _fdotstep, length = results
        return cls(f, fd, s, step, length, ref_time)


def _exact(events, ns, stat, func, oversample) -> Plane:
    from .efsearch import folding_search

    results = folding_search(
        events, ns.fmin, ns.fmax, step=ns.step, func=func, oversample=oversample, nbin=ns.nbin, expocorr=ns.expocorr,
        fdotmin=ns.fdotmin or 0, fdotmax=ns.fdotmax or 0, segment_size=ns.segment_size, **stat.func_kwargs)
    return Plane.from_results(results, events.gti[0, 0])


def _fast(events, ns, stat, func, oversample) -> Plane:
    from .efsearch import search_with_qffa

    nbin = ns.nbin
    if nbin / stat.nharm < 8:
        nbin = stat.nharm * 8
        warnings.warn(f"The number of bins is too small for Z search.Increasing to {nbin}")
    # an explicit --fdotmax not above --fdotmin switches the fdot dimension off
    search_fdot = not (ns.fdotmax is not None and ns.fdotmax <= (ns.fdotmin or 0))
    results = search_with_qffa(
        events.time, ns.fmin, ns.fmax, fdot=ns.mean_fdot, fddot=ns.mean_fddot, nbin=nbin, n=stat.nharm, nprof=None,
        npfact=ns.npfact, oversample=oversample, search_fdot=search_fdot, force_memmap=ns.force_memmap)
    return Plane.from_results(results, (events.time[-1] + events.time[0]) / 2)


def _ffa(events, ns, stat, func, oversample) -> Plane:
    from .efsearch import search_with_ffa

    warnings.warn("The Fast Folding Algorithm functionality is experimental. Use with care, and feel free to report "
                  "any issues.")
    results = search_with_ffa(events.time, ns.fmin, ns.fmax, nbin=ns.nbin, n=stat.nharm)
    return Plane.from_results(results, events.time[0])


_STRATEGIES = {"exact": _exact, "fast": _fast, "ffa": _ffa}


# ---------------------------------------------------------------------------------------------
# pipeline stages
# ---------------------------------------------------------------------------------------------
def filter_energy(ev, emin, emax):
    """Events with emin <= energy < emax (``energy`` column, else ``pi``); hendrics/io.py:161-205 reduced to the
    columns this path carries.  Returns ``(events, label)``."""
    for attr, label in (("energy", "Energy"), ("pi", "PI")):
        col = getattr(ev, attr, None)
        if col is not None:
            break
    else:
        warnings.warn("No energy information in event list while filtering between "
                      f"{emin} and {emax}. Definition of events.energy is now based on PI.")
        return ev, ""
    col = np.asarray(col)
    keep = (col >= (-np.inf if emin is None else emin)) & (col < (np.inf if emax is None else emax))
    out = copy.copy(ev)
    out.time = ev.time[keep]
    for attr in ("pi", "energy"):
        if getattr(ev, attr, None) is not None:
            setattr(out, attr, np.asarray(getattr(ev, attr))[keep])
    return out, label


def _rank0() -> bool:
    from . import dist

    return dist.rank() == 0


def _transient_scan(events, ns, stat, oversample, root):
    from .efsearch import transient_search

    tr = transient_search(events.time, ns.fmin, ns.fmax, fdot=0, nbin=ns.nbin, n=stat.nharm,
                          nprof=ns.n_transient_intervals, oversample=oversample, force_memmap=ns.force_memmap)
    if _rank0():
        np.savez(root + "_transient.npz", nave=tr.nave, freqs=tr.freqs, times=tr.times, stats=np.asarray(tr.stats),
                 f0=tr.f0, f1=tr.f1, fdot=tr.fdot, oversample=tr.oversample)


def _dynamical_pass(events, ns, stat, func, oversample, step):
    """The same search in --dynstep windows (hendrics/efsearch.py:1912-1922 draws Dyn.png from it): Dyn.npz here."""
    from .efsearch import dyn_folding_search

    if not _rank0():
        return
    t, f, s = dyn_folding_search(events, ns.fmin, ns.fmax, step=step, func=func, oversample=oversample,
                                 time_step=ns.dynstep, **stat.func_kwargs)
    np.savez("Dyn.npz", times=t, frequencies=f, stats=s)


def fit_peak(frequencies, stats, center_freq, obs_length=None, baseline=0):
    """Least-squares sinc^2 (``obs_length`` given) or Gaussian around a candidate; stands in for ``fit``
    (hendrics/efsearch.py:313-336), which returns astropy models.  Returns a dict of parameters."""
    from scipy import optimize

    y = stats - baseline
    amp0 = y[np.argmin(np.abs(frequencies - center_freq))]
    if obs_length is not None:
        kind, names, p0 = "sinc", ("amplitude", "mean"), [amp0, center_freq]

        def model(x, amp, mean):
            return amp * np.sinc((x - mean) * obs_length) ** 2
    else:
        kind, names = "gaussian", ("amplitude", "mean", "stddev")
        p0 = [amp0, center_freq, 2 * (frequencies[1] - frequencies[0])]

        def model(x, amp, mean, std):
            return amp * np.exp(-0.5 * ((x - mean) / std) ** 2)
    try:
        popt, _ = optimize.curve_fit(model, frequencies, y, p0=p0, maxfev=2000)
    except Exception:
        popt = p0
    return {"model": kind, "baseline": baseline, **{k: float(v) for k, v in zip(names, popt)}}


def _candidate_peaks(ef, ns) -> Optional[np.ndarray]:
    if ns.find_candidates:
        return ef.find_peaks(conflevel=ns.conflevel)[0]
    if ns.fit_frequency is not None:
        ef.peaks, ef.peak_stat = np.array([ns.fit_frequency]), [0]
        return ef.peaks
    return None


def _peak_fits(plane: Plane, peaks, ns, stat, oversample) -> list:
    if not ns.fit_candidates:
        return []
    if peaks is None or len(peaks) == 0:
        warnings.warn("No peaks detected")
        return []
    if np.ndim(plane.frequencies) > 1:
        warnings.warn("Candidate fitting is only available for 1-d (frequency-only) searches")
        return []
    if ns.fast or ns.ffa:
        return []
    curve = ns.curve.lower()
    if curve not in ("sinc", "gaussian"):
        raise ValueError("`--curve` arg must be sinc or gaussian")
    half_width = 5 * oversample * plane.step
    fits = []
    for f in peaks:
        near = np.abs(plane.frequencies - f) < half_width
        fits.append(fit_peak(plane.frequencies[near], plane.stats[near], f,
                             obs_length=plane.length if curve == "sinc" else None, baseline=stat.baseline))
    return fits


def search_one_file(fname, ns, func) -> Optional[str]:
    """Run the whole pipeline on one file; returns the periodogram file name (None when only the transient scan was
    requested)."""
    log.info(f"Treating {fname}")
    stat = _statistic(func, ns)
    strategy = _strategy_name(ns)
    root = output_root(fname, ns, stat)

    ftype, events = get_file_type(fname)
    if ftype not in ("events", "lc"):
        raise ValueError(f"{fname}: only event lists and light curves can be searched")
    mjdref = getattr(events, "mjdref", 0) if ftype == "events" else 0
    if ftype == "events" and (ns.emin is not None or ns.emax is not None):
        events, _ = filter_energy(events, ns.emin, ns.emax)
    oversample = ns.oversample if ns.oversample is not None else (4 * stat.nharm if ns.fast else 2)

    if ns.transient:
        if ftype == "lc":
            log.error("Transient search not yet available for light curves")
        else:
            _transient_scan(events, ns, stat, oversample, root)
            if strategy == "exact":
                return None

    plane = _STRATEGIES[strategy](events, ns, stat, func, oversample)
    if plane.length > ns.dynstep and strategy == "exact":
        _dynamical_pass(events, ns, stat, func, oversample, plane.step)

    span = events.time.max() - events.time.min()
    segment_size = np.min([span, ns.segment_size])
    ef = EFPeriodogram(
        plane.frequencies, plane.stats, stat.kind, ns.nbin, ns.N, fdots=plane.fdots, M=span // segment_size,
        segment_size=segment_size, filename=fname, parfile=ns.deorbit_par, emin=ns.emin, emax=ns.emax, mjdref=mjdref,
        pepoch=mjdref + plane.ref_time / 86400, oversample=ns.oversample)
    ef.upperlim = pf_upper_limit(np.max(plane.stats), events.time.size, n=ns.N)
    ef.ncounts = events.time.size
    peaks = _candidate_peaks(ef, ns)
    ef.best_fits = _peak_fits(plane, peaks, ns, stat, oversample)
    ef.oversample = oversample

    outfile = root + HEN_FILE_EXTENSION
    if _rank0():
        save_folding(ef, outfile)
        if ns.fast:
            from .candidates import analyze_qffa_results

            analyze_qffa_results(outfile)
    return outfile


def _validate(ns, func):
    from .search import z_n_search

    if ns.fit_candidates:
        ns.find_candidates = ns.fit_frequency is None
    if ns.fast and func is not z_n_search:
        raise ValueError("The fast option is only available for z searches")
    if ns.deorbit_par is not None:
        raise NotImplementedError("--deorbit-par (PINT) is outside the B200 search hot path (SURVEY.md section 8f)")


def _common_main(args, func):
    ns = _common_parser(args)
    _validate(ns, func)
    outfiles = (search_one_file(fname, ns, func) for fname in ns.files)
    return [o for o in outfiles if o is not None]


def main_efsearch(args=None):
    """The ``HENefsearch`` command line (hendrics/efsearch.py:1986-1988)."""
    from .search import epoch_folding_search

    return _common_main(args, epoch_folding_search)


def main_zsearch(args=None):
    """The ``HENzsearch`` command line (hendrics/efsearch.py:1991-1993)."""
    from .search import z_n_search

    return _common_main(args, z_n_search)
