"""Host-side mirror of the search entry points of ``hendrics/efsearch.py``.

Same function names, arguments, return tuples, output file names and error behaviour as the
reference for the hot path:

    folding_search           hendrics/efsearch.py:1141-1193
    search_with_qffa_step    hendrics/efsearch.py:876-925
    search_with_qffa         hendrics/efsearch.py:928-1089
    _common_parser/_main     hendrics/efsearch.py:1568-1983   -> driver.py (stand-alone) / rebind.py (inside HENDRICS)
    main_efsearch/zsearch    hendrics/efsearch.py:1986-1993   -> driver.py
    analyze_qffa_results     hendrics/efsearch.py:1335-1565   -> candidates.py

All arithmetic on events runs in the CUDA kernels behind ``include/hx_search.h``; this module only
builds trial grids, shards them over ranks (``dist.py``) and assembles result arrays.
"""

from __future__ import annotations

import copy
import logging
import os
import warnings

import numpy as np

from . import _cabi
from .base import find_peaks_in_image  # noqa: F401  (re-exported like hendrics/efsearch.py:33)
from .ffa import ffa_search, h_test  # noqa: F401  (h_test is re-exported like hendrics/efsearch.py:46,85)
from .search import epoch_folding_search, z_n_search

log = logging.getLogger("hendrics_b200")

__all__ = [
    "folding_search", "search_with_qffa", "search_with_qffa_step", "check_phase_error_after_casting_to_double",
    "dyn_folding_search", "transient_search", "TransientResults", "analyze_qffa_results", "main_efsearch", "main_zsearch", "z_n_search",
    "epoch_folding_search", "qffa_centres", "search_with_ffa", "h_test",
]


def check_phase_error_after_casting_to_double(tref, f, fdot=0):
    """hendrics/efsearch.py:199-205."""
    times = np.array(np.random.normal(tref, 0.1, 1000), dtype=np.longdouble)
    times_dbl = times.astype(np.double)
    phase = times * f + 0.5 * times**2 * fdot
    phase_dbl = times_dbl * np.double(f) + 0.5 * times_dbl**2 * np.double(fdot)
    return np.max(np.abs(phase_dbl - phase))


# ---------------------------------------------------------------------------------------------
# exact fold search
# ---------------------------------------------------------------------------------------------
class TrialGrid:
    """The (f, fdot) trial values of an exact search, built the way hendrics/efsearch.py:1159-1177 builds them -- the
    grid IS the interface (statistics are compared element by element), so the rules are kept to the letter:

    * default steps: ``1 / oversample / length`` in f, ``1 / oversample / length**2`` in fdot, ``length`` being the
      time of the last event after the GTI start;
    * ``np.arange(lo, hi + eps, step)`` with ``eps = 1e-8 * step`` (f) and ``1e-2 * fdotstep`` (fdot), so that a
      degenerate range (lo == hi) still holds one trial.
    """

    F_EPS, FDOT_EPS = 1e-8, 1e-2

    def __init__(self, length, fmin, fmax, step=None, oversample=2, fdotmin=0, fdotmax=0, fdotstep=None):
        self.length = length
        self.step = 1 / oversample / length if step is None else step
        self.fdotstep = 1 / oversample / length**2 if fdotstep is None else fdotstep
        self.freqs = np.arange(fmin, fmax + self.F_EPS * self.step, self.step)
        self.fdots = np.arange(fdotmin, fdotmax + self.FDOT_EPS * self.fdotstep, self.fdotstep)

    def describe(self):
        what = f"{len(self.freqs)} frequencies"
        return what + (f" and {len(self.fdots)} fdots" if len(self.fdots) > 1 else "")


def folding_search(events, fmin, fmax, step=None, func=epoch_folding_search, oversample=2, fdotmin=0, fdotmax=0,
                   fdotstep=None, expocorr=False, **kwargs):
    """Drop-in for hendrics/efsearch.py:1141-1193: search ``events`` (needs ``.time`` and ``.gti``; ``.counts`` makes
    it a light curve, i.e. a weighted search) over the trial grid of :class:`TrialGrid` with ``func``.  Returns
    ``(frequencies, stats, step, length)`` for a single fdot, else ``(frequencies, fdots, stats, step, fdotstep,
    length)`` with 2-D arrays of shape (n_fdot, n_freq).

    With ``torch.distributed`` initialised on several ranks the frequency axis is sharded over them and the statistic
    slices are all-gathered (``dist.sharded_search``); every rank returns the full arrays."""
    from . import dist

    t_start = events.gti[0, 0]
    times = (events.time - t_start).astype(np.float64)
    grid = TrialGrid(times[-1], fmin, fmax, step=step, oversample=oversample, fdotmin=fdotmin, fdotmax=fdotmax,
                     fdotstep=fdotstep)
    log.info("Searching " + grid.describe())
    search_kwargs = dict(
        fdots=grid.fdots, expocorr=expocorr,
        gti=(events.gti - t_start).astype(np.float64) if expocorr else None,
        weights=getattr(events, "counts", 1),
    )
    found = dist.sharded_search(func, times, grid.freqs, **search_kwargs, **kwargs)
    if len(found) == 2:
        return found[0], found[1], grid.step, grid.length
    return found[0], found[1], found[2], grid.step, grid.fdotstep, grid.length


from .orbital import decide_binary_parameters, folding_orbital_search  # noqa: E402,F401  (hendrics/efsearch.py:208-310)


def time_intervals_from_gtis(gtis, segment_size, fraction_step=1):
    """``stingray.gti.time_intervals_from_gtis`` [UPSTREAM-RECALLED]: start/stop of all
    ``segment_size``-long intervals that fit in the GTIs."""
    starts, stops = [], []
    for g in np.asarray(gtis, dtype=np.float64):
        if g[1] - g[0] < segment_size:
            continue
        newtimes = np.arange(g[0], g[1] - segment_size + 1e-9 * segment_size, np.float64(segment_size) * fraction_step)
        starts.extend(newtimes)
        stops.extend(newtimes + segment_size)
    return np.asarray(starts), np.asarray(stops)


def dyn_folding_search(events, fmin, fmax, step=None, func=epoch_folding_search, oversample=2, time_step=128, **kwargs):
    """hendrics/efsearch.py:1196-1243 without the matplotlib figure: the same search repeated in
    ``time_step``-long windows.  Returns ``(times, frequencies, stats[n_windows, n_freq])``.

    With this package's own searches and windows too small to fill a GPU (the CLI's default: 128 s) all windows are
    folded by ONE launch (``search.windowed_searches`` -> ``hx_fold_windows_dev``) and come back in one copy; otherwise
    ``func`` is called window by window like in the reference."""
    from . import search as _search

    if step is None:
        step = 1 / oversample / time_step
    gti = np.copy(events.gti)
    length = np.diff(gti, axis=1)
    if not np.any(length > time_step):
        gti = np.array([[gti[0, 0], gti[-1, 1]]])
    start, stop = time_intervals_from_gtis(gti, time_step)
    frequencies = np.arange(fmin, fmax, step)
    ev_time = np.asanyarray(events.time)
    times = (start + stop) / 2

    if func in (z_n_search, epoch_folding_search) and len(start) > 0 and len(start) <= 65535 and frequencies.size > 0:
        kw = dict(kwargs)
        nharm = kw.pop("nharm", 4) if func is z_n_search else 0
        nbin = kw.pop("nbin", 128)
        plain = (not kw.pop("expocorr", False) and kw.pop("gti", None) is None and np.isscalar(kw.pop("weights", 1))
                 and np.all(np.asarray(kw.pop("fdots", 0)) == 0) and np.isinf(kw.pop("segment_size", np.inf)) and not kw)
        small = ev_time.size / len(start) * frequencies.size <= _search.WINDOWED_MAX_UNITS_PER_WINDOW
        if plain and small and len(start) * frequencies.size * 8 <= (1 << 30):
            res = _search.windowed_searches(nharm, ev_time, frequencies, start, stop, nbin=nbin)
            if res is not None:
                return times, frequencies, res[0][:, 0, :]

    stats = []
    is_sorted = ev_time.size < 2 or bool(np.all(ev_time[1:] >= ev_time[:-1]))
    for st, sp in zip(start, stop):
        if is_sorted:  # same selection as the mask below, without a pass over the whole list per window
            times_filt = ev_time[np.searchsorted(ev_time, st, side="left"):np.searchsorted(ev_time, sp, side="left")]
        else:
            times_filt = ev_time[(ev_time >= st) & (ev_time < sp)]
        try:
            frequencies, stat = func(times_filt, frequencies, **kwargs)
            stats.append(stat)
        except _cabi.HxError:
            raise
        except Exception:
            stats.append(np.zeros_like(frequencies))
    times = (start + stop) / 2
    return times, frequencies, np.array(stats)


# ---------------------------------------------------------------------------------------------
# --fast
# ---------------------------------------------------------------------------------------------
def _shift_grids(nbin, npfact, oversample, search_fdot):
    linbinshifts = np.linspace(-nbin * npfact, nbin * npfact, int(oversample * npfact))
    if search_fdot:
        quabinshifts = np.linspace(-nbin * npfact, nbin * npfact, int(oversample * npfact))
    else:
        quabinshifts = np.array([0])
    return np.meshgrid(linbinshifts, quabinshifts, indexing="ij")


def _qffa_device_call(times_d, t_first, t_last, centres, nbin, nprof, L, Q, n, want_profiles=False):
    import torch

    cents = np.ascontiguousarray(np.asarray(centres, dtype=np.float64).reshape(-1, 3))
    mf, mfd, mfdd = (np.ascontiguousarray(cents[:, i]) for i in range(3))
    Lf, Qf = _cabi.as_f64(L.ravel()), _cabi.as_f64(Q.ravel())
    nc = cents.shape[0]
    stats = torch.empty((nc, Lf.size), dtype=torch.float64, device=times_d.device)
    prof = torch.empty((nc, nprof, nbin), dtype=torch.int32, device=times_d.device) if want_profiles else None
    with torch.cuda.device(times_d.device):
        rc = _cabi.lib().hx_qffa_search_dev(
            times_d.data_ptr(), times_d.numel(), float(t_first), float(t_last), _cabi.dptr(mf), _cabi.dptr(mfd),
            _cabi.dptr(mfdd), nc, int(nbin), int(nprof), _cabi.dptr(Lf), _cabi.dptr(Qf), Lf.size, int(n),
            stats.data_ptr(), prof.data_ptr() if want_profiles else None, torch.cuda.current_stream().cuda_stream,
        )
    _cabi.check(rc)
    stats = stats.cpu().numpy().reshape((nc,) + L.shape)
    if want_profiles:
        return stats, (prof.to(torch.int64) & 0xFFFFFFFF).cpu().numpy()
    return stats


def _to_device(times, device=None):
    import torch

    from .search import _device

    return torch.from_numpy(np.ascontiguousarray(times, dtype=np.float64)).to(_device(device))


def search_with_qffa_step(
    times,
    mean_f,
    mean_fdot=0,
    mean_fddot=0,
    nbin=16,
    nprof=64,
    npfact=2,
    oversample=8,
    n=1,
    search_fdot=True,
    return_profiles=False,
):
    """Single step of the quasi-fast folding algorithm (hendrics/efsearch.py:876-925).

    ``times`` must be fp64, sorted and centred like in ``search_with_qffa``.  Returns
    ``(fgrid, fdotgrid, stats)`` of shape ``(nL, nQ)`` [+ the (nprof, nbin) profiles]."""
    times = np.asarray(times, dtype=np.float64)
    t1, t0 = times[-1], times[0]
    L, Q = _shift_grids(nbin, npfact, oversample, search_fdot)
    dphi = 1 / nbin
    delta_t = (t1 - t0) / 2
    bin_to_frequency = dphi / delta_t
    bin_to_fdot = 2 * dphi / delta_t**2
    res = _qffa_device_call(_to_device(times), t0, t1, [(mean_f, mean_fdot, mean_fddot)], nbin, nprof, L, Q, n,
                            want_profiles=return_profiles)
    stats = res[0][0] if return_profiles else res[0]
    out = (L * bin_to_frequency + mean_f, Q * bin_to_fdot + mean_fdot, stats)
    if return_profiles:
        out = out + (res[1][0],)
    return out


def qffa_centres(f0, f1, length, npfact=2, fdot=0, fddot=0):
    """Centre (mean_f, mean_fdot, mean_fddot) list and step of hendrics/efsearch.py:1004-1032."""
    frequency = (f0 + f1) / 2
    step = 4 * npfact / length
    niter = int(np.rint((f1 - f0) / step)) + 2
    allvalues = list(range(-(niter // 2), niter // 2))
    if allvalues == []:
        allvalues = [0]
    cents = []
    for i in allvalues:
        offset = step * i
        fdot_offset = 0
        cents.append((np.double(frequency + offset + 0.12 * step), np.double(fdot + fdot_offset), np.double(fddot)))
    return cents, step


def search_with_qffa(
    times,
    f0,
    f1,
    fdot=0,
    fddot=0,
    nbin=16,
    nprof=None,
    npfact=2,
    oversample=8,
    n=1,
    search_fdot=True,
    t0=None,
    t1=None,
    silent=False,
    force_memmap=False,
):
    """'Quite fast folding' algorithm (hendrics/efsearch.py:928-1089).

    The serial loop over centre frequencies of the reference becomes one ``hx_qffa_search_dev`` call
    per rank: centres are sharded over ranks when ``torch.distributed`` is initialised, and each
    rank streams the events once per group of centres."""
    from . import dist

    if nprof is None:
        # nprof > total_delta_phi * nbin to get all the signal inside one bin in a sub-integration
        nprof = 4 * 2 * nbin * npfact

    times = copy.deepcopy(np.asarray(times))

    if t0 is None:
        t0 = times.min()
    if t1 is None:
        t1 = times.max()
    meantime = (t1 + t0) / 2
    times -= meantime

    maxerr = check_phase_error_after_casting_to_double(np.max(times), f1, fdot)
    if maxerr > 1 / nbin / 10:
        warnings.warn(f"Maximum error on the phase expected when casting to double: {maxerr}")
        warnings.warn("Casting to double produces non-negligible phase errors. Please use shorter light curves.")

    times = times.astype(np.double)
    length = t1 - t0
    cents, step = qffa_centres(f0, f1, length, npfact, fdot, fddot)

    L, Q = _shift_grids(nbin, npfact, oversample, search_fdot)
    dphi = 1 / nbin
    delta_t = (times[-1] - times[0]) / 2
    bin_to_frequency = dphi / delta_t
    bin_to_fdot = 2 * dphi / delta_t**2

    times_d = _to_device(times)

    def local(cents_slice):
        return _qffa_device_call(times_d, times[0], times[-1], cents_slice, nbin, nprof, L, Q, n)

    stats = dist.sharded_rows(local, np.asarray(cents, dtype=np.float64))  # (n_centres, nL, nQ)

    nL = L.shape[0]
    all_fgrid_shape = (nL * len(cents), L.shape[1])
    log.info(f"Initializing result arrays of shape {all_fgrid_shape}")
    if all_fgrid_shape[0] * all_fgrid_shape[1] > 1e7 or force_memmap:
        import tempfile

        log.info("Large result arrays detected, using memory-mapped files to reduce memory usage.")
        names = [tempfile.NamedTemporaryFile("w+").name + "_hen.npy" for _ in range(3)]
        all_fgrid, all_fdotgrid, all_stats = (
            np.lib.format.open_memmap(nm, mode="w+", dtype=np.float64, shape=all_fgrid_shape) for nm in names
        )
    else:
        all_fgrid = np.zeros(all_fgrid_shape)
        all_fdotgrid = np.zeros(all_fgrid_shape)
        all_stats = np.zeros(all_fgrid_shape)
    for ii, (mean_f, mean_fdot, _) in enumerate(cents):
        all_fgrid[ii * nL : (ii + 1) * nL, :] = L * bin_to_frequency + mean_f
        all_fdotgrid[ii * nL : (ii + 1) * nL, :] = Q * bin_to_fdot + mean_fdot
        all_stats[ii * nL : (ii + 1) * nL, :] = stats[ii]

    step = np.median(np.diff(all_fgrid[:, 0]))
    fdotstep = np.median(np.diff(all_fdotgrid[0]))
    if search_fdot:
        return all_fgrid.T, all_fdotgrid.T, all_stats.T, step, fdotstep, length
    else:
        return all_fgrid.T[0], all_stats.T[0], step, length


# ---------------------------------------------------------------------------------------------
# transient search (widening: SURVEY.md section 8f, rank 4)
# ---------------------------------------------------------------------------------------------
def search_with_ffa(times, f0, f1, nbin=16, n=1, t0=None, t1=None):
    """Fast-folding-algorithm search: mirror of ``hendrics/efsearch.py:1092-1138``.

    The light curve (bin time = 1/f1/nbin) is built on the GPU by ``hx_hist1d_dev`` and never leaves it;
    ``ffa_search`` folds it (``hx_ffa_dev``).  Returns ``(frequencies, stats, None, length)`` like the reference;
    ``n`` is accepted and ignored like there (the reference's ``ffa_search`` always computes Z^2_2).
    """
    import torch

    times = np.asarray(times)
    if t0 is None:
        t0 = times[0]
    if t1 is None:
        t1 = times[-1]
    length = np.asarray(t1 - t0).astype(np.double)
    p0 = 1 / f1
    p1 = 1 / f0
    dt = p0 / nbin
    bins = int(np.rint(length / dt))
    if times.dtype == np.float64:
        # `(times - t0).astype(np.double)` is one fp64 subtraction per event: done inside the histogram kernel (x0 = t0)
        x, x0 = _to_device(times), float(t0)
    else:  # extended-precision times: subtract on the host in their own precision, then round, as the reference does
        x, x0 = _to_device((times - t0).astype(np.double)), 0.0
    counts = torch.empty(bins, dtype=torch.int32, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().hx_hist1d_dev(x.data_ptr(), x.numel(), x0, 0.0, float(length), bins, counts.data_ptr(), stream))
    bin_periods, stats = ffa_search(counts, dt, p0, p1)
    return 1 / bin_periods, stats, None, length


class TransientResults:
    """Same fields as hendrics/efsearch.py:488-496."""

    oversample: int = None
    f0: float = None
    f1: float = None
    fdot: float = None
    nave: np.array = None
    freqs: np.array = None
    times: np.array = None
    stats: np.array = None


def transient_search(
    times, f0, f1, fdot=0, nbin=16, nprof=None, n=1, t0=None, t1=None, oversample=4, force_memmap=False
):
    """Search for transient pulsations (hendrics/efsearch.py:499-650).

    For every centre frequency the events are folded into ``nprof`` time slices (K2) and the Z^2_n of
    the sub-profiles averaged over 1, 2, 4, ... slices is computed from per-slice harmonic coefficients
    (``hx_transient_search_dev``).  Centres are sharded over ranks like in ``search_with_qffa``."""
    from . import dist

    if nprof is None:
        nprof = 4 * 2 * nbin

    times = copy.deepcopy(np.asarray(times))
    if t0 is None:
        t0 = times.min()
    if t1 is None:
        t1 = times.max()
    meantime = (t1 + t0) / 2
    times -= meantime

    maxerr = check_phase_error_after_casting_to_double(np.max(times), f1, fdot)
    log.info(f"Maximum error on the phase expected when casting to double: {maxerr}")
    if maxerr > 1 / nbin / 10:
        warnings.warn("Casting to double produces non-negligible phase errors. Please use shorter light curves.")

    times = times.astype(np.double)
    length = t1 - t0
    frequency = (f0 + f1) / 2
    step = 1 / length / oversample
    niter = int(np.rint((f1 - f0) / step)) + 2
    allvalues = list(range(-(niter // 2), niter // 2))
    if allvalues == []:
        allvalues = [0]
    cents = np.array([[np.double(frequency + step * i + 0.12 * step), np.double(fdot + 0)] for i in allvalues])
    dt = (times[-1] - times[0]) / nprof
    # `nprof = int(2 ** np.log2(nprof))` (hendrics/efsearch.py:436-438) is not a power-of-two truncation, but the
    # floating-point round trip does drop the LAST time slice for some nprof (10 -> 9, 100 -> 99, 200 -> 199): the
    # events are histogrammed into nprof slices and the averaging sees the first nprof_used of them
    nprof_used = int(2 ** np.log2(nprof))
    n_levels = int(np.log2(nprof_used))
    nave = 2 ** np.arange(n_levels)

    times_d = _to_device(times)

    def local(rows):
        import torch

        rows = np.ascontiguousarray(rows, dtype=np.float64)
        mf, mfd = np.ascontiguousarray(rows[:, 0]), np.ascontiguousarray(rows[:, 1])
        out = torch.empty((rows.shape[0], n_levels, nprof_used), dtype=torch.float64, device=times_d.device)
        with torch.cuda.device(times_d.device):
            rc = _cabi.lib().hx_transient_search_dev(
                times_d.data_ptr(), times_d.numel(), float(times[0]), float(times[-1]), _cabi.dptr(mf),
                _cabi.dptr(mfd), rows.shape[0], int(nbin), int(nprof), int(nprof_used), int(n), int(n_levels), out.data_ptr(),
                torch.cuda.current_stream().cuda_stream,
            )
        _cabi.check(rc)
        return out.cpu().numpy()

    all_results = dist.sharded_rows(local, cents)  # (centres, levels, nprof)
    final_shape = (nave.size, all_results.shape[2], all_results.shape[0])
    if np.prod(final_shape) > 1e7 or force_memmap:
        import tempfile

        tmp = tempfile.NamedTemporaryFile(delete=True).name + "_hen.npy"
        stats = np.lib.format.open_memmap(tmp, mode="w+", dtype=np.float64, shape=final_shape)
    else:
        stats = np.empty(final_shape)
    for i in range(nave.size):
        stats[i] = all_results[:, i, :].T

    results = TransientResults()
    results.oversample = oversample
    results.f0 = f0
    results.f1 = f1
    results.fdot = fdot
    results.nave = nave
    results.freqs = cents[:, 0].copy()
    results.times = dt * np.arange(all_results.shape[2])
    results.stats = stats
    return results


# ---------------------------------------------------------------------------------------------
# candidate post-processing and the command line live in their own modules; same names as hendrics/efsearch.py
# ---------------------------------------------------------------------------------------------
from .candidates import (  # noqa: E402,F401
    CAND_COLUMNS,
    _analyze_qffa_results,
    analyze_qffa_results,
    get_boundaries_from_level,
    get_xy_boundaries_from_level,
    print_qffa_results,
)
from .driver import _common_main, _common_parser, filter_energy, main_efsearch, main_zsearch  # noqa: E402,F401
