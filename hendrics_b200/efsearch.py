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
# --fast and the transient scan: fastfold.py
# ---------------------------------------------------------------------------------------------
from .fastfold import (  # noqa: E402,F401
    TransientResults,
    _qffa_device_call,
    _shift_grids,
    _to_device,
    check_phase_error_after_casting_to_double,
    qffa_centres,
    search_with_qffa,
    search_with_qffa_step,
    transient_search,
)


# ---------------------------------------------------------------------------------------------
# --ffa
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
