"""Host-side mirror of the search entry points of ``hendrics/efsearch.py``.

Same function names, arguments, return tuples, output file names and error behaviour as the
reference for the hot path:

    folding_search           hendrics/efsearch.py:1141-1193
    search_with_qffa_step    hendrics/efsearch.py:876-925
    search_with_qffa         hendrics/efsearch.py:928-1089
    _common_parser/_main     hendrics/efsearch.py:1568-1983
    main_efsearch/zsearch    hendrics/efsearch.py:1986-1993
    analyze_qffa_results     hendrics/efsearch.py:1335-1565 (candidate table, reduced: no astropy Table)

All arithmetic on events runs in the CUDA kernels behind ``include/hx_search.h``; this module only
builds trial grids, shards them over ranks (``dist.py``) and assembles result arrays.
"""

from __future__ import annotations

import argparse
import copy
import logging
import os
import warnings

import numpy as np

from . import _cabi
from .base import (
    HENDRICS_STAR_VALUE,
    a_from_ssig,
    assign_value_if_none,
    check_negative_numbers_in_args,
    find_peaks_in_image,
    fold_detection_level,
    hen_root,
    pf_upper_limit,
    power_confidence_limits,
    z2_n_detection_level,
)
from .io import HEN_FILE_EXTENSION, EFPeriodogram, get_file_type, load_folding, save_folding
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
def folding_search(
    events,
    fmin,
    fmax,
    step=None,
    func=epoch_folding_search,
    oversample=2,
    fdotmin=0,
    fdotmax=0,
    fdotstep=None,
    expocorr=False,
    **kwargs,
):
    """hendrics/efsearch.py:1141-1193.  ``events`` needs ``.time`` and ``.gti``.

    When ``torch.distributed`` is initialised with more than one rank the frequency grid is
    sharded over the ranks and the statistic slices are all-gathered (``dist.sharded_search``);
    every rank returns the full arrays."""
    from . import dist

    times = (events.time - events.gti[0, 0]).astype(np.float64)
    weights = 1
    if hasattr(events, "counts"):
        weights = events.counts

    length = times[-1]

    if step is None:
        step = 1 / oversample / length
    if fdotstep is None:
        fdotstep = 1 / oversample / length**2
    gti = None
    if expocorr:
        gti = (events.gti - events.gti[0, 0]).astype(np.float64)

    # epsilon is needed if fmin == fmax
    epsilon = 1e-8 * step
    trial_freqs = np.arange(fmin, fmax + epsilon, step)
    fdotepsilon = 1e-2 * fdotstep
    trial_fdots = np.arange(fdotmin, fdotmax + fdotepsilon, fdotstep)
    if len(trial_fdots) > 1:
        log.info(f"Searching {len(trial_freqs)} frequencies and {len(trial_fdots)} fdots")
    else:
        log.info(f"Searching {len(trial_freqs)} frequencies")

    results = dist.sharded_search(
        func, times, trial_freqs, fdots=trial_fdots, expocorr=expocorr, gti=gti, weights=weights, **kwargs
    )
    if len(results) == 2:
        frequencies, stats = results
        return frequencies, stats, step, length
    elif len(results) == 3:
        frequencies, fdots, stats = results
        return frequencies, fdots, stats, step, fdotstep, length


TWOPI = 2 * np.pi
D_OMEGA_FACTOR = 2 * np.sqrt(3)


def _save_df_to_csv(df, csv_file, reset=False):
    """hendrics/efsearch.py:189-196."""
    if not os.path.exists(csv_file) or reset:
        mode, header = "w", True
    else:
        mode, header = "a", False
    df.to_csv(csv_file, header=header, index=False, mode=mode)


def decide_binary_parameters(length, freq_range, porb_range, asini_range, fdot_range=[0, 0], NMAX=10,
                             csv_file="db.csv", reset=False):
    """hendrics/efsearch.py:208-260: random (freq, fdot, X, Porb) rows for ``folding_orbital_search``."""
    import pandas as pd

    count = 0
    omega_range = [1 / porb_range[1], 1 / porb_range[0]]
    columns = ["freq", "fdot", "X", "Porb", "done", "max_stat", "min_stat", "best_T0"]
    df = 1 / length
    log.info(f"Recommended frequency steps: {int(np.diff(freq_range)[0] // df + 1)}")
    while count < NMAX:
        if count > 0:  # only the first loop deletes the file
            reset = False
        block_of_data = []
        freq = np.random.uniform(freq_range[0], freq_range[1])
        fdot = np.random.uniform(fdot_range[0], fdot_range[1])
        dX = 1 / (TWOPI * freq)
        nX = int(np.diff(asini_range)[0] // dX) + 1
        Xs = np.random.uniform(asini_range[0], asini_range[1], nX)
        for X in Xs:
            dOmega = 1 / (TWOPI * freq * X * length) * D_OMEGA_FACTOR
            nOmega = int(np.diff(omega_range)[0] // dOmega) + 1
            Omegas = np.random.uniform(omega_range[0], omega_range[1], nOmega)
            for Omega in Omegas:
                block_of_data.append([freq, fdot, X, TWOPI / Omega, False, 0.0, 0.0, 0.0])
        df = pd.DataFrame(block_of_data, columns=columns)
        _save_df_to_csv(df, csv_file, reset=reset)
        count += 1
    return csv_file


def folding_orbital_search(events, parameter_csv_file, chunksize=100, outfile="out.csv", fun=epoch_folding_search,
                           **fun_kwargs):
    """hendrics/efsearch.py:263-310.  For every parameter row the loop over the trial epochs T0 (two demodulation
    passes + one single-trial search each, :294-302) is ONE launch of ``hx_orbital_search_dev`` when ``fun`` is this
    package's ``epoch_folding_search`` / ``z_n_search`` with plain keywords; any other ``fun`` is called T0 by T0 like
    the reference does.  As in the reference the T0s are drawn with ``np.random.uniform`` and chunks whose index does
    not start at 0 are skipped (``chunk["done"][0]`` raises for them, :274-277)."""
    import pandas as pd

    from .search import orbital_fold_stats

    times = (events.time - events.gti[0, 0]).astype(np.float64)
    batched = fun in (epoch_folding_search, z_n_search) and set(fun_kwargs) <= {"nbin", "nharm"}
    for chunk in pd.read_csv(parameter_csv_file, chunksize=chunksize):
        try:
            chunk["done"][0]
        except Exception:
            continue
        for i in range(len(chunk)):
            if chunk["done"][i]:
                continue
            row = chunk.iloc[i]
            freq, fdot, X, Porb = np.array([row["freq"], row["fdot"], row["X"], row["Porb"]], dtype=np.float64)
            dT0 = min(1 / (TWOPI**2 * freq) * Porb / X, Porb / 10)
            max_stats, min_stats, best_T0 = 0, 1e32, None
            T0s = np.random.uniform(0, Porb, int(Porb // dT0 + 1))
            if batched:
                nharm = fun_kwargs.get("nharm", 4) if fun is z_n_search else 0
                all_stats = orbital_fold_stats(times, freq, fdot, X, Porb, T0s, nbin=fun_kwargs.get("nbin", 128),
                                               nharm=nharm)
            else:
                all_stats = []
                for T0 in T0s:
                    new_values = times - X * np.sin(2 * np.pi * (times - T0) / Porb)
                    new_values = new_values - X * np.sin(2 * np.pi * (new_values - T0) / Porb)
                    all_stats.append(fun(new_values, np.array([freq]), fdots=fdot, **fun_kwargs)[1][0])
            for T0, stat in zip(T0s, all_stats):  # the reference's running max / min (:298-302), NaNs never win
                if stat > max_stats:
                    max_stats, best_T0 = stat, T0
                if stat < min_stats:
                    min_stats = stat
            idx = chunk.index[i]
            chunk.iloc[idx, chunk.columns.get_loc("max_stat")] = max_stats
            chunk.iloc[idx, chunk.columns.get_loc("min_stat")] = min_stats
            chunk.iloc[idx, chunk.columns.get_loc("best_T0")] = best_T0
            chunk.iloc[idx, chunk.columns.get_loc("done")] = True
        _save_df_to_csv(chunk, outfile)


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
# candidates
# ---------------------------------------------------------------------------------------------
def get_boundaries_from_level(x, y, level, x0):
    """hendrics/efsearch.py:1299-1332."""
    max_idx = np.argmin(np.abs(x - x0))
    idx = max_idx
    min_x = max_x = x0
    while idx > 0 and y[idx] > level:
        min_x = x[idx]
        idx -= 1
    idx = max_idx
    while idx < y.size and y[idx] > level:
        max_x = x[idx]
        idx += 1
    return min_x, max_x


CAND_COLUMNS = [
    "fname", "mjd", "power", "f", "f_err_n", "f_err_p", "fdot", "fdot_err_n", "fdot_err_p", "fddot", "power_cl_0.9",
    "pulse_amp", "pulse_amp_err", "pulse_amp_cl_0.1", "pulse_amp_cl_0.9", "pulse_amp_ul_0.9", "f_idx", "fdot_idx",
    "fddot_idx",
]


def _analyze_qffa_results(ef, fname=None):
    """Best five candidates of a (typically 2-D) search plane: hendrics/efsearch.py:1335-1545.

    Returns a list of dict rows with the reference's column names (the reference builds an astropy
    Table; astropy is not available here).  2-D error boxes, which the reference traces with
    matplotlib contours (:1262-1296), are reported along the two axes through the peak instead."""
    if not hasattr(ef, "M") or ef.M is None:
        ef.M = 1
    ntrial = ef.stat.size
    if getattr(ef, "oversample", None) is not None:
        ntrial = int(ntrial / ef.oversample)
    if ef.kind == "Z2n":
        detlev = z2_n_detection_level(epsilon=0.001, n=int(ef.N), ntrial=ntrial, n_summed_spectra=int(ef.M))
    else:
        detlev = fold_detection_level(nbin=int(ef.nbin), epsilon=0.001, ntrial=ntrial)
    n_cands = 5
    best_cands = find_peaks_in_image(ef.stat, n=n_cands)
    fddot = ef.fddots if getattr(ef, "fddots", None) is not None else 0
    rows = []
    for i, idx in enumerate(best_cands):
        f_idx = fdot_idx = fddot_idx = 0
        if len(ef.stat.shape) > 1 and ef.stat.shape[0] > 1:
            f_idx, fdot_idx = int(idx[0]), int(idx[1])
            f, fdot = ef.freq[f_idx, fdot_idx], ef.fdots[f_idx, fdot_idx]
            max_stat = ef.stat[f_idx, fdot_idx]
            sig_e1_m, sig_e1 = power_confidence_limits(max_stat, c=0.68, n=ef.N)
            if max_stat < sig_e1_m:
                fmin = fmax = fdotmin = fdotmax = np.nan
            else:
                fmin, fmax = get_boundaries_from_level(ef.freq[f_idx, :], ef.stat[f_idx, :], sig_e1_m, f)
                fdotmin, fdotmax = get_boundaries_from_level(ef.fdots[:, fdot_idx], ef.stat[:, fdot_idx], sig_e1_m, fdot)
        elif len(ef.stat.shape) == 1:
            f_idx = int(idx[0])
            f = ef.freq[f_idx]
            max_stat = ef.stat[f_idx]
            sig_e1_m, sig_e1 = power_confidence_limits(max_stat, c=0.68, n=ef.N)
            fmin, fmax = get_boundaries_from_level(ef.freq, ef.stat, sig_e1_m, f)
            fdot = fdotmin = fdotmax = 0
        else:
            raise ValueError("Did not understand stats shape.")
        if ef.ncounts is None:
            continue
        sig_0, sig_1 = power_confidence_limits(max_stat, c=0.90, n=ef.N)
        amp = amp_err = amp_ul = amp_1 = amp_0 = np.nan
        if max_stat < detlev:
            amp_ul = a_from_ssig(sig_1, ef.ncounts) * 100
        else:
            amp = a_from_ssig(max_stat, ef.ncounts) * 100
            amp_err = a_from_ssig(sig_e1, ef.ncounts) * 100 - amp
            amp_0 = a_from_ssig(sig_0, ef.ncounts) * 100
            amp_1 = a_from_ssig(sig_1, ef.ncounts) * 100
        rows.append(dict(zip(CAND_COLUMNS, [
            ef.filename, ef.pepoch, max_stat, f, fmin - f, fmax - f, fdot, fdotmin - fdot, fdotmax - fdot, fddot,
            sig_0, amp, amp_err, amp_0, amp_1, amp_ul, f_idx, fdot_idx, fddot_idx,
        ])))
        if max_stat < detlev:
            continue  # only one candidate below the detection level
    return rows


def analyze_qffa_results(fname):
    """hendrics/efsearch.py:1548-1565: load, find the best candidates, write ``<fname>_best_cands.csv``."""
    import csv

    ef = load_folding(fname)
    rows = _analyze_qffa_results(ef, fname=fname)
    with open(fname + "_best_cands.csv", "w", newline="") as fobj:
        w = csv.DictWriter(fobj, fieldnames=CAND_COLUMNS)
        w.writeheader()
        for r in rows:
            w.writerow(r)
    return ef, rows


# ---------------------------------------------------------------------------------------------
# CLI (flag surface of hendrics/efsearch.py:1568-1768)
# ---------------------------------------------------------------------------------------------
def _common_parser(args=None):
    description = "Search for pulsars using the epoch folding or the Z_n^2 algorithm"
    parser = argparse.ArgumentParser(description=description)

    parser.add_argument("files", help="List of files", nargs="+")
    parser.add_argument("-f", "--fmin", type=float, required=True, help="Minimum frequency to fold")
    parser.add_argument("-F", "--fmax", type=float, required=True, help="Maximum frequency to fold")
    parser.add_argument("--emin", default=None, type=float, help="Minimum energy (or PI if uncalibrated) to plot")
    parser.add_argument("--emax", default=None, type=float, help="Maximum energy (or PI if uncalibrated) to plot")
    parser.add_argument("--mean-fdot", type=float, required=False, default=0,
                        help="Mean fdot to fold (only useful when using --fast)")
    parser.add_argument("--mean-fddot", type=float, required=False, default=0,
                        help="Mean fddot to fold (only useful when using --fast)")
    parser.add_argument("--fdotmin", type=float, required=False, help="Minimum fdot to fold", default=None)
    parser.add_argument("--fdotmax", type=float, required=False, help="Maximum fdot to fold", default=None)
    parser.add_argument("--dynstep", type=int, required=False, help="Dynamical EF step", default=128)
    parser.add_argument("--npfact", type=int, required=False, help="Size of search parameter space", default=2)
    parser.add_argument("--n-transient-intervals", type=int, required=False, default=None,
                        help="Number of transient intervals to investigate")
    parser.add_argument("-n", "--nbin", default=128, type=int, help="Number of phase bins of the profile")
    parser.add_argument("--segment-size", default=1e32, type=float,
                        help="Size of the event list segment to use (default None, implying the whole observation)")
    parser.add_argument("--step", default=None, type=float,
                        help="Step size of the frequency axis. Defaults to 1/oversample/obs_length. ")
    parser.add_argument("--oversample", default=None, type=float,
                        help="Oversampling factor - frequency resolution improvement w.r.t. the standard FFT's "
                        "1/obs_length.")
    parser.add_argument("--fast", default=False, action="store_true",
                        help="Use a faster folding algorithm. It automatically searches for the first spin "
                        "derivative using an optimized step. This option ignores expocorr, fdotmin/max, "
                        "segment-size, and step")
    parser.add_argument("--ffa", default=False, action="store_true",
                        help="Use *the* Fast Folding Algorithm by Staelin+69. No accelerated search allowed at the moment. Only recommended to search for slow pulsars.")
    parser.add_argument("--transient", default=False, action="store_true",
                        help="Look for transient emission (the dynamic Z search arrays are saved as "
                        "<root>_transient.npz; the reference's animated GIF needs matplotlib)")
    parser.add_argument("--expocorr", default=False, action="store_true",
                        help="Correct for the exposure of the profile bins. This method is *much* slower, but it is useful for very slow pulsars, where data gaps due to occultation or SAA passages can significantly alter the exposure of different profile bins.")
    parser.add_argument("--find-candidates", default=False, action="store_true",
                        help="Find pulsation candidates using thresholding")
    parser.add_argument("--conflevel", default=99, type=float, help="percent confidence level for thresholding [0-100).")
    parser.add_argument("--fit-candidates", default=False, action="store_true",
                        help="Fit the candidate peaks in the periodogram")
    parser.add_argument("--curve", default="sinc", type=str, help="Kind of curve to use (sinc or Gaussian)")
    parser.add_argument("--fit-frequency", type=float, help="Force the candidate frequency to FIT_FREQUENCY")
    # Only relevant to z search
    parser.add_argument("-N", default=2, type=int,
                        help="The number of harmonics to use in the search (the 'N' in Z^2_N; only relevant to Z search!)")
    parser.add_argument("--force-memmap", default=False, action="store_true",
                        help="Force the use of memory-mapped files")
    parser.add_argument("-p", "--deorbit-par", type=str, default=None,
                        help="Deorbit data with this parameter file (requires PINT; not available in hendrics_b200)")
    parser.add_argument("--loglevel", type=str, default="WARNING",
                        help="use given logging level (one between INFO, WARNING, ERROR, CRITICAL, DEBUG)")
    parser.add_argument("--debug", default=False, action="store_true", help="set DEBUG logging level")

    args = check_negative_numbers_in_args(args)
    args = parser.parse_args(args)
    if args.debug:
        args.loglevel = "DEBUG"
    log.setLevel(args.loglevel)
    return args


def filter_energy(ev, emin, emax):
    """hendrics/io.py:161-205 reduced to the columns this path carries."""
    if getattr(ev, "energy", None) is not None:
        energy, elabel = np.asarray(ev.energy), "Energy"
    elif getattr(ev, "pi", None) is not None:
        energy, elabel = np.asarray(ev.pi), "PI"
    else:
        warnings.warn("No energy information in event list while filtering between "
                      f"{emin} and {emax}. Definition of events.energy is now based on PI.")
        return ev, ""
    lo = -np.inf if emin is None else emin
    hi = np.inf if emax is None else emax
    good = (energy >= lo) & (energy < hi)
    new = copy.copy(ev)
    new.time = ev.time[good]
    for attr in ("pi", "energy"):
        if getattr(ev, attr, None) is not None:
            setattr(new, attr, np.asarray(getattr(ev, attr))[good])
    return new, elabel


def _fit_peak(frequencies, stats, center_freq, obs_length=None, baseline=0):
    """Reduced stand-in for ``fit`` (hendrics/efsearch.py:313-336): least-squares sinc^2 (or Gaussian
    when ``obs_length`` is None) around a candidate; returns a plain dict of parameters (the
    reference returns astropy models, which are unavailable here)."""
    from scipy import optimize

    y = stats - baseline
    amp0 = y[np.argmin(np.abs(frequencies - center_freq))]
    if obs_length is not None:
        def model(x, amp, mean):
            return amp * np.sinc((x - mean) * obs_length) ** 2
        p0 = [amp0, center_freq]
        kind = "sinc"
    else:
        df = frequencies[1] - frequencies[0]
        def model(x, amp, mean, std):
            return amp * np.exp(-0.5 * ((x - mean) / std) ** 2)
        p0 = [amp0, center_freq, 2 * df]
        kind = "gaussian"
    try:
        popt, _ = optimize.curve_fit(model, frequencies, y, p0=p0, maxfev=2000)
    except Exception:
        popt = p0
    names = ["amplitude", "mean", "stddev"][: len(popt)]
    out = {"model": kind, "baseline": baseline}
    out.update({k: float(v) for k, v in zip(names, popt)})
    return out


def _common_main(args, func):
    args = _common_parser(args)
    files = args.files
    if args.fit_candidates and args.fit_frequency is None:
        args.find_candidates = True
    elif args.fit_candidates and args.fit_frequency is not None:
        args.find_candidates = False

    if func != z_n_search and args.fast:
        raise ValueError("The fast option is only available for z searches")
    for flag, name in [(args.deorbit_par is not None, "--deorbit-par (PINT)")]:
        if flag:
            raise NotImplementedError(f"{name} is outside the B200 search hot path (SURVEY.md section 8f)")

    outfiles = []
    for i_f, fname in enumerate(files):
        log.info(f"Treating {fname}")
        mjdref = 0
        kwargs = {}
        baseline = args.nbin
        kind = "EF"
        kind_label = kind
        n = 1
        if func == z_n_search:
            n = args.N
            kwargs = {"nharm": args.N}
            baseline = args.N
            kind = "Z2n"
            kind_label = f"Z2{n}"
        ftype, events = get_file_type(fname)

        out_fname = hen_root(fname) + f"_{kind_label}"
        if args.emin is not None or args.emax is not None:
            emin = assign_value_if_none(args.emin, HENDRICS_STAR_VALUE)
            emax = assign_value_if_none(args.emax, HENDRICS_STAR_VALUE)
            out_fname += f"_{emin:g}-{emax:g}keV"
        if args.fmin is not None or args.fmax is not None:
            fmin = assign_value_if_none(args.fmin, HENDRICS_STAR_VALUE)
            fmax = assign_value_if_none(args.fmax, HENDRICS_STAR_VALUE)
            out_fname += f"_{fmin:g}-{fmax:g}Hz"
        if args.fast:
            out_fname += "_fast"
        elif args.ffa:
            out_fname += "_ffa"
        if args.mean_fdot is not None and not np.isclose(args.mean_fdot * 1e10, 0):
            out_fname += f"_fd{args.mean_fdot * 1e10:g}e-10s-2"

        if ftype == "events":
            if hasattr(events, "mjdref"):
                mjdref = events.mjdref
            if args.emin is not None or args.emax is not None:
                events, elabel = filter_energy(events, args.emin, args.emax)
        elif ftype != "lc":
            raise ValueError(f"{fname}: only event lists and light curves can be searched")

        if args.fast:
            oversample = assign_value_if_none(args.oversample, 4 * n)
        else:
            oversample = assign_value_if_none(args.oversample, 2)

        if args.transient and ftype == "lc":
            log.error("Transient search not yet available for light curves")
        if args.transient and ftype == "events":
            # efsearch.py:1831-1845; the reference renders an animated GIF (matplotlib + imageio, both
            # unavailable here): the arrays it is drawn from are saved instead
            tr = transient_search(
                events.time, args.fmin, args.fmax, fdot=0, nbin=args.nbin, n=n, nprof=args.n_transient_intervals,
                oversample=oversample, force_memmap=args.force_memmap,
            )
            from . import dist as _dist

            if _dist.rank() == 0:
                np.savez(out_fname + "_transient.npz", nave=tr.nave, freqs=tr.freqs, times=tr.times,
                         stats=np.asarray(tr.stats), f0=tr.f0, f1=tr.f1, fdot=tr.fdot, oversample=tr.oversample)
            if not args.fast and not args.ffa:
                continue

        if not args.fast and not args.ffa:
            fdotmin = args.fdotmin if args.fdotmin is not None else 0
            fdotmax = args.fdotmax if args.fdotmax is not None else 0
            results = folding_search(
                events, args.fmin, args.fmax, step=args.step, func=func, oversample=oversample, nbin=args.nbin,
                expocorr=args.expocorr, fdotmin=fdotmin, fdotmax=fdotmax, segment_size=args.segment_size, **kwargs,
            )
            ref_time = events.gti[0, 0]
        elif args.fast:
            fdotmin = args.fdotmin if args.fdotmin is not None else 0
            fdotmax = args.fdotmax if args.fdotmax is not None else 0
            search_fdot = True
            if args.fdotmax is not None and fdotmax <= fdotmin:
                search_fdot = False
            nbin = args.nbin
            if nbin / n < 8:
                nbin = n * 8
                warnings.warn(f"The number of bins is too small for Z search.Increasing to {nbin}")
            results = search_with_qffa(
                events.time, args.fmin, args.fmax, fdot=args.mean_fdot, fddot=args.mean_fddot, nbin=nbin, n=n,
                nprof=None, npfact=args.npfact, oversample=oversample, search_fdot=search_fdot,
                force_memmap=args.force_memmap,
            )
            ref_time = (events.time[-1] + events.time[0]) / 2
        elif args.ffa:
            warnings.warn(
                "The Fast Folding Algorithm functionality is experimental. Use"
                " with care, and feel free to report any issues."
            )
            results = search_with_ffa(events.time, args.fmin, args.fmax, nbin=args.nbin, n=n)
            ref_time = events.time[0]

        length = events.time.max() - events.time.min()
        segment_size = np.min([length, args.segment_size])
        M = length // segment_size

        fdots = 0
        if len(results) == 4:
            frequencies, stats, step, length = results
        elif len(results) == 6:
            frequencies, fdots, stats, step, fdotsteps, length = results

        if length > args.dynstep and not (args.fast or args.ffa):
            # efsearch.py:1912-1922: the reference runs the dynamical search here, draws "Dyn.png" with matplotlib
            # and discards the arrays.  Without matplotlib the arrays the figure is drawn from are saved instead.
            from . import dist as _dist

            if _dist.rank() == 0:
                dyn_t, dyn_f, dyn_s = dyn_folding_search(
                    events, args.fmin, args.fmax, step=step, func=func, oversample=oversample, time_step=args.dynstep,
                    **kwargs,
                )
                np.savez("Dyn.npz", times=dyn_t, frequencies=dyn_f, stats=dyn_s)

        efperiodogram = EFPeriodogram(
            frequencies, stats, kind, args.nbin, args.N, fdots=fdots, M=M, segment_size=segment_size, filename=fname,
            parfile=args.deorbit_par, emin=args.emin, emax=args.emax, mjdref=mjdref,
            pepoch=mjdref + ref_time / 86400, oversample=args.oversample,
        )
        efperiodogram.upperlim = pf_upper_limit(np.max(stats), events.time.size, n=args.N)
        efperiodogram.ncounts = events.time.size
        best_peaks = None
        if args.find_candidates:
            best_peaks, best_stat = efperiodogram.find_peaks(conflevel=args.conflevel)
        elif args.fit_frequency is not None:
            best_peaks = np.array([args.fit_frequency])
            efperiodogram.peaks = best_peaks
            efperiodogram.peak_stat = [0]

        best_models = []
        detected = best_peaks is not None and len(best_peaks) > 0
        if args.fit_candidates and not detected:
            warnings.warn("No peaks detected")
        elif args.fit_candidates and np.ndim(frequencies) > 1:
            warnings.warn("Candidate fitting is only available for 1-d (frequency-only) searches")
        elif args.fit_candidates and not (args.fast or args.ffa):
            search_width = 5 * oversample * step
            for f in best_peaks:
                good = np.abs(frequencies - f) < search_width
                if args.curve.lower() == "sinc":
                    best_fun = _fit_peak(frequencies[good], stats[good], f, obs_length=length, baseline=baseline)
                elif args.curve.lower() == "gaussian":
                    best_fun = _fit_peak(frequencies[good], stats[good], f, baseline=baseline)
                else:
                    raise ValueError("`--curve` arg must be sinc or gaussian")
                best_models.append(best_fun)

        efperiodogram.best_fits = best_models
        efperiodogram.oversample = oversample

        from . import dist

        if dist.rank() == 0:
            save_folding(efperiodogram, out_fname + HEN_FILE_EXTENSION)
            if args.fast:
                analyze_qffa_results(out_fname + HEN_FILE_EXTENSION)
        outfiles.append(out_fname + HEN_FILE_EXTENSION)
    return outfiles


def main_efsearch(args=None):
    """Main function called by the `HENefsearch` command line script."""
    return _common_main(args, epoch_folding_search)


def main_zsearch(args=None):
    """Main function called by the `HENzsearch` command line script."""
    return _common_main(args, z_n_search)
