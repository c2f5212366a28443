"""CPU ORACLE for the HENDRICS pulsar periodicity-search hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product (``hendrics_b200``) never does: it fails loudly when
its CUDA library is missing.

What it restates (file:line are relative to the reference tree, ``/root/reference``):

* in-tree, read directly:
    hendrics/efsearch.py:345-363   ``mod`` / ``shift_and_sum``
    hendrics/efsearch.py:366-404   ``z_n_fast``                    (unbinned-capable Z^2_n)
    hendrics/efsearch.py:817-851   ``_fast_step``
    hendrics/efsearch.py:854-873   ``_fast_phase`` / ``_fast_phase_fdot`` / ``_fast_phase_fddot``
    hendrics/efsearch.py:876-925   ``search_with_qffa_step``
    hendrics/efsearch.py:928-1089  ``search_with_qffa``
    hendrics/efsearch.py:1141-1193 ``folding_search``
    hendrics/efsearch.py:407-650   ``_average_and_z_sub_search`` / ``_transient_search_step`` / ``transient_search``
    hendrics/efsearch.py:199-205   ``check_phase_error_after_casting_to_double``
    hendrics/ffa.py:38-157         cached harmonics, ``_z_n_fast_cached`` (binned Z^2_n)
    hendrics/ffa.py:210-395        ``ffa_step`` / ``_ffa`` / ``ffa`` / ``_quick_rebin`` / ``ffa_search`` (--ffa)
    hendrics/efsearch.py:1092-1138 ``search_with_ffa``;  hendrics/base.py:621-660 ``hist1d_numba_seq``
    hendrics/base.py:667-693       ``_hist2d_numba_seq`` / ``hist2d_numba_seq``
    hendrics/base.py:1109-1148     ``find_peaks_in_image`` (rough / no-skimage branch)
* out-of-tree (third-party dependency ``stingray>=2.3.2``, pyproject.toml:41, imported at
  hendrics/efsearch.py:15-18; its source is NOT under /root/reference and it is not
  installable here): ``stingray.pulse.search._folding_search / z_n_search /
  epoch_folding_search / _pulse_phase_fast / _profile_fast`` and
  ``stingray.pulse.pulsar.ef_profile_stat / z_n_binned_events``, restated from the
  published stingray 2.x algorithm.

PARITY STATUS: **parity unpinned at the stingray boundary** -- the reference's own
tests hold no golden vector for z_n_search/epoch_folding_search (unseeded inputs, 1/T
tolerances).  What *is* pinned, in tests/test_oracle.py, are the
reference's doctest known answers (Z^2_n == 50 for the 1+sin profile, histogram ==
np.histogram2d on in-range data, toy-image top-2 peaks, ...).

Declared upstream ambiguities:
  A1  ``rezero_to_first_event``  times -= times[0] inside ``_folding_search``      (default True)
  A2  ``drop_last_event``        segment mask is half-open => t == t[-1] dropped   (default True)
  A3  fdot term evaluated as ((0.5*t)*t)*fd, each product separately rounded       (no FMA)
  A4  a phase that rounds to exactly 1.0 (or phase*nbin == nbin for non power-of-two nbin)
      makes upstream's bincount one element too long; here the count is clamped into bin nbin-1.

numba is used when importable (like the reference's hendrics/compat/compatibility.py:6-52),
otherwise the same bodies run as plain Python/numpy.
"""

from __future__ import annotations

import copy
import math
import warnings

import numpy as np

try:  # same policy as hendrics/compat/compatibility.py:6-52
    from numba import njit, prange

    HAS_NUMBA = True
except Exception:  # pragma: no cover - numba is present in this image
    HAS_NUMBA = False

    def njit(*args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]

        def deco(f):
            return f

        return deco

    prange = range


# --------------------------------------------------------------------------------------
# stingray.pulse.search._pulse_phase_fast / _profile_fast  [UPSTREAM-RECALLED, Appendix A]
# --------------------------------------------------------------------------------------
@njit()
def pulse_phase_fast(time, f, fdot, buffer_array):
    """phase = t*f + 0.5*t*t*fdot ; phase -= floor(phase)   (separately rounded, A3)."""
    for i in range(len(time)):
        buffer_array[i] = time[i] * f + 0.5 * time[i] * time[i] * fdot
        buffer_array[i] -= np.floor(buffer_array[i])
    return buffer_array


@njit()
def profile_fast(phase, nbin):
    """int64 histogram of floor(phase*nbin); overflow clamped into nbin-1 (A4)."""
    prof = np.zeros(nbin, dtype=np.int64)
    for i in range(len(phase)):
        b = np.int64(np.floor(phase[i] * nbin))
        if b >= nbin:
            b = nbin - 1
        if b < 0:
            b = 0
        prof[b] += 1
    return prof


def profile_fast_upstream(phase, nbin):
    """The literal upstream formulation (sentinel + bincount); used to show that
    ``profile_fast`` equals it whenever no phase rounds up to the right edge."""
    phase_bin = np.zeros(len(phase) + 2, dtype=np.int64)
    phase_bin[-1] = nbin - 1
    phase_bin[: len(phase)] = np.floor(phase * nbin).astype(np.int64)
    bc = np.bincount(phase_bin)
    bc[0] -= 1
    bc[-1] -= 1
    return bc


# --------------------------------------------------------------------------------------
# stingray.pulse.pulsar.ef_profile_stat / z_n_binned_events   [UPSTREAM-RECALLED]
# in-tree twin of the Z formula: hendrics/ffa.py:38-157
# --------------------------------------------------------------------------------------
def ef_profile_stat(profile, err=None):
    """chi^2 = sum((p - mean)^2 / err^2), err = sqrt(mean) by default."""
    profile = np.asarray(profile, dtype=np.float64)
    mean = np.mean(profile)
    if err is None:
        err = np.sqrt(mean)
    return np.sum((profile - mean) ** 2 / err**2)


def cached_sin_harmonics(nbin, z_n_n):
    """hendrics/ffa.py:38-46."""
    twopiphases = np.pi * 2 * np.arange(nbin) * (1.0 / nbin)
    return np.tile(np.sin(twopiphases), z_n_n)


def cached_cos_harmonics(nbin, z_n_n):
    """hendrics/ffa.py:49-56."""
    twopiphases = np.pi * 2 * np.arange(nbin) * (1.0 / nbin)
    return np.tile(np.cos(twopiphases), z_n_n)


@njit()
def _z_n_fast_cached(norm, cached_sin, cached_cos, n=2):
    """hendrics/ffa.py:59-104 -- Z^2_n of a binned profile, harmonic k read as the
    stride-k slice of the n-times tiled table."""
    total_norm = np.sum(norm)
    result = 0.0
    N = norm.size
    for k in range(1, n + 1):
        result += (
            np.sum(cached_cos[: N * k : k] * norm) ** 2
            + np.sum(cached_sin[: N * k : k] * norm) ** 2
        )
    return 2 / total_norm * result


def z_n_fast_cached(norm, n=2):
    """hendrics/ffa.py:107-136."""
    norm = np.asarray(norm, dtype=np.float64)
    return _z_n_fast_cached(norm, cached_sin_harmonics(norm.size, n), cached_cos_harmonics(norm.size, n), n)


@njit()
def _z_n_fast_cached_all(norm, cached_sin, cached_cos, ks):
    """hendrics/ffa.py:139-157."""
    total_norm = np.sum(norm)
    all_zs = np.zeros(ks.size)
    N = norm.size
    total_sum = 0.0
    for k in ks:
        local_z = (
            np.sum(cached_cos[: N * k : k] * norm) ** 2
            + np.sum(cached_sin[: N * k : k] * norm) ** 2
        )
        all_zs[k - 1] = local_z + total_sum
        total_sum += local_z
    return 2 / total_norm * all_zs


def z_n_fast_cached_all(norm, nmax=20):
    """hendrics/ffa.py:160-194."""
    norm = np.asarray(norm, dtype=np.float64)
    ks = np.arange(1, nmax + 1, dtype=np.int64)
    return ks, _z_n_fast_cached_all(
        norm, cached_sin_harmonics(norm.size, nmax), cached_cos_harmonics(norm.size, nmax), ks
    )


def h_test(norm, nmax=20):
    """hendrics/ffa.py:197-211."""
    ks, zs = z_n_fast_cached_all(norm, nmax=nmax)
    hs = zs - 4 * ks + 4
    idx = np.argmax(hs)
    return hs[idx], ks[idx]


def z_n_binned_events(profile, n):
    """stingray.pulse.pulsar.z_n_binned_events: same formula as hendrics/ffa.py:139-157."""
    return z_n_fast_cached(np.asarray(profile, dtype=np.float64), n)


@njit(fastmath=True)
def z_n_fast(phase, norm, n=2):
    """hendrics/efsearch.py:366-404 (k*phi by repeated addition)."""
    total_norm = np.sum(norm)
    result = 0.0
    kph = np.zeros_like(phase)
    for k in range(1, n + 1):
        kph += phase
        result += np.sum(np.cos(kph) * norm) ** 2 + np.sum(np.sin(kph) * norm) ** 2
    return 2 / total_norm * result


def z_n_events(phase, n):
    """Unbinned Z^2_n of photon phases (cycles); oracle of the ``binned=False`` extension:
    hendrics/efsearch.py:366-404 with norm == 1, evaluated without fastmath."""
    ph = 2 * np.pi * np.asarray(phase, dtype=np.float64)
    res = 0.0
    for k in range(1, n + 1):
        res += np.sum(np.cos(k * ph)) ** 2 + np.sum(np.sin(k * ph)) ** 2
    return 2.0 / ph.size * res


# --------------------------------------------------------------------------------------
# stingray.pulse.search._folding_search / z_n_search / epoch_folding_search
# --------------------------------------------------------------------------------------
def select_segments(times, segment_size=np.inf, rezero_to_first_event=True, drop_last_event=True):
    """Event selection of upstream ``_folding_search`` (Appendix A: A1, A2).

    Returns the re-zeroed float64 times and the list of boolean masks of the segments
    that are actually folded."""
    times = np.asarray(times)
    if rezero_to_first_event:
        times = (times - times[0]).astype(np.float64)
    else:
        times = times.astype(np.float64)
    length = times[-1]
    if length < segment_size:
        segment_size = length
    start_times = np.arange(times[0], times[-1], segment_size)
    masks = []
    for s in start_times:
        if drop_last_event:
            good = (times >= s) & (times < s + segment_size)
        else:
            good = (times >= s) & (times <= s + segment_size)
        ts = times[good]
        if len(ts) < 1 or ts[-1] - ts[0] < 0.2 * segment_size:
            continue
        masks.append(good)
    return times, masks


@njit()
def _stat_from_profile(prof, nharm, ctab, stab):
    """nharm == 0: ef_profile_stat; else z_n_binned_events (stride-k table reads)."""
    nbin = prof.size
    if nharm == 0:
        mean = 0.0
        for j in range(nbin):
            mean += prof[j]
        mean /= nbin
        chi = 0.0
        for j in range(nbin):
            d = prof[j] - mean
            chi += d * d / mean
        return chi
    tot = 0.0
    for j in range(nbin):
        tot += prof[j]
    res = 0.0
    for k in range(1, nharm + 1):
        c = 0.0
        s = 0.0
        for j in range(nbin):
            idx = (j * k) % nbin
            c += ctab[idx] * prof[j]
            s += stab[idx] * prof[j]
        res += c * c + s * s
    return 2.0 / tot * res


@njit()
def _fold_one(ts, f, fd, nbin, prof):
    """One trial: fused _pulse_phase_fast + _profile_fast (bit-identical arithmetic)."""
    prof[:] = 0
    for i in range(ts.size):
        ph = ts[i] * f + 0.5 * ts[i] * ts[i] * fd
        ph -= np.floor(ph)
        b = np.int64(np.floor(ph * nbin))
        if b >= nbin:
            b = nbin - 1
        if b < 0:
            b = 0
        prof[b] += 1


@njit()
def _search_serial(ts, freqs, fdots, nbin, nharm, ctab, stab, stats, profiles, want_profiles):
    """The upstream hot loop: serial over fdots then freqs (Appendix A)."""
    prof = np.zeros(nbin, dtype=np.int64)
    for i in range(fdots.size):
        for j in range(freqs.size):
            _fold_one(ts, freqs[j], fdots[i], nbin, prof)
            stats[i, j] += _stat_from_profile(prof, nharm, ctab, stab)
            if want_profiles:
                for b in range(nbin):
                    profiles[i, j, b] += prof[b]


@njit(parallel=True)
def _search_parallel(ts, freqs, fdots, nbin, nharm, ctab, stab, stats):
    """Same arithmetic, trials spread over host threads (CPU-baseline 'all cores' leg)."""
    nf = freqs.size
    ntr = fdots.size * nf
    for g in prange(ntr):
        i = g // nf
        j = g - i * nf
        prof = np.zeros(nbin, dtype=np.int64)
        _fold_one(ts, freqs[j], fdots[i], nbin, prof)
        stats[i, j] += _stat_from_profile(prof, nharm, ctab, stab)


def _tables(nbin):
    tw = np.pi * 2 * np.arange(nbin) * (1.0 / nbin)
    return np.cos(tw), np.sin(tw)


def folding_search_core(
    times,
    frequencies,
    fdots=0,
    nbin=128,
    nharm=0,
    segment_size=np.inf,
    rezero_to_first_event=True,
    drop_last_event=True,
    return_profiles=False,
    parallel=False,
):
    """Upstream ``_folding_search`` fast branch, for EF (nharm=0) or Z^2_n (nharm>=1).

    Returns (fgrid, stats) when a single fdot is given, else (fgrid, fdgrid, stats) with
    shape (n_fdot, n_freq).  With ``return_profiles`` the int64 profiles summed over
    segments, shape (n_fdot, n_freq, nbin), are appended."""
    frequencies = np.atleast_1d(np.asarray(frequencies, dtype=np.float64))
    fdots = np.atleast_1d(np.asarray(fdots, dtype=np.float64))
    fgrid, fdgrid = np.meshgrid(frequencies, fdots)
    stats = np.zeros_like(fgrid)
    profiles = np.zeros(fgrid.shape + (nbin,), dtype=np.int64) if return_profiles else np.zeros((1, 1, 1), dtype=np.int64)
    times, masks = select_segments(times, segment_size, rezero_to_first_event, drop_last_event)
    ctab, stab = _tables(nbin)
    count = 0
    for good in masks:
        ts = np.ascontiguousarray(times[good])
        if parallel and not return_profiles:
            _search_parallel(ts, frequencies, fdots, nbin, nharm, ctab, stab, stats)
        else:
            _search_serial(ts, frequencies, fdots, nbin, nharm, ctab, stab, stats, profiles, return_profiles)
        count += 1
    with np.errstate(invalid="ignore", divide="ignore"):
        stats = stats / count  # Appendix A: unguarded; no folded segment gives 0/0 = NaN (and numpy's warning)
    if fgrid.shape[0] == 1:
        out = (fgrid.flatten(), stats.flatten())
    else:
        out = (fgrid, fdgrid, stats)
    if return_profiles:
        out = out + (profiles,)
    return out


# --------------------------------------------------------------------------------------
# The SLOW branch of upstream's searches: stingray.pulse.pulsar.pulse_phase / phase_exposure /
# fold_events, reached when ``expocorr`` is set or ``weights`` is an array (light-curve input:
# hendrics/efsearch.py:1155-1157, 1164-1167).
# [UPSTREAM-RECALLED] -- restated from the published stingray 2.x sources, not checkable here.
# --------------------------------------------------------------------------------------
def pulse_phase(times, *frequency_derivatives, ph0=0, to_1=True):
    """stingray.pulse.pulsar.pulse_phase: ph = ph0 + sum_i times**(i+1) * f_i / (i+1)!."""
    ph = ph0
    for i_f, f in enumerate(frequency_derivatives):
        ph = ph + 1 / math.factorial(i_f + 1) * times ** (i_f + 1) * f
    if to_1:
        ph = ph - np.floor(ph)
    return ph


def phase_exposure(start_time, stop_time, period, nbin=16, gti=None):
    """stingray.pulse.pulsar.phase_exposure: relative exposure of each phase bin, long-double
    arithmetic on fp64 inputs, accumulated into an fp64 array."""
    if gti is None:
        gti = np.array([[start_time, stop_time]])
    start_time = np.longdouble(start_time)
    stop_time = np.longdouble(stop_time)
    period = np.longdouble(period)
    gti = np.array(gti, dtype=np.longdouble)

    expo = np.zeros(nbin)
    phs = np.linspace(0, 1, nbin + 1)
    phs = np.array(list(zip(phs[0:-1], phs[1:])))

    good = np.logical_and(gti[:, 0] < stop_time, gti[:, 1] > start_time)
    gti = gti[good]
    for g in gti:
        g0 = g[0]
        g1 = g[1]
        if g0 < start_time:
            g0 = start_time
        if g1 > stop_time:
            g1 = stop_time
        length = g1 - g0
        nraw = length / period
        nper = nraw.astype(int)
        expo += nper / nbin
        start_phase = np.fmod(g0 / period, 1)
        end_phase = nraw - nper + start_phase
        limits = [[start_phase, end_phase]]
        if end_phase > 1:
            limits = [[0, end_phase - 1], [start_phase, 1]]
        for l in limits:
            l0 = l[0]
            l1 = l[1]
            goodbins = np.logical_and(phs[:, 0] <= l1, phs[:, 1] >= l0)
            idxs = np.arange(len(phs), dtype=int)[goodbins]
            for i in idxs:
                start = np.max([phs[i, 0], l0])
                stop = np.min([phs[i, 1], l1])
                w = stop - start
                expo[i] += w
    return expo / np.max(expo)


def fold_events(times, *frequency_derivatives, nbin=16, weights=1, gti=None, ref_time=0, expocorr=False):
    """stingray.pulse.pulsar.fold_events, ``mode="ef"``: (weighted) phase histogram over the
    explicit bin edges ``linspace(0, 1, nbin + 1)`` (np.histogram => searchsorted semantics, right
    edge of the last bin included), optionally divided by the per-bin exposure.
    Returns (bin centres, profile, profile error)."""
    times = np.asarray(times, dtype=np.float64)
    if gti is None:
        gti = [[times[0], times[-1]]]
    gti = np.asanyarray(gti)
    if not isinstance(weights, np.ndarray) and not isinstance(weights, (list, tuple)):
        weights = weights * np.ones(len(times))
    weights = np.asarray(weights)
    gti = gti - ref_time
    times = times - ref_time
    dt = np.abs(times[1] - times[0])
    start_time = times[0]
    stop_time = times[-1] + dt
    phases = pulse_phase(times, *frequency_derivatives, to_1=True)
    gti_phases = pulse_phase(gti, *frequency_derivatives, to_1=False)
    start_phase, stop_phase = pulse_phase(np.array([start_time, stop_time]), *frequency_derivatives, to_1=False)
    bins = np.linspace(0, 1, nbin + 1)
    raw_profile, bins = np.histogram(phases, bins=bins, weights=weights)
    raw_profile_err = np.sqrt(np.histogram(phases, bins=bins, weights=weights**2)[0])
    bins = (bins[1:] + bins[:-1]) / 2
    if expocorr:
        expo_norm = phase_exposure(start_phase, stop_phase, 1, nbin, gti=gti_phases)
    else:
        expo_norm = 1
    return bins, raw_profile / expo_norm, raw_profile_err / expo_norm


def folding_search_slow(
    times,
    frequencies,
    fdots=0,
    nbin=128,
    nharm=0,
    segment_size=np.inf,
    expocorr=False,
    gti=None,
    weights=1,
    rezero_to_first_event=True,
    drop_last_event=True,
    return_profiles=False,
):
    """Upstream ``_folding_search(..., use_times=True)``: per trial ``fold_events`` on the segment's
    times (array-valued keyword arguments as long as ``times`` are masked like them; the GTIs are
    passed through UNSHIFTED, as upstream does) followed by ``ef_profile_stat`` (nharm == 0) or
    ``z_n_binned_events``."""
    frequencies = np.atleast_1d(np.asarray(frequencies, dtype=np.float64))
    fdots = np.atleast_1d(np.asarray(fdots, dtype=np.float64))
    fgrid, fdgrid = np.meshgrid(frequencies, fdots)
    stats = np.zeros_like(fgrid)
    profiles = np.zeros(fgrid.shape + (nbin,)) if return_profiles else None
    n_all = len(times)
    times, masks = select_segments(times, segment_size, rezero_to_first_event, drop_last_event)
    count = 0
    for good in masks:
        ts = times[good]
        w = weights
        if isinstance(weights, (np.ndarray, list, tuple)) and len(weights) == n_all:
            w = np.asarray(weights)[good]
        g = gti
        if gti is not None and len(gti) == n_all:  # upstream masks ANY keyword of that length
            g = np.asarray(gti)[good]
        for i in range(fgrid.shape[0]):
            for j in range(fgrid.shape[1]):
                prof = fold_events(ts, fgrid[i, j], fdgrid[i, j], nbin=nbin, weights=w, gti=g, expocorr=expocorr)[1]
                if nharm == 0:
                    stats[i, j] += ef_profile_stat(prof)
                else:
                    stats[i, j] += z_n_binned_events(prof, nharm)
                if return_profiles:
                    profiles[i, j] += prof
        count += 1
    with np.errstate(invalid="ignore", divide="ignore"):
        stats = stats / count  # Appendix A: unguarded; no folded segment gives 0/0 = NaN (and numpy's warning)
    if fgrid.shape[0] == 1:
        out = (fgrid.flatten(), stats.flatten())
    else:
        out = (fgrid, fdgrid, stats)
    if return_profiles:
        out = out + (profiles,)
    return out


def _slow_branch(expocorr, weights):
    return bool(expocorr) or isinstance(weights, (np.ndarray, list, tuple))


def epoch_folding_search(times, frequencies, nbin=128, segment_size=np.inf, expocorr=False, gti=None, weights=1, fdots=0, **kw):
    """stingray.pulse.search.epoch_folding_search."""
    if expocorr and gti is None:
        raise ValueError("To calculate exposure correction, you need to specify the GTIs")
    if _slow_branch(expocorr, weights):
        return folding_search_slow(times, frequencies, fdots=fdots, nbin=nbin, nharm=0, segment_size=segment_size,
                                   expocorr=expocorr, gti=gti, weights=weights, **kw)
    return folding_search_core(times, frequencies, fdots=fdots, nbin=nbin, nharm=0, segment_size=segment_size, **kw)


def z_n_search(times, frequencies, nharm=4, nbin=128, segment_size=np.inf, expocorr=False, weights=1, gti=None, fdots=0, **kw):
    """stingray.pulse.search.z_n_search (binned)."""
    if expocorr and gti is None:
        raise ValueError("To calculate exposure correction, you need to specify the GTIs")
    if _slow_branch(expocorr, weights):
        return folding_search_slow(times, frequencies, fdots=fdots, nbin=nbin, nharm=nharm, segment_size=segment_size,
                                   expocorr=expocorr, gti=gti, weights=weights, **kw)
    return folding_search_core(times, frequencies, fdots=fdots, nbin=nbin, nharm=nharm, segment_size=segment_size, **kw)


def folding_search(
    event_times,
    gti_start,
    fmin,
    fmax,
    step=None,
    func=epoch_folding_search,
    oversample=2,
    fdotmin=0,
    fdotmax=0,
    fdotstep=None,
    **kwargs,
):
    """hendrics/efsearch.py:1141-1193 on plain arrays (events.time, events.gti[0, 0])."""
    times = (np.asarray(event_times) - gti_start).astype(np.float64)
    length = times[-1]
    if step is None:
        step = 1 / oversample / length
    if fdotstep is None:
        fdotstep = 1 / oversample / length**2
    epsilon = 1e-8 * step
    trial_freqs = np.arange(fmin, fmax + epsilon, step)
    fdotepsilon = 1e-2 * fdotstep
    trial_fdots = np.arange(fdotmin, fdotmax + fdotepsilon, fdotstep)
    results = func(times, trial_freqs, fdots=trial_fdots, **kwargs)
    if len(results) == 2:
        frequencies, stats = results
        return frequencies, stats, step, length
    frequencies, fdots, stats = results
    return frequencies, fdots, stats, step, fdotstep, length


# --------------------------------------------------------------------------------------
# hendrics/efsearch.py:263-310 folding_orbital_search (T0 loop of one parameter row)
# --------------------------------------------------------------------------------------
def sin_correctly_rounded(a):
    """fp64 sin rounded from an extended-precision evaluation (x87 long double: 64-bit mantissa, so the fp64 result is
    the correctly rounded one except within ~2^-11 ulp of a rounding boundary).  ``np.sin`` on fp64 is NOT this
    function: numpy dispatches to SIMD kernels with errors up to ~1 ulp that depend on the CPU's vector extensions
    (0.12 % of random arguments differ in the last bit on the AVX-512 build measured in tests/test_oracle.py), so the
    reference's demodulated times are only defined up to the libm / SIMD flavour it runs on."""
    return np.sin(np.asarray(a, dtype=np.longdouble)).astype(np.float64)


def orbital_demodulate(times, T0, X, Porb, sin=np.sin):
    """efsearch.py:295-296: two fixed-point iterations of the Roemer-delay removal (``sin``: the reference's
    ``np.sin``; ``sin_correctly_rounded`` to measure how much the choice of sin implementation moves)."""
    new_values = times - X * sin(2 * np.pi * (times - T0) / Porb)
    new_values = new_values - X * sin(2 * np.pi * (new_values - T0) / Porb)
    return new_values


def orbital_stats(times, freq, fdot, X, Porb, T0s, fun=None, return_profiles=False, sin=np.sin, **fun_kwargs):
    """efsearch.py:294-302: ``fun(new_values, np.array([freq]), fdots=fdot, **fun_kwargs)[1][0]`` for every T0."""
    if fun is None:
        fun = epoch_folding_search
    stats = np.zeros(len(T0s))
    profs = []
    for k, T0 in enumerate(T0s):
        nv = orbital_demodulate(times, T0, X, Porb, sin=sin)
        res = fun(nv, np.array([freq]), fdots=fdot, return_profiles=return_profiles, **fun_kwargs)
        stats[k] = res[1][0]
        if return_profiles:
            profs.append(res[2][0, 0])
    return (stats, np.array(profs)) if return_profiles else stats


# --------------------------------------------------------------------------------------
# --fast ("quasi-fast folding")
# --------------------------------------------------------------------------------------
ONE_SIXTH = 1 / 6


def fast_phase(ts, mean_f):
    """hendrics/efsearch.py:870-873."""
    phases = ts * mean_f
    return phases - np.floor(phases)


def fast_phase_fdot(ts, mean_f, mean_fdot=0):
    """hendrics/efsearch.py:854-857."""
    phases = ts * mean_f + 0.5 * ts * ts * mean_fdot
    return phases - np.floor(phases)


def fast_phase_fddot(ts, mean_f, mean_fdot=0, mean_fddot=0):
    """hendrics/efsearch.py:863-867."""
    tssq = ts * ts
    phases = ts * mean_f + 0.5 * tssq * mean_fdot + ONE_SIXTH * tssq * ts * mean_fddot
    return phases - np.floor(phases)


@njit()
def _hist2d_seq(H, x, y, bins, ranges):
    """hendrics/base.py:667-677 (right edges excluded, truncating index)."""
    delta0 = 1 / ((ranges[0, 1] - ranges[0, 0]) / bins[0])
    delta1 = 1 / ((ranges[1, 1] - ranges[1, 0]) / bins[1])
    for t in range(x.size):
        i = (x[t] - ranges[0, 0]) * delta0
        j = (y[t] - ranges[1, 0]) * delta1
        if 0 <= i < bins[0] and 0 <= j < bins[1]:
            H[int(i), int(j)] += 1
    return H


def histogram2d(x, y, bins, range):
    """hendrics/base.py:680-693, 913-945 (unweighted numba variant)."""
    H = np.zeros((bins[0], bins[1]), dtype=np.uint64)
    return _hist2d_seq(
        H,
        np.asarray(x, dtype=np.float64),
        np.asarray(y, dtype=np.float64),
        np.asarray(list(bins), dtype=np.int64),
        np.asarray(range, dtype=np.float64),
    )


@njit()
def shift_and_sum(repeated_profiles, lshift, qshift, splat_prof, base_shift, quadbaseshift):
    """hendrics/efsearch.py:350-363."""
    nprof = repeated_profiles.shape[0]
    nbin = splat_prof.size
    twonbin = nbin * 2
    splat_prof[:] = 0.0
    for k in range(nprof):
        total_shift = base_shift[k] * lshift + quadbaseshift[k] * qshift
        total_shift = np.mod(np.rint(total_shift), nbin)
        total_shift_int = int(total_shift)
        splat_prof[:] += repeated_profiles[k, nbin - total_shift_int : twonbin - total_shift_int]
    return splat_prof


@njit()
def _fast_step_impl(profiles, L, Q, nbin, n, return_splats):
    twopiphases = 2 * np.pi * np.arange(0, 1, 1 / nbin)
    cached_cos = np.zeros(n * nbin)
    cached_sin = np.zeros(n * nbin)
    for i in range(n):
        cached_cos[i * nbin : (i + 1) * nbin] = np.cos(twopiphases)
        cached_sin[i * nbin : (i + 1) * nbin] = np.sin(twopiphases)
    stats = np.zeros_like(L)
    repeated_profiles = np.hstack((profiles, profiles, profiles))
    nprof = repeated_profiles.shape[0]
    base_shift = np.linspace(-1, 1, nprof)
    quad_base_shift = base_shift**2
    splats = np.zeros((L.shape[0] if return_splats else 1, L.shape[1] if return_splats else 1, nbin))
    for i in range(L.shape[0]):
        splat_prof = np.zeros(nbin)
        for j in range(L.shape[1]):
            splat_prof = shift_and_sum(repeated_profiles, L[i, j], Q[i, j], splat_prof, base_shift, quad_base_shift)
            # note: reference passes (cos, sin) into the (sin, cos) slots -- symmetric (efsearch.py:848)
            stats[i, j] = _z_n_fast_cached(splat_prof, cached_cos, cached_sin, n)
            if return_splats:
                splats[i, j, :] = splat_prof
    return stats, splats


def fast_step(profiles, L, Q, nbin, n=2, return_splats=False):
    """hendrics/efsearch.py:817-851."""
    stats, splats = _fast_step_impl(np.asarray(profiles, dtype=np.float64), L, Q, nbin, n, return_splats)
    if return_splats:
        return stats, splats
    return stats


def search_with_qffa_step(
    times, mean_f, mean_fdot=0, mean_fddot=0, nbin=16, nprof=64, npfact=2, oversample=8, n=1, search_fdot=True, return_profiles=False
):
    """hendrics/efsearch.py:876-925."""
    if mean_fddot != 0:
        phases = fast_phase_fddot(times, mean_f, mean_fdot, mean_fddot)
    elif mean_fdot != 0:
        phases = fast_phase_fdot(times, mean_f, mean_fdot)
    else:
        phases = fast_phase(times, mean_f)
    profiles = histogram2d(phases, times, range=[[0, 1], [times[0], times[-1]]], bins=(nbin, nprof)).T
    t1, t0 = times[-1], times[0]
    linbinshifts = np.linspace(-nbin * npfact, nbin * npfact, int(oversample * npfact))
    if search_fdot:
        quabinshifts = np.linspace(-nbin * npfact, nbin * npfact, int(oversample * npfact))
    else:
        quabinshifts = np.array([0])
    dphi = 1 / nbin
    delta_t = (t1 - t0) / 2
    bin_to_frequency = dphi / delta_t
    bin_to_fdot = 2 * dphi / delta_t**2
    L, Q = np.meshgrid(linbinshifts, quabinshifts, indexing="ij")
    stats = fast_step(profiles, L, Q, nbin, n=n)
    out = (L * bin_to_frequency + mean_f, Q * bin_to_fdot + mean_fdot, stats)
    if return_profiles:
        out = out + (profiles,)
    return out


def check_phase_error_after_casting_to_double(tref, f, fdot=0, rng=None):
    """hendrics/efsearch.py:199-205 (the reference draws from the unseeded global RNG)."""
    rng = np.random if rng is None else rng
    times = np.array(rng.normal(tref, 0.1, 1000), dtype=np.longdouble)
    times_dbl = times.astype(np.double)
    phase = times * f + 0.5 * times**2 * fdot
    phase_dbl = times_dbl * np.double(f) + 0.5 * times_dbl**2 * np.double(fdot)
    return np.max(np.abs(phase_dbl - phase))


def qffa_centres(f0, f1, length, npfact=2, fdot=0, fddot=0):
    """Centre (mean_f, mean_fdot, mean_fddot) list of hendrics/efsearch.py:1004-1032."""
    frequency = (f0 + f1) / 2
    step = 4 * npfact / length
    niter = int(np.rint((f1 - f0) / step)) + 2
    allvalues = list(range(-(niter // 2), niter // 2))
    if allvalues == []:
        allvalues = [0]
    cents = []
    for i in allvalues:
        offset = step * i
        cents.append((np.double(frequency + offset + 0.12 * step), np.double(fdot + 0), np.double(fddot)))
    return cents, step


def search_with_qffa(
    times, f0, f1, fdot=0, fddot=0, nbin=16, nprof=None, npfact=2, oversample=8, n=1, search_fdot=True, t0=None, t1=None, **_ignored
):
    """hendrics/efsearch.py:928-1089 (in-memory result arrays)."""
    if nprof is None:
        nprof = 4 * 2 * nbin * npfact
    times = copy.deepcopy(np.asarray(times))
    if t0 is None:
        t0 = times.min()
    if t1 is None:
        t1 = times.max()
    meantime = (t1 + t0) / 2
    times = times - meantime
    maxerr = check_phase_error_after_casting_to_double(np.max(times), f1, fdot)
    if maxerr > 1 / nbin / 10:
        warnings.warn(f"Maximum error on the phase expected when casting to double: {maxerr}")
    times = times.astype(np.double)
    length = t1 - t0
    cents, step = qffa_centres(f0, f1, length, npfact, fdot, fddot)
    all_f, all_fd, all_s = [], [], []
    for mean_f, mean_fdot, mean_fddot in cents:
        fgrid, fdotgrid, stats = search_with_qffa_step(
            times, mean_f, mean_fdot=mean_fdot, mean_fddot=mean_fddot, nbin=nbin, nprof=nprof, npfact=npfact,
            oversample=oversample, n=n, search_fdot=search_fdot,
        )
        all_f.append(fgrid)
        all_fd.append(fdotgrid)
        all_s.append(stats)
    all_fgrid = np.concatenate(all_f, axis=0)
    all_fdotgrid = np.concatenate(all_fd, axis=0)
    all_stats = np.concatenate(all_s, axis=0)
    step = np.median(np.diff(all_fgrid[:, 0]))
    fdotstep = np.median(np.diff(all_fdotgrid[0]))
    if search_fdot:
        return all_fgrid.T, all_fdotgrid.T, all_stats.T, step, fdotstep, length
    return all_fgrid.T[0], all_stats.T[0], step, length


# --------------------------------------------------------------------------------------
# transient search (hendrics/efsearch.py:407-650)
# --------------------------------------------------------------------------------------
def average_and_z_sub_search(profiles, n=2):
    """hendrics/efsearch.py:407-465.  `int(2 ** np.log2(nprof))` is no power-of-two truncation, but it does drop the
    last slice for some nprof (10 -> 9, 100 -> 99: floating-point round trip), exactly as in the reference."""
    profiles = np.asarray(profiles, dtype=np.float64)
    nprof = len(profiles)
    nprof = int(2 ** np.log2(nprof))
    profiles = profiles[:nprof]
    nbin = len(profiles[0])
    n_log_ave_max = int(np.log2(nprof))
    results = np.zeros((n_log_ave_max, nprof))
    twopiphases = 2 * np.pi * np.arange(0, 1, 1 / nbin)
    n_ave = 2 ** np.arange(n_log_ave_max)
    for ave_i in range(len(n_ave)):
        n_ave_i = n_ave[ave_i]
        shape_0 = int(profiles.shape[0] / n_ave_i)
        for i in range(shape_0):
            new_profiles = np.sum(profiles[i * n_ave_i : (i + 1) * n_ave_i], axis=0)
            if np.sum(new_profiles) == 0:
                continue
            z = z_n_fast(twopiphases, new_profiles, n)
            results[ave_i, i * n_ave_i : (i + 1) * n_ave_i] = z
    return n_ave, results


def transient_search_step(times, mean_f, mean_fdot=0, nbin=16, nprof=64, n=1):
    """hendrics/efsearch.py:468-485."""
    phases = fast_phase_fdot(times, mean_f, mean_fdot)
    profiles = histogram2d(phases, times, range=[[0, 1], [times[0], times[-1]]], bins=(nbin, nprof)).T
    return average_and_z_sub_search(profiles, n=n)


def transient_centres(f0, f1, length, oversample=4, fdot=0):
    """Centre list of hendrics/efsearch.py:573-592."""
    frequency = (f0 + f1) / 2
    step = 1 / length / oversample
    niter = int(np.rint((f1 - f0) / step)) + 2
    allvalues = list(range(-(niter // 2), niter // 2))
    if allvalues == []:
        allvalues = [0]
    return [(np.double(frequency + step * i + 0.12 * step), np.double(fdot + 0)) for i in allvalues]


def transient_search(times, f0, f1, fdot=0, nbin=16, nprof=None, n=1, t0=None, t1=None, oversample=4):
    """hendrics/efsearch.py:499-650; returns (nave, freqs, times, stats[nave, nprof, n_centres])."""
    if nprof is None:
        nprof = 4 * 2 * nbin
    times = copy.deepcopy(np.asarray(times))
    if t0 is None:
        t0 = times.min()
    if t1 is None:
        t1 = times.max()
    meantime = (t1 + t0) / 2
    times = times - meantime
    times = times.astype(np.double)
    length = t1 - t0
    cents = transient_centres(f0, f1, length, oversample, fdot)
    dt = (times[-1] - times[0]) / nprof
    all_results, all_freqs = [], []
    for mean_f, mean_fdot in cents:
        nave, results = transient_search_step(times, mean_f, mean_fdot=mean_fdot, nbin=nbin, nprof=nprof, n=n)
        all_results.append(results)
        all_freqs.append(mean_f)
    all_results = np.asarray(all_results)  # (centres, levels, nprof)
    tgrid = dt * np.arange(all_results.shape[2])
    stats = np.empty((nave.size, all_results.shape[2], all_results.shape[0]))
    for i in range(nave.size):
        stats[i] = all_results[:, i, :].T
    return nave, np.asarray(all_freqs), tgrid, stats


# --------------------------------------------------------------------------------------
# candidates
# --------------------------------------------------------------------------------------
def find_peaks_in_image(image, n=5):
    """hendrics/base.py:1140-1146 (the branch taken when skimage is absent / rough=True)."""
    image = np.asarray(image)
    best_cands = [np.unravel_index(idx, image.shape) for idx in np.argpartition(image.flatten(), -n)[-n:]]
    best_stats = [image[bci] for bci in best_cands]
    return np.asarray(best_cands)[np.argsort(best_stats)][::-1]


def topk_flat(stats, k):
    """Deterministic top-k used for the 'identical candidate list' check: value descending,
    ties broken by ascending flat index (last bullet)."""
    flat = np.asarray(stats).ravel()
    k = min(k, flat.size)
    order = np.lexsort((np.arange(flat.size), -flat))[:k]
    return order.astype(np.int64), flat[order]


def search_best_peaks(x, stat, threshold):
    """stingray.pulse.search.search_best_peaks [UPSTREAM-RECALLED]: contiguous regions with
    stat >= threshold -> arg-max of each, sorted by stat descending."""
    x = np.asarray(x)
    stat = np.asarray(stat)
    peaks = stat >= threshold
    best_x, best_stat = [], []
    i = 0
    n = stat.size
    while i < n:
        if not peaks[i]:
            i += 1
            continue
        j = i
        while j < n and peaks[j]:
            j += 1
        m = i + int(np.argmax(stat[i:j]))
        best_x.append(x[m])
        best_stat.append(stat[m])
        i = j
    order = np.argsort(best_stat)[::-1]
    return [best_x[o] for o in order], [best_stat[o] for o in order]


# --------------------------------------------------------------------------------------
# Fast folding algorithm (HENzsearch/HENefsearch --ffa): hendrics/ffa.py:210-395,
# hendrics/efsearch.py:1092-1138, hendrics/base.py:621-630 (1-D histogram).  In-tree code, so this
# section is PINNED: tests/golden/ffa_*.npz come from the reference's own functions.
# --------------------------------------------------------------------------------------
@njit()
def _hist1d_seq(H, tracks, bins, lo, hi):
    """hendrics/base.py:621-630 (right edge excluded, truncating index)."""
    delta = 1 / ((hi - lo) / bins)
    for t in range(tracks.size):
        i = (tracks[t] - lo) * delta
        if 0 <= i < bins:
            H[int(i)] += 1
    return H


def histogram(a, bins, range):
    """hendrics/base.py:633-660, 947-972 (unweighted numba variant; counts come back in a.dtype)."""
    a = np.asarray(a, dtype=np.float64)
    return _hist1d_seq(np.zeros(int(bins), dtype=np.float64), a, int(bins), float(range[0]), float(range[1]))


@njit()
def ffa_step_pow(step):
    """hendrics/ffa.py:234-236."""
    return 2 ** (step + 1)


@njit()
def ffa_shift(prof_n, step):
    """hendrics/ffa.py:239-241."""
    return (prof_n % ffa_step_pow(step) + 1) // 2


@njit()
def ffa_start_value(prof_n, step):
    """hendrics/ffa.py:244-267."""
    val = 0
    sp = ffa_step_pow(step)
    val += prof_n // sp * sp
    if step >= 1:
        val += (prof_n - val) // 2
    return val


@njit()
def ffa_step(array, step, ntables):
    """hendrics/ffa.py:291-308: table prof_n of the next level = table `start` + table `start + 2**step` rolled by
    -shift (roll(a, -s)[j] == a[(j + s) % n], ffa.py:215-231)."""
    out = np.copy(array)
    jump = 2**step
    nb = array.shape[1]
    for prof_n in range(ntables):
        start = ffa_start_value(prof_n, step)
        sh = ffa_shift(prof_n, step) % nb
        for j in range(nb):
            jj = j + sh
            if jj >= nb:
                jj -= nb
            out[prof_n, j] = array[start, j] + array[start + jump, jj]
    return out


def _ffa(array_reshaped, bin_period, ntables, z_n_n=2):
    """hendrics/ffa.py:311-334."""
    periods = np.array([bin_period + n / (ntables - 1) for n in range(ntables)])
    for step in range(int(np.log2(ntables))):
        array_reshaped = ffa_step(array_reshaped, step, ntables)
    # the reference builds its tables with np.arange(0, 1, 1 / nbin), whose length is nbin or nbin + 1 depending on
    # the rounding of 1 / nbin; the tiling below (and hence harmonic k >= 2) inherits that -- kept as is.
    twopiphases = np.pi * 2 * np.arange(0, 1, 1 / array_reshaped.shape[1])
    cached_cos = np.tile(np.cos(twopiphases), z_n_n)
    cached_sin = np.tile(np.sin(twopiphases), z_n_n)
    stats = np.zeros(ntables)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(array_reshaped.shape[0]):
            # argument order as at ffa.py:332 (cos and sin swapped w.r.t. the signature; symmetric)
            stats[i] = _z_n_fast_cached(array_reshaped[i, :], cached_cos, cached_sin, z_n_n)
    return periods, stats


def ffa_tables(bin_period, z_n_n=2):
    """The (cos, sin) values harmonic k = 1..z_n_n of `_ffa` multiplies bin j with: tab[0, k-1, j], tab[1, k-1, j]
    (stride-k slices of the tiled tables, ffa.py:98-102, 320-328)."""
    twopiphases = np.pi * 2 * np.arange(0, 1, 1 / bin_period)
    cc = np.tile(np.cos(twopiphases), z_n_n)
    ss = np.tile(np.sin(twopiphases), z_n_n)
    N = bin_period
    tab = np.zeros((2, z_n_n, N))
    for k in range(1, z_n_n + 1):
        tab[0, k - 1] = cc[: N * k : k]
        tab[1, k - 1] = ss[: N * k : k]
    return tab


def ffa(array, bin_period, z_n_n=2):
    """hendrics/ffa.py:337-349."""
    N_raw = len(array)
    ntables = int(2 ** np.ceil(np.log2(N_raw // bin_period + 1)))
    if ntables <= 1:
        return np.zeros(1), np.zeros(1)
    N = ntables * bin_period
    new_arr = np.zeros(N)
    new_arr[:N_raw] = array
    return _ffa(new_arr.reshape([ntables, bin_period]), bin_period, ntables, z_n_n=z_n_n)


def _quick_rebin(counts, current_rebin):
    """hendrics/ffa.py:352-363."""
    n = int(counts.size // current_rebin)
    return np.sum(counts[: n * current_rebin].reshape(n, current_rebin), axis=1)


def ffa_search(counts, dt, period_min, period_max):
    """hendrics/ffa.py:366-395 (always Z^2_2: `ffa` is called without z_n_n)."""
    counts = np.array(counts)
    pmin = np.floor(period_min / dt)
    pmax = np.ceil(period_max / dt)
    bin_periods = None
    stats = None
    current_rebin = 1
    rebinned_counts = counts
    for bin_period in np.arange(pmin, pmax + 1, dtype=int):
        rebin = int(2 ** np.floor(np.log2(bin_period / pmin)))
        if rebin > current_rebin:
            current_rebin = rebin
            rebinned_counts = _quick_rebin(counts, current_rebin)
        if bin_period % current_rebin != 0:
            continue
        per, st = ffa(rebinned_counts, bin_period // current_rebin)
        per *= current_rebin
        if per[0] == 0:
            continue
        elif bin_periods is None:
            bin_periods = per[:-1] * dt
            stats = st[:-1]
        else:
            bin_periods = np.concatenate((bin_periods, per[:-1] * dt))
            stats = np.concatenate((stats, st[:-1]))
    return bin_periods, stats


def search_with_ffa(times, f0, f1, nbin=16, n=1, t0=None, t1=None):
    """hendrics/efsearch.py:1092-1138 (`n` is accepted and ignored, as in the reference)."""
    times = np.asarray(times)
    if t0 is None:
        t0 = times[0]
    if t1 is None:
        t1 = times[-1]
    length = (t1 - t0).astype(np.double)
    p0 = 1 / f1
    p1 = 1 / f0
    dt = p0 / nbin
    counts = histogram((times - t0).astype(np.double), range=[0, length], bins=int(np.rint(length / dt)))
    bin_periods, stats = ffa_search(counts, dt, p0, p1)
    return 1 / bin_periods, stats, None, length


# --------------------------------------------------------------------------------------
# long-double reference for the "more accurate than the reference" fixed-point mode
# --------------------------------------------------------------------------------------
def exact_bins(ts, f, fd, nbin):
    """floor(frac(t*f + 0.5*t^2*fd) * nbin) evaluated in exact rational arithmetic on the
    fp64 inputs (Python ints via float.as_integer_ratio); used for bin-flip counts."""
    from fractions import Fraction

    F = Fraction(float(f))
    FD = Fraction(float(fd))
    out = np.empty(len(ts), dtype=np.int64)
    for i, t in enumerate(ts):
        T = Fraction(float(t))
        ph = T * F + Fraction(1, 2) * T * T * FD
        fr = ph - math.floor(ph)
        out[i] = int(math.floor(fr * nbin))
    return out


# --------------------------------------------------------------------------------------
# synthetic HENfake-style event lists (benchmark inputs) -- shared by tests/bench
# --------------------------------------------------------------------------------------
def simulate_events(n_events, tstop, f0=1 / 0.101, amp=0.25, seed=1, tstart=0.0):
    """N sorted fp64 arrival times on [tstart, tstop) from rate ~ 1 + amp*cos(2 pi f0 t),
    by thinning a uniform sample (recipe of tests/test_efsearch.py:46-62 restated)."""
    rng = np.random.default_rng(seed)
    out = np.empty(0)
    while out.size < n_events:
        m = int((n_events - out.size) * (1 + amp) * 1.05) + 16
        t = rng.uniform(tstart, tstop, m)
        keep = rng.uniform(0, 1 + amp, m) < 1 + amp * np.cos(2 * np.pi * f0 * t)
        out = np.concatenate([out, t[keep]])
    out = out[:n_events]
    out.sort()
    return out
