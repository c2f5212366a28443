"""Drop-in replacements for ``stingray.pulse.search.z_n_search`` / ``epoch_folding_search``.

These are the two module-level names HENDRICS imports at ``hendrics/efsearch.py:15-18`` and
passes as ``func=`` to ``folding_search`` (``hendrics/efsearch.py:1141-1193``).  Signatures,
keyword names, return shapes and error behaviour follow the upstream functions (SURVEY.md
section 8b / Appendix A); the per-trial fold + statistic runs in the sm_100a kernel K1
(``csrc/hx_fold.cu``) through the C ABI (``include/hx_search.h``).

Host code here is plumbing only: event selection per segment (half-open mask, like upstream),
device buffers (torch), and the call into ``hx_fold_search_dev``.  No CPU fallback exists.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np

from . import _cabi

__all__ = ["z_n_search", "epoch_folding_search", "fold_profiles", "fold_profiles_weighted", "orbital_fold_stats",
           "z_n_search_unbinned", "windowed_searches", "fold_windows_device",
           "UPSTREAM_SEMANTICS"]

# The two places where the restatement of stingray's `_folding_search` could not be checked against
# its source.  A maintainer with stingray at hand can flip
# them without touching any kernel; the oracle has the same switches.
UPSTREAM_SEMANTICS = {
    "rezero_to_first_event": True,  # A1: times -= times[0] before folding
    "drop_last_event": True,        # A2: segment mask is [s, s + segment_size): t == t[-1] is not folded
}


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise _cabi.HxError(_cabi.HX_ERR_CUDA, "no CUDA device available; hendrics_b200 has no CPU fallback")
    return torch


def _device(device):
    torch = _torch()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _aligned(t):
    """K1 streams events with 16-byte bulk copies: the first element must be 16-byte aligned."""
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def fold_stats_device(ts_d, freqs_d, fdots_d, nbin, nharm, flags=0, want_profiles=False):
    """One K1 call on device tensors.  Returns stats (n_fd, n_f) [and uint32-valued profiles
    (n_fd, n_f, nbin) as an int32 tensor]."""
    torch = _torch()
    n_f, n_fd = freqs_d.numel(), fdots_d.numel()
    stats = torch.empty((n_fd, n_f), dtype=torch.float64, device=ts_d.device)
    prof = torch.empty((n_fd, n_f, nbin), dtype=torch.int32, device=ts_d.device) if want_profiles else None
    ts_d = _aligned(ts_d.contiguous())
    with torch.cuda.device(ts_d.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = _cabi.lib().hx_fold_search_dev(
            ts_d.data_ptr(), ts_d.numel(), freqs_d.data_ptr(), n_f, fdots_d.data_ptr(), n_fd, int(nbin), int(nharm),
            int(flags), stats.data_ptr(), prof.data_ptr() if want_profiles else None, stream,
        )
    _cabi.check(rc)
    return (stats, prof) if want_profiles else stats


# A window (segment / dynamical-search step) this small cannot fill the GPU by itself: all windows go through ONE
# hx_fold_windows_dev launch.  Above it, one K1 launch per window is the faster plan (K1 is ~10x quicker per unit).
WINDOWED_MAX_UNITS_PER_WINDOW = 1.0e8

WIN_SEGMENTS, WIN_SEARCHES = 0, 1


def fold_windows_device(ts_d, win_start, win_stop, mode, freqs_d, fdots_d, nbin, nharm, origin=0.0, segment_size=0.0,
                        want_sum=False):
    """One ``hx_fold_windows_dev`` call on a sorted device event list.  Returns ``(stats (n_win, n_fd, n_f) tensor,
    status int32[n_win] (1 folded, 0 skipped, 2 upstream raises), sum (n_fd, n_f) tensor or None)``."""
    import ctypes

    torch = _torch()
    win_start = np.ascontiguousarray(win_start, dtype=np.float64)
    win_stop = np.ascontiguousarray(win_stop, dtype=np.float64)
    n_win, n_f, n_fd = win_start.size, freqs_d.numel(), fdots_d.numel()
    stats = torch.empty((n_win, n_fd, n_f), dtype=torch.float64, device=ts_d.device)
    total = torch.empty((n_fd, n_f), dtype=torch.float64, device=ts_d.device) if want_sum else None
    status = np.zeros(n_win, dtype=np.int32)
    ts_d = ts_d.contiguous()
    with torch.cuda.device(ts_d.device):
        rc = _cabi.lib().hx_fold_windows_dev(
            ts_d.data_ptr(), ts_d.numel(), _cabi.dptr(win_start), _cabi.dptr(win_stop), n_win, int(mode), float(origin),
            float(segment_size), freqs_d.data_ptr(), n_f, fdots_d.data_ptr(), n_fd, int(nbin), int(nharm),
            stats.data_ptr(), total.data_ptr() if want_sum else None, status.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
            torch.cuda.current_stream().cuda_stream,
        )
    _cabi.check(rc)
    return stats, status, total


def _segment_slices(times_d, start_times, segment_size, half_open):
    """``[(lo, hi, t[lo], t[hi - 1]), ...]``: the events with ``s <= t < s + segment_size`` (``<=`` when not half-open) of
    every segment start ``s`` as index ranges of the SORTED device array; None when the array is not sorted."""
    torch = _torch()
    if not _is_sorted_device(times_d):
        return None
    starts = torch.as_tensor(np.asarray(start_times, dtype=np.float64), device=times_d.device)
    lo = torch.searchsorted(times_d, starts, right=False)
    hi = torch.searchsorted(times_d, starts + float(segment_size), right=not half_open)
    n = times_d.numel()
    edges = torch.stack([times_d[lo.clamp(max=n - 1)], times_d[(hi - 1).clamp(min=0)]])
    lo, hi, edges = lo.cpu().numpy(), hi.cpu().numpy(), edges.cpu().numpy()
    return [(int(a), int(b), float(x), float(y)) for a, b, x, y in zip(lo, hi, edges[0], edges[1])]


def _is_sorted_device(t_d):
    return t_d.numel() < 2 or bool((t_d[1:] >= t_d[:-1]).all().item())


def windowed_searches(nharm, times, frequencies, starts, stops, nbin=128, fdots=0, device=None):
    """``[func(times[(times >= st) & (times < sp)], frequencies, ...) for st, sp in zip(starts, stops)]`` -- the loop of
    ``dyn_folding_search`` (hendrics/efsearch.py:1221-1231) -- in one launch, for sorted ``times``.  Returns
    ``(stats (n_windows, n_fdot, n_freq) numpy, status int32[n_windows])``; rows with status 0 are NaN (upstream's 0/0),
    rows with status 2 are zeros (upstream raises; the caller stores zeros).  None when ``times`` is not sorted."""
    from .dist import _to_host

    torch = _torch()
    dev = _device(device)
    times = np.asanyarray(times)
    if times.dtype != np.float64:
        return None  # extended-precision times are re-zeroed on the host per window: keep the per-window calls
    frequencies = np.atleast_1d(np.asanyarray(frequencies).astype(np.float64))
    fdots = np.atleast_1d(np.asanyarray(fdots).astype(np.float64))
    ts_d = torch.from_numpy(np.ascontiguousarray(times)).to(dev)
    if not _is_sorted_device(ts_d):
        return None
    f_d = torch.from_numpy(np.ascontiguousarray(frequencies)).to(dev)
    fd_d = torch.from_numpy(np.ascontiguousarray(fdots)).to(dev)
    stats, status, _ = fold_windows_device(ts_d, starts, stops, WIN_SEARCHES, f_d, fd_d, nbin, nharm)
    return _to_host(stats), status


def _folding_search(
    nharm,
    times,
    frequencies,
    nbin=128,
    segment_size=np.inf,
    fdots=0,
    device=None,
    flags=0,
    return_profiles=False,
    exact_phase=False,
    _device_stats=False,
):
    """Upstream ``_folding_search`` fast branch with the (fdot, f) double loop
    replaced by one K1 launch per segment.

    ``exact_phase=True`` (extension, NOT the parity path) evaluates the phase mod 1 from error-free
    products instead of the reference's fp64 phase: for very long baselines (|t f| >~ 1e10 cycles)
    the reference itself warns that fp64 is not enough (hendrics/efsearch.py:199-205, 993-1000)."""
    if exact_phase:
        flags = int(flags) | _cabi.HX_FOLD_EXACT_PHASE
    torch = _torch()
    dev = _device(device)
    frequencies = np.atleast_1d(np.asanyarray(frequencies).astype(np.float64))
    fdots = np.atleast_1d(np.asanyarray(fdots).astype(np.float64))
    grid_shape = (fdots.size, frequencies.size)
    rezero = UPSTREAM_SEMANTICS["rezero_to_first_event"]
    if isinstance(times, torch.Tensor):
        # already on the device (dist.upload_replicated: sharded H2D + NVLink all-gather of the replicated event list)
        if times.dtype != torch.float64 or times.dim() != 1:
            raise TypeError("device event lists must be 1-D float64 tensors")
        if times.numel() == 0:
            raise ValueError("No events to fold")
        times_d = times.to(dev)
        ends = times_d[[0, -1]].cpu().numpy()
        t_first = np.float64(ends[0] - ends[0]) if rezero else np.float64(ends[0])
        length = np.float64(ends[1] - ends[0]) if rezero else np.float64(ends[1])
        if rezero:
            times_d = times_d - float(ends[0])
        times = None
    else:
        times = np.asanyarray(times)
        if times.size == 0:
            raise ValueError("No events to fold")
    if times is None:
        pass
    elif times.dtype == np.float64:
        # A1: re-zero to the first event.  The fp64 subtraction is done on the device (same IEEE
        # operation as numpy's) to spare two host passes over the event list.
        t_first = np.float64(times[0] - times[0]) if rezero else np.float64(times[0])
        length = np.float64(times[-1] - times[0]) if rezero else np.float64(times[-1])
        times_d = torch.from_numpy(np.ascontiguousarray(times)).to(dev)
        if rezero:
            times_d = times_d - float(times[0])
    else:  # e.g. long double mission times: subtract in the input precision, then cast (efsearch.py:1154)
        times = ((times - times[0]) if rezero else times).astype(np.float64)
        t_first, length = times[0], times[-1]
        times_d = torch.from_numpy(np.ascontiguousarray(times)).to(dev)
    if length < segment_size:
        segment_size = length
    start_times = np.arange(t_first, length, segment_size)

    freqs_d = torch.from_numpy(np.ascontiguousarray(frequencies)).to(dev)
    fdots_d = torch.from_numpy(np.ascontiguousarray(fdots)).to(dev)
    stats_d = torch.zeros(grid_shape, dtype=torch.float64, device=dev)
    prof_d = torch.zeros(grid_shape + (nbin,), dtype=torch.int64, device=dev) if return_profiles else None
    count = 0
    n_seg = len(start_times)
    if (n_seg > 1 and n_seg <= 65535 and not return_profiles and flags == 0 and UPSTREAM_SEMANTICS["drop_last_event"]
            and times_d.numel() / n_seg * grid_shape[0] * grid_shape[1] <= WINDOWED_MAX_UNITS_PER_WINDOW
            and n_seg * grid_shape[0] * grid_shape[1] * 8 <= (1 << 30) and _is_sorted_device(times_d)):
        # many small segments (--segment-size): one launch for all of them, summed in segment order on the device
        _, status, total = fold_windows_device(times_d, start_times, start_times + segment_size, WIN_SEGMENTS, freqs_d,
                                               fdots_d, nbin, nharm, origin=0.0, segment_size=segment_size, want_sum=True)
        return _grids_and_stats(frequencies, fdots, total, int(np.count_nonzero(status == 1)), device_stats=_device_stats)
    # Sorted times (the usual case): a segment is a SLICE of the device array -- its bounds come from one searchsorted
    # for all segments and one small D2H copy -- instead of a boolean mask + compaction over the whole list per segment.
    ranges = _segment_slices(times_d, start_times, segment_size, UPSTREAM_SEMANTICS["drop_last_event"])
    for k, s in enumerate(start_times):
        if ranges is not None:
            lo, hi, t_lo, t_hi = ranges[k]
            if hi <= lo:
                continue
            ts_d, span = times_d[lo:hi], t_hi - t_lo
        else:
            if UPSTREAM_SEMANTICS["drop_last_event"]:
                good = (times_d >= s) & (times_d < s + segment_size)  # A2: half-open
            else:
                good = (times_d >= s) & (times_d <= s + segment_size)
            ts_d = times_d[good]
            if ts_d.numel() < 1:
                continue
            span = float(ts_d[-1] - ts_d[0])
        if span < 0.2 * segment_size:
            continue
        res = fold_stats_device(ts_d, freqs_d, fdots_d, nbin, nharm, flags=flags, want_profiles=return_profiles)
        if return_profiles:
            stats_d += res[0]
            prof_d += res[1].to(torch.int64) & 0xFFFFFFFF
        else:
            stats_d += res
        count += 1
    out = _grids_and_stats(frequencies, fdots, stats_d, count, device_stats=_device_stats)
    if return_profiles:
        out = out + (prof_d.cpu().numpy(),)
    return out


def _grids_and_stats(frequencies, fdots, stats_d, count, device_stats=False):
    """Return tuple of upstream's ``_folding_search``: ``(fgrid.flatten(), stats.flatten() / count)`` for a single
    fdot, else ``(fgrid, fdgrid, stats / count)`` with shape (n_fdot, n_freq).  The division (same IEEE operation as
    numpy's) runs on the device and the 1-D case skips the meshgrid: no extra host passes over 10^6-trial arrays.
    ``device_stats`` (multi-GPU path, ``dist.sharded_search``): the statistics stay a device tensor -- they go
    straight into the NCCL all-gather -- and no grids are built for the slice."""
    from .dist import _to_host

    torch = _torch()
    stats_d = stats_d / count if count > 0 else torch.full_like(stats_d, float("nan"))  # upstream: 0 / 0
    if device_stats:
        return (frequencies, stats_d) if fdots.size == 1 else (frequencies, fdots, stats_d)
    stats = _to_host(stats_d)
    if fdots.size == 1:
        return frequencies, stats.reshape(-1)  # `frequencies` is already this call's own astype() copy
    fgrid, fdgrid = np.meshgrid(frequencies, fdots)
    return fgrid, fdgrid, stats


def _is_slow_branch(expocorr, weights):
    """Upstream takes the `fold_events` branch when the exposure correction is asked for or the weights
    are an array (``expocorr or not HAS_NUMBA or isinstance(weights, Iterable)``)."""
    return bool(expocorr) or isinstance(weights, Iterable)


def wfold_stats_device(ts_d, w_d, unit_weight, freqs_d, fdots_d, nbin, nharm, gti=None, start_time=0.0,
                       stop_time=0.0, want_profiles=False):
    """One K1w call on device tensors (``hx_wfold_search_dev``).  ``w_d`` None: every event weighs
    ``unit_weight``.  ``gti`` (host, (n, 2)) switches the exposure correction on."""
    torch = _torch()
    n_f, n_fd = freqs_d.numel(), fdots_d.numel()
    stats = torch.empty((n_fd, n_f), dtype=torch.float64, device=ts_d.device)
    prof = torch.empty((n_fd, n_f, nbin), dtype=torch.float64, device=ts_d.device) if want_profiles else None
    ts_d = _aligned(ts_d.contiguous())
    if w_d is not None:
        w_d = _aligned(w_d.contiguous())
    n_gti = 0
    gti_p = None
    if gti is not None:
        gti = np.ascontiguousarray(gti, dtype=np.float64).reshape(-1, 2)
        n_gti, gti_p = gti.shape[0], _cabi.dptr(gti)
    with torch.cuda.device(ts_d.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = _cabi.lib().hx_wfold_search_dev(
            ts_d.data_ptr(), w_d.data_ptr() if w_d is not None else None, float(unit_weight), ts_d.numel(),
            freqs_d.data_ptr(), n_f, fdots_d.data_ptr(), n_fd, int(nbin), int(nharm), gti_p, n_gti,
            float(start_time), float(stop_time), stats.data_ptr(), prof.data_ptr() if want_profiles else None, stream,
        )
    _cabi.check(rc)
    return (stats, prof) if want_profiles else stats


def _folding_search_weighted(
    nharm, times, frequencies, nbin=128, segment_size=np.inf, fdots=0, expocorr=False, gti=None, weights=1,
    device=None, return_profiles=False, _device_stats=False,
):
    """Upstream ``_folding_search(..., use_times=True)`` with the per-trial
    ``fold_events`` + statistic replaced by one K1w launch per segment.  As upstream: times are re-zeroed to
    the first event (A1), array keywords as long as ``times`` are masked like the events, the GTIs are handed
    to ``fold_events`` as they come (NOT re-zeroed with the times), ``start_time = ts[0]`` and
    ``stop_time = ts[-1] + |ts[1] - ts[0]|`` bound the exposure."""
    torch = _torch()
    dev = _device(device)
    frequencies = np.atleast_1d(np.asanyarray(frequencies).astype(np.float64))
    fdots = np.atleast_1d(np.asanyarray(fdots).astype(np.float64))
    grid_shape = (fdots.size, frequencies.size)
    times = np.asanyarray(times)
    if times.size == 0:
        raise ValueError("No events to fold")
    n_all = times.size
    if UPSTREAM_SEMANTICS["rezero_to_first_event"]:
        times = (times - times[0]).astype(np.float64)
    else:
        times = times.astype(np.float64)
    length = times[-1]
    if length < segment_size:
        segment_size = length
    start_times = np.arange(times[0], times[-1], segment_size)

    times_d = torch.from_numpy(np.ascontiguousarray(times)).to(dev)
    w_all = None
    unit_weight = 1.0
    if isinstance(weights, Iterable):
        weights = np.asanyarray(weights)
        if weights.shape != (n_all,):
            raise ValueError("weights must be a scalar or have one entry per event")
        w_all = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)
    else:
        unit_weight = float(weights)
    gti_all = None
    if expocorr:
        gti_all = np.asanyarray(gti).astype(np.float64)
    freqs_d = torch.from_numpy(np.ascontiguousarray(frequencies)).to(dev)
    fdots_d = torch.from_numpy(np.ascontiguousarray(fdots)).to(dev)
    stats_d = torch.zeros(grid_shape, dtype=torch.float64, device=dev)
    prof_d = torch.zeros(grid_shape + (nbin,), dtype=torch.float64, device=dev) if return_profiles else None
    count = 0
    for s in start_times:
        if UPSTREAM_SEMANTICS["drop_last_event"]:
            good = (times_d >= s) & (times_d < s + segment_size)
        else:
            good = (times_d >= s) & (times_d <= s + segment_size)
        ts_d = times_d[good]
        n_good = ts_d.numel()
        if n_good < 1:
            continue
        edge = ts_d[[0, 1, -1] if n_good > 1 else [0, 0, -1]].cpu().numpy()
        if edge[2] - edge[0] < 0.2 * segment_size:
            continue
        if n_good < 2:
            raise IndexError("fold_events needs at least two events per segment")  # upstream: times[1]
        g = gti_all
        if g is not None and len(g) == n_all:  # upstream masks every keyword as long as `times`
            g = g[good.cpu().numpy()]
        res = wfold_stats_device(
            ts_d, w_all[good] if w_all is not None else None, unit_weight, freqs_d, fdots_d, nbin, nharm, gti=g,
            start_time=edge[0], stop_time=edge[2] + np.abs(edge[1] - edge[0]), want_profiles=return_profiles,
        )
        if return_profiles:
            stats_d += res[0]
            prof_d += res[1]
        else:
            stats_d += res
        count += 1
    out = _grids_and_stats(frequencies, fdots, stats_d, count, device_stats=_device_stats)
    if return_profiles:
        out = out + (prof_d.cpu().numpy(),)
    return out


def epoch_folding_search(
    times, frequencies, nbin=128, segment_size=np.inf, expocorr=False, gti=None, weights=1, fdots=0, **kwargs
):
    """Epoch-folding chi^2 periodogram.  Same call as ``stingray.pulse.search.epoch_folding_search``.

    Returns ``(fgrid, stats)`` (1-D) when a single fdot is searched, else
    ``(fgrid, fdgrid, stats)`` of shape ``(n_fdot, n_freq)``.  ``expocorr`` / array ``weights`` select the
    weighted kernel K1w, like upstream's ``fold_events`` branch."""
    if expocorr and gti is None:
        raise ValueError("To calculate exposure correction, you need to specify the GTIs")
    if _is_slow_branch(expocorr, weights):
        return _folding_search_weighted(0, times, frequencies, nbin=nbin, segment_size=segment_size, fdots=fdots,
                                        expocorr=expocorr, gti=gti, weights=weights, **kwargs)
    return _folding_search(0, times, frequencies, nbin=nbin, segment_size=segment_size, fdots=fdots, **kwargs)


def z_n_search(
    times,
    frequencies,
    nharm=4,
    nbin=128,
    segment_size=np.inf,
    expocorr=False,
    weights=1,
    gti=None,
    fdots=0,
    **kwargs,
):
    """Binned Z^2_n periodogram.  Same call as ``stingray.pulse.search.z_n_search``."""
    if expocorr and gti is None:
        raise ValueError("To calculate exposure correction, you need to specify the GTIs")
    if nharm < 1:
        raise ValueError("nharm must be >= 1")
    if _is_slow_branch(expocorr, weights):
        return _folding_search_weighted(nharm, times, frequencies, nbin=nbin, segment_size=segment_size, fdots=fdots,
                                        expocorr=expocorr, gti=gti, weights=weights, **kwargs)
    return _folding_search(nharm, times, frequencies, nbin=nbin, segment_size=segment_size, fdots=fdots, **kwargs)


# dist.sharded_search may ask these for their statistics as a device tensor (`_device_stats=True`)
epoch_folding_search._hx_device_stats = True
z_n_search._hx_device_stats = True


def fold_profiles(times, frequencies, fdots=0, nbin=128, nharm=0, device=None, flags=0, exact_phase=False):
    """Profiles AND statistics of one fold over already-selected, already-zeroed times.

    Returns ``(stats (n_fd, n_f), profiles (n_fd, n_f, nbin) int64)``; used by the parity tests
    for the bit-exact profile-count check."""
    torch = _torch()
    dev = _device(device)
    if exact_phase:
        flags = int(flags) | _cabi.HX_FOLD_EXACT_PHASE
    ts_d = torch.from_numpy(np.ascontiguousarray(times, dtype=np.float64)).to(dev)
    f_d = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(frequencies), dtype=np.float64)).to(dev)
    fd_d = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(fdots), dtype=np.float64)).to(dev)
    stats, prof = fold_stats_device(ts_d, f_d, fd_d, nbin, nharm, flags=flags, want_profiles=True)
    return stats.cpu().numpy(), (prof.to(torch.int64) & 0xFFFFFFFF).cpu().numpy()


def fold_profiles_weighted(times, frequencies, fdots=0, nbin=128, nharm=0, weights=1, gti=None, device=None):
    """Profiles AND statistics of one weighted (K1w) fold over already-selected times, i.e. upstream's
    ``fold_events(times, f, fd, nbin=, weights=, gti=, expocorr=gti is not None)`` for every trial.

    Returns ``(stats (n_fd, n_f), profiles (n_fd, n_f, nbin) float64)``; used by the parity tests."""
    torch = _torch()
    dev = _device(device)
    times = np.ascontiguousarray(times, dtype=np.float64)
    ts_d = torch.from_numpy(times).to(dev)
    f_d = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(frequencies), dtype=np.float64)).to(dev)
    fd_d = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(fdots), dtype=np.float64)).to(dev)
    w_d, unit = None, 1.0
    if isinstance(weights, Iterable):
        w_d = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)
    else:
        unit = float(weights)
    stats, prof = wfold_stats_device(
        ts_d, w_d, unit, f_d, fd_d, nbin, nharm, gti=gti, start_time=times[0],
        stop_time=times[-1] + np.abs(times[1] - times[0]), want_profiles=True,
    )
    return stats.cpu().numpy(), prof.cpu().numpy()


def orbital_fold_stats(times, freq, fdot, X, Porb, T0s, nbin=128, nharm=0, device=None, return_profiles=False):
    """Statistic of the single trial ``(freq, fdot)`` on the event times demodulated for every trial epoch of
    ascending node in ``T0s`` (``hx_orbital_search_dev``): the T0 loop of hendrics/efsearch.py:294-302 in one launch.
    ``nharm == 0``: epoch-folding chi^2, else Z^2_nharm.  Returns ``stats[len(T0s)]`` (and the uint32 profiles)."""
    torch = _torch()
    dev = _device(device)
    ts_d = _aligned(torch.from_numpy(np.ascontiguousarray(times, dtype=np.float64)).to(dev))
    t0_d = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(T0s), dtype=np.float64)).to(dev)
    stats = torch.empty(t0_d.numel(), dtype=torch.float64, device=dev)
    prof = torch.empty((t0_d.numel(), nbin), dtype=torch.int32, device=dev) if return_profiles else None
    with torch.cuda.device(dev):
        rc = _cabi.lib().hx_orbital_search_dev(
            ts_d.data_ptr(), ts_d.numel(), float(freq), float(fdot), float(X), float(Porb), t0_d.data_ptr(),
            t0_d.numel(), int(nbin), int(nharm), stats.data_ptr(), prof.data_ptr() if return_profiles else None,
            torch.cuda.current_stream().cuda_stream,
        )
    _cabi.check(rc)
    if return_profiles:
        return stats.cpu().numpy(), (prof.to(torch.int64) & 0xFFFFFFFF).cpu().numpy()
    return stats.cpu().numpy()


def z_n_search_unbinned(times, frequencies, nharm=2, fdots=0, device=None):
    """Unbinned Z^2_n (extension, K4): oracle formula ``hendrics/efsearch.py:366-404`` with unit
    weights.  Same return convention as ``z_n_search``."""
    torch = _torch()
    dev = _device(device)
    frequencies = np.atleast_1d(np.asanyarray(frequencies).astype(np.float64))
    fdots = np.atleast_1d(np.asanyarray(fdots).astype(np.float64))
    fgrid, fdgrid = np.meshgrid(frequencies, fdots)
    ts_d = torch.from_numpy(np.ascontiguousarray(times, dtype=np.float64)).to(dev)
    f_d = torch.from_numpy(frequencies).to(dev)
    fd_d = torch.from_numpy(fdots).to(dev)
    stats = torch.empty(fgrid.shape, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = _cabi.lib().hx_zn_unbinned_dev(
            ts_d.data_ptr(), ts_d.numel(), f_d.data_ptr(), f_d.numel(), fd_d.data_ptr(), fd_d.numel(), int(nharm),
            stats.data_ptr(), torch.cuda.current_stream().cuda_stream,
        )
    _cabi.check(rc)
    stats = stats.cpu().numpy()
    if fgrid.shape[0] == 1:
        return fgrid.flatten(), stats.flatten()
    return fgrid, fdgrid, stats
