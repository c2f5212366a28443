"""Host side of the pre-binned searches: ``--fast`` (quasi-fast folding) and the transient scan.

    search_with_qffa_step   hendrics/efsearch.py:876-925
    search_with_qffa        hendrics/efsearch.py:928-1089
    transient_search        hendrics/efsearch.py:499-650  (+ TransientResults :488-496)

The reference loops over centre frequencies in Python and calls numba kernels per centre.  Here the per-centre work is
the CUDA path (K2 + K3, ``hx_qffa_search_dev`` / ``hx_transient_search_dev``) for ALL centres of a rank at once; this
module only prepares what the kernels need (centred fp64 times, the ladder of centre frequencies, the (L, Q) bin-shift
plane) and lays the statistics out in the reference's result arrays.  The numbers that define the output grids -- ladder
step ``4 npfact / length`` (``1 / length / oversample`` for the transient scan), the ``+0.12 step`` offset, ``rint(...) +
2`` centres, ``linspace(-nbin npfact, nbin npfact, oversample npfact)`` shifts, bin -> frequency / fdot factors -- are
the interface and are kept to the letter (``tests/golden/qffa_full*.npz``, ``transient_*.npz``).
"""

from __future__ import annotations

import logging
import tempfile
import warnings
from dataclasses import dataclass

import numpy as np

from . import _cabi

log = logging.getLogger("hendrics_b200")

MEMMAP_THRESHOLD = 1e7  # result elements above which the reference switches to memory-mapped arrays


def check_phase_error_after_casting_to_double(tref, f, fdot=0):
    """Largest phase difference (cycles) between an extended-precision and an fp64 evaluation of ``t f + t^2 fdot / 2``
    for 1000 times scattered by 0.1 s around ``tref`` (hendrics/efsearch.py:199-205; draws from numpy's global RNG like
    the reference)."""
    t_ext = np.array(np.random.normal(tref, 0.1, 1000), dtype=np.longdouble)
    t_dbl = t_ext.astype(np.double)
    exact = t_ext * f + 0.5 * t_ext**2 * fdot
    rounded = t_dbl * np.double(f) + 0.5 * t_dbl**2 * np.double(fdot)
    return np.max(np.abs(rounded - exact))


# ---------------------------------------------------------------------------------------------
# ingredients
# ---------------------------------------------------------------------------------------------
@dataclass
class CentredTimes:
    """Event times re-centred on the middle of [t0, t1] (in the caller's precision) and then cast to fp64."""

    times: np.ndarray
    length: float

    @classmethod
    def prepare(cls, times, t0, t1, f_top, fdot, nbin, announce):
        times = np.array(times)  # a copy: the caller's array is never modified
        lo = times.min() if t0 is None else t0
        hi = times.max() if t1 is None else t1
        times -= (hi + lo) / 2
        err = check_phase_error_after_casting_to_double(np.max(times), f_top, fdot)
        announce(err, err > 1 / nbin / 10)
        return cls(times.astype(np.double), hi - lo)

    @property
    def first(self):
        return self.times[0]

    @property
    def last(self):
        return self.times[-1]

    def on_device(self):
        """The searches that use this are collective when torch.distributed is initialised (centres sharded over the
        ranks): the replicated event list goes up in shards + one NVLink all-gather (dist.upload_replicated)."""
        from .dist import upload_replicated

        return upload_replicated(self.times)


def centre_ladder(f0, f1, step):
    """Centre frequencies ``(f0 + f1) / 2 + (i + 0.12) step`` for ``i`` in ``[-(n // 2), n // 2)``, ``n = rint((f1 - f0)
    / step) + 2`` (a single centre at i = 0 when that range is empty): hendrics/efsearch.py:1004-1032, 592-600."""
    n = int(np.rint((f1 - f0) / step)) + 2
    rungs = np.arange(-(n // 2), n // 2) if n // 2 > 0 else np.array([0])
    middle = (f0 + f1) / 2
    return np.array([np.double(middle + step * i + 0.12 * step) for i in rungs.tolist()])


def qffa_centres(f0, f1, length, npfact=2, fdot=0, fddot=0):
    """``([(mean_f, mean_fdot, mean_fddot), ...], step)`` of a ``--fast`` search."""
    step = 4 * npfact / length
    return [(f, np.double(fdot + 0), np.double(fddot)) for f in centre_ladder(f0, f1, step)], step


class ShiftPlane:
    """The (L, Q) plane of linear / quadratic bin shifts tried around every centre, and its physical scale."""

    def __init__(self, nbin, npfact, oversample, search_fdot):
        reach, n = nbin * npfact, int(oversample * npfact)
        lin = np.linspace(-reach, reach, n)
        quad = np.linspace(-reach, reach, n) if search_fdot else np.array([0])
        self.L, self.Q = np.meshgrid(lin, quad, indexing="ij")
        self.nbin = nbin

    def scale(self, t_first, t_last):
        """(Hz per bin of linear shift, Hz/s per bin of quadratic shift) for events spanning [t_first, t_last]."""
        half_span = (t_last - t_first) / 2
        dphi = 1 / self.nbin
        return dphi / half_span, 2 * dphi / half_span**2


def _shift_grids(nbin, npfact, oversample, search_fdot):
    plane = ShiftPlane(nbin, npfact, oversample, search_fdot)
    return plane.L, plane.Q


def _result_array(shape, force_memmap, fill=None):
    """In memory, or -- like the reference for large searches -- a memory-mapped temporary ``.npy`` file."""
    if np.prod(shape) > MEMMAP_THRESHOLD or force_memmap:
        log.info("Large result arrays detected, using memory-mapped files to reduce memory usage.")
        name = tempfile.NamedTemporaryFile("w+").name + "_hen.npy"
        return np.lib.format.open_memmap(name, mode="w+", dtype=np.float64, shape=tuple(int(s) for s in shape))
    return np.zeros(shape) if fill is None else np.full(shape, fill)


def _qffa_device_call(times_d, t_first, t_last, centres, nbin, nprof, L, Q, n, want_profiles=False):
    """``hx_qffa_search_dev`` for a list of (mean_f, mean_fdot, mean_fddot): statistics (n_centres,) + L.shape
    [, profiles (n_centres, nprof, nbin) int64]."""
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


def _qffa_host_call(times, centres, nbin, nprof, L, Q, n):
    """``hx_qffa_search`` straight from the HOST event array: the library uploads it itself -- in chunks, while K2 already
    folds the chunks that have arrived (fused path, large lists) -- so a single-GPU search of 1e8 events does not wait
    for its 0.8 GB copy first.  Statistics (n_centres,) + L.shape."""
    import torch

    times = np.ascontiguousarray(times, dtype=np.float64)
    cents = np.ascontiguousarray(np.asarray(centres, dtype=np.float64).reshape(-1, 3))
    mf, mfd, mfdd = (np.ascontiguousarray(cents[:, i]) for i in range(3))
    Lf, Qf = _cabi.as_f64(L.ravel()), _cabi.as_f64(Q.ravel())
    stats = np.empty((cents.shape[0], Lf.size), dtype=np.float64)
    from .search import _device

    dev = _device(None)
    with torch.cuda.device(dev):
        rc = _cabi.lib().hx_qffa_search(
            _cabi.dptr(times), times.size, _cabi.dptr(mf), _cabi.dptr(mfd), _cabi.dptr(mfdd), cents.shape[0], int(nbin),
            int(nprof), _cabi.dptr(Lf), _cabi.dptr(Qf), Lf.size, int(n), _cabi.dptr(stats), None, dev.index or 0)
    _cabi.check(rc)
    return stats.reshape((cents.shape[0],) + L.shape)


def _to_device(times, device=None):
    import torch

    from .search import _device

    return torch.from_numpy(np.ascontiguousarray(times, dtype=np.float64)).to(_device(device))


# ---------------------------------------------------------------------------------------------
# --fast
# ---------------------------------------------------------------------------------------------
def search_with_qffa_step(times, mean_f, mean_fdot=0, mean_fddot=0, nbin=16, nprof=64, npfact=2, oversample=8, n=1,
                          search_fdot=True, return_profiles=False):
    """One centre of the quasi-fast folding search (hendrics/efsearch.py:876-925): fold ``times`` (fp64, sorted, centred
    as in :func:`search_with_qffa`) at (mean_f, mean_fdot, mean_fddot) into ``nprof`` time slices of ``nbin`` bins and
    take the Z^2_n of the slices summed under every (linear, quadratic) bin-shift pair.  Returns ``(fgrid, fdotgrid,
    stats)`` of shape (nL, nQ) [+ the (nprof, nbin) profiles]."""
    times = np.asarray(times, dtype=np.float64)
    plane = ShiftPlane(nbin, npfact, oversample, search_fdot)
    to_f, to_fdot = plane.scale(times[0], times[-1])
    found = _qffa_device_call(_to_device(times), times[0], times[-1], [(mean_f, mean_fdot, mean_fddot)], nbin, nprof,
                              plane.L, plane.Q, n, want_profiles=return_profiles)
    stats, profiles = (found[0][0], found[1][0]) if return_profiles else (found[0], None)
    out = (plane.L * to_f + mean_f, plane.Q * to_fdot + mean_fdot, stats)
    return out + (profiles,) if return_profiles else out


def search_with_qffa(times, f0, f1, fdot=0, fddot=0, nbin=16, nprof=None, npfact=2, oversample=8, n=1, search_fdot=True,
                     t0=None, t1=None, silent=False, force_memmap=False):
    """'Quite fast folding' search between ``f0`` and ``f1`` (hendrics/efsearch.py:928-1089).  Returns ``(fgrid,
    fdotgrid, stats, step, fdotstep, length)`` with 2-D arrays of shape (nQ, n_centres * nL), or ``(freqs, stats, step,
    length)`` without the fdot dimension.

    The reference's serial loop over centre frequencies is one ``hx_qffa_search_dev`` call per rank: the centres are
    sharded over the ranks when ``torch.distributed`` is initialised and every rank streams the events once."""
    from . import dist

    if nprof is None:
        nprof = 4 * 2 * nbin * npfact  # enough slices for the signal to stay inside one bin during a slice

    def announce(err, too_large):
        if too_large:
            warnings.warn(f"Maximum error on the phase expected when casting to double: {err}")
            warnings.warn("Casting to double produces non-negligible phase errors. Please use shorter light curves.")

    ev = CentredTimes.prepare(times, t0, t1, f1, fdot, nbin, announce)
    centres, _ = qffa_centres(f0, f1, ev.length, npfact, fdot, fddot)
    plane = ShiftPlane(nbin, npfact, oversample, search_fdot)
    to_f, to_fdot = plane.scale(ev.first, ev.last)
    if dist.is_distributed():  # replicated list: 1/N per rank over PCIe + one NVLink all-gather, then the device entry
        times_d = ev.on_device()

        def fold(rows):
            return _qffa_device_call(times_d, ev.first, ev.last, rows, nbin, nprof, plane.L, plane.Q, n)
    else:  # one GPU: the host entry streams the events up while K2 folds what has arrived
        def fold(rows):
            return _qffa_host_call(ev.times, rows, nbin, nprof, plane.L, plane.Q, n)

    stats = dist.sharded_rows(fold, np.asarray(centres, dtype=np.float64))  # (n_centres, nL, nQ)

    nL, nQ = plane.L.shape
    shape = (nL * len(centres), nQ)
    log.info(f"Initializing result arrays of shape {shape}")
    fgrid, fdotgrid, statgrid = (_result_array(shape, force_memmap) for _ in range(3))
    mean_f = np.array([c[0] for c in centres])[:, None, None]
    mean_fdot = np.array([c[1] for c in centres])[:, None, None]
    fgrid[:] = (plane.L * to_f + mean_f).reshape(shape)
    fdotgrid[:] = (plane.Q * to_fdot + mean_fdot).reshape(shape)
    statgrid[:] = np.asarray(stats).reshape(shape)

    step = np.median(np.diff(fgrid[:, 0]))
    if not search_fdot:
        return fgrid.T[0], statgrid.T[0], step, ev.length
    return fgrid.T, fdotgrid.T, statgrid.T, step, np.median(np.diff(fdotgrid[0])), ev.length


# ---------------------------------------------------------------------------------------------
# transient scan
# ---------------------------------------------------------------------------------------------
class TransientResults:
    """Same attributes as hendrics/efsearch.py:488-496."""

    oversample: int = None
    f0: float = None
    f1: float = None
    fdot: float = None
    nave: np.array = None
    freqs: np.array = None
    times: np.array = None
    stats: np.array = None


def transient_search(times, f0, f1, fdot=0, nbin=16, nprof=None, n=1, t0=None, t1=None, oversample=4,
                     force_memmap=False):
    """Scan for pulsations that are present only part of the time (hendrics/efsearch.py:499-650): for every centre
    frequency of the ladder, the Z^2_n of the profile accumulated over 1, 2, 4, ... consecutive time slices, for every
    position of such a group.  ``results.stats`` has shape (levels, slices, centres).

    K2 folds the events into ``nprof`` slices per centre; ``hx_transient_search_dev`` forms the statistics of all
    averaging levels from the per-slice harmonic coefficients.  Centres are sharded over ranks."""
    from . import dist

    if nprof is None:
        nprof = 4 * 2 * nbin

    def announce(err, too_large):
        log.info(f"Maximum error on the phase expected when casting to double: {err}")
        if too_large:
            warnings.warn("Casting to double produces non-negligible phase errors. Please use shorter light curves.")

    ev = CentredTimes.prepare(times, t0, t1, f1, fdot, nbin, announce)
    freqs = centre_ladder(f0, f1, 1 / ev.length / oversample)
    centres = np.stack([freqs, np.full(freqs.shape, np.double(fdot + 0))], axis=1)
    # `nprof = int(2 ** np.log2(nprof))` (hendrics/efsearch.py:436-438) is not a power-of-two truncation, but the
    # floating-point round trip does drop the LAST time slice for some nprof (10 -> 9, 100 -> 99, 200 -> 199): the
    # events are histogrammed into nprof slices and the averaging sees the first nprof_used of them
    nprof_used = int(2 ** np.log2(nprof))
    n_levels = int(np.log2(nprof_used))
    times_d = ev.on_device()

    def local(rows):
        import torch

        rows = np.ascontiguousarray(rows, dtype=np.float64)
        mf, mfd = np.ascontiguousarray(rows[:, 0]), np.ascontiguousarray(rows[:, 1])
        out = torch.empty((rows.shape[0], n_levels, nprof_used), dtype=torch.float64, device=times_d.device)
        with torch.cuda.device(times_d.device):
            rc = _cabi.lib().hx_transient_search_dev(
                times_d.data_ptr(), times_d.numel(), float(ev.first), float(ev.last), _cabi.dptr(mf), _cabi.dptr(mfd),
                rows.shape[0], int(nbin), int(nprof), int(nprof_used), int(n), int(n_levels), out.data_ptr(),
                torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc)
        return out.cpu().numpy()

    per_centre = dist.sharded_rows(local, centres)  # (centres, levels, slices)
    stats = _result_array((n_levels, per_centre.shape[2], per_centre.shape[0]), force_memmap)
    stats[:] = np.transpose(per_centre, (1, 2, 0))

    results = TransientResults()
    results.oversample, results.f0, results.f1, results.fdot = oversample, f0, f1, fdot
    results.nave = 2 ** np.arange(n_levels)
    results.freqs = freqs.copy()
    results.times = (ev.last - ev.first) / nprof * np.arange(per_centre.shape[2])
    results.stats = stats
    return results
