"""Host-side mirror of ``hendrics/ffa.py`` (fast folding algorithm, ``HENzsearch --ffa``).

    ffa_search            hendrics/ffa.py:366-395
    ffa                   hendrics/ffa.py:337-349  (-> ``_ffa`` :311-334, ``ffa_step`` :291-308)
    _quick_rebin          hendrics/ffa.py:352-363
    z_n_fast_cached(_all) hendrics/ffa.py:107-188, h_test :191-207   (single-profile helpers, numpy)

The folding itself (all ``log2(ntables)`` steps and the Z^2_n of every table) runs in ``hx_ffa_dev``
(``csrc/hx_ffa.cu``); this module builds the trial-period grid, the harmonic tables (with numpy, exactly as the
reference does) and loops over the integer bin periods.  Light curves are integer counts (``search_with_ffa`` always
histograms an event list): they are kept as uint32 on the GPU, where the reference adds integer-valued doubles --
the sums are the same numbers.  A light curve with fractional or negative counts is refused (``ValueError``).
"""

from __future__ import annotations

import functools

import numpy as np

from . import _cabi

__all__ = ["ffa_search", "ffa", "h_test", "z_n_fast_cached", "z_n_fast_cached_all"]


# ---- single-profile statistics (host helpers; the searches use the CUDA epilogues) ----------------------------------
def cached_sin_harmonics(nbin, z_n_n):
    """hendrics/ffa.py:38-46."""
    twopiphases = np.pi * 2 * np.arange(0, 1, 1.0 / nbin)
    return np.tile(np.sin(twopiphases), z_n_n)


def cached_cos_harmonics(nbin, z_n_n):
    """hendrics/ffa.py:49-56."""
    twopiphases = np.pi * 2 * np.arange(0, 1, 1.0 / nbin)
    return np.tile(np.cos(twopiphases), z_n_n)


def z_n_fast_cached(norm, n=2):
    """Z^2_n of a binned profile, hendrics/ffa.py:59-136."""
    norm = np.asarray(norm, dtype=np.float64)
    cs, cc = cached_sin_harmonics(norm.size, n), cached_cos_harmonics(norm.size, n)
    N = norm.size
    result = 0.0
    for k in range(1, n + 1):
        result += np.sum(cc[: N * k : k] * norm) ** 2 + np.sum(cs[: N * k : k] * norm) ** 2
    return 2 / np.sum(norm) * result


def z_n_fast_cached_all(norm, nmax=20):
    """Z^2_1 .. Z^2_nmax of a binned profile, hendrics/ffa.py:139-188."""
    norm = np.asarray(norm, dtype=np.float64)
    cs, cc = cached_sin_harmonics(norm.size, nmax), cached_cos_harmonics(norm.size, nmax)
    ks = np.arange(1, nmax + 1, dtype=int)
    N = norm.size
    all_zs = np.zeros(ks.size)
    total = 0.0
    for k in ks:
        local = np.sum(cc[: N * k : k] * norm) ** 2 + np.sum(cs[: N * k : k] * norm) ** 2
        all_zs[k - 1] = local + total
        total += local
    return ks, 2 / np.sum(norm) * all_zs


def h_test(norm, nmax=20):
    """H statistics (de Jager+89), hendrics/ffa.py:191-207."""
    ks, zs = z_n_fast_cached_all(norm, nmax=nmax)
    hs = zs - 4 * ks + 4
    idx = np.argmax(hs)
    return hs[idx], ks[idx]


# ---- FFA ---------------------------------------------------------------------------------------------------------------
def _ffa_tables(bin_period, z_n_n=2):
    """[2][z_n_n][bin_period] cos / sin values harmonic k multiplies bin j with: the stride-k slices (ffa.py:98-102) of
    the tables ``_ffa`` tiles from ``np.arange(0, 1, 1 / nbin)`` (ffa.py:320-328) -- whose length is nbin + 1 for some
    nbin (49, 98, 103, ...), which skews the harmonics k >= 2 in the reference; reproduced here."""
    twopiphases = np.pi * 2 * np.arange(0, 1, 1 / bin_period)
    cc = np.tile(np.cos(twopiphases), z_n_n)
    ss = np.tile(np.sin(twopiphases), z_n_n)
    N = int(bin_period)
    tab = np.zeros((2, z_n_n, N))
    for k in range(1, z_n_n + 1):
        tab[0, k - 1] = cc[: N * k : k]
        tab[1, k - 1] = ss[: N * k : k]
    return tab


def _counts_to_device(counts, device=None):
    """uint32 light curve on the GPU (as an int32 tensor: only its storage is used)."""
    import torch

    from .search import _device

    if isinstance(counts, torch.Tensor):
        if counts.dtype != torch.int32 or not counts.is_cuda:
            raise TypeError("device light curves must be int32 CUDA tensors holding uint32 counts")
        c = counts.contiguous()
    else:
        a = np.asarray(counts)
        if a.ndim != 1:
            raise ValueError("counts must be one-dimensional")
        if a.dtype.kind == "f":
            if not np.all(np.isfinite(a)) or np.any(a != np.rint(a)):
                raise ValueError("hendrics_b200.ffa folds integer counts; got a light curve with fractional values")
        elif a.dtype.kind not in "iu":
            raise TypeError(f"unsupported light-curve dtype {a.dtype}")
        if a.size and (a.min() < 0 or a.max() >= 2**32):
            raise ValueError("counts must be in [0, 2^32)")
        c = torch.from_numpy(a.astype(np.uint32).view(np.int32)).to(_device(device))
    if c.numel() and int(c.sum(dtype=torch.int64).item()) >= 2**32:  # device light curves are assumed non-negative as uint32
        raise ValueError("total counts must be below 2^32 (uint32 table sums)")
    return c


@functools.lru_cache(maxsize=4)
def _period_fractions(ntables):
    """n / (ntables - 1), n = 0 .. ntables - 1 (ffa.py:314); the same array for every bin period of a search"""
    f = np.arange(ntables) / (ntables - 1)
    f.setflags(write=False)
    return f


_PINNED = {}  # (device index, ntables) -> pinned host buffer for the statistics of one FFA


def _ffa_device(counts_d, bin_period, z_n_n=2, reuse_host_buffer=False):
    """``ffa`` on a device light curve: (periods, stats) as host arrays, ffa.py:337-349.  With ``reuse_host_buffer`` the
    statistics come back as a view of a cached pinned buffer that the next call overwrites (``ffa_search`` copies them)."""
    import torch

    N_raw = counts_d.numel()
    bin_period = int(bin_period)
    ntables = int(2 ** np.ceil(np.log2(N_raw // bin_period + 1)))
    if ntables <= 1:
        return np.zeros(1), np.zeros(1)
    periods = bin_period + _period_fractions(ntables)  # == [bin_period + n / (ntables - 1) ...]
    tab = np.ascontiguousarray(_ffa_tables(bin_period, z_n_n))
    stats = torch.empty(ntables, dtype=torch.float64, device=counts_d.device)
    with torch.cuda.device(counts_d.device):
        stream = torch.cuda.current_stream(counts_d.device).cuda_stream
        _cabi.check(_cabi.lib().hx_ffa_dev(counts_d.data_ptr(), N_raw, bin_period, int(z_n_n), _cabi.dptr(tab),
                                           stats.data_ptr(), ntables, stream))
    if not reuse_host_buffer:
        return periods, stats.cpu().numpy()
    key = (counts_d.device.index, ntables)
    host = _PINNED.get(key)
    if host is None:
        _PINNED.clear()  # one size at a time: a search uses a single ntables (two when the light curve is rebinned)
        host = _PINNED[key] = torch.empty(ntables, dtype=torch.float64, pin_memory=True)
    host.copy_(stats)
    return periods, host.numpy()


def ffa(array, bin_period, z_n_n=2):
    """Fast folding algorithm search at one integer bin period (hendrics/ffa.py:337-349): returns
    ``(periods, stats)``, ``ntables`` trial periods ``bin_period + i / (ntables - 1)`` (in bins) and their Z^2_n."""
    return _ffa_device(_counts_to_device(array), bin_period, z_n_n=z_n_n)


def _quick_rebin(counts_d, current_rebin):
    """hendrics/ffa.py:352-363 on the device."""
    import torch

    n = int(counts_d.numel() // current_rebin)
    out = torch.empty(n, dtype=torch.int32, device=counts_d.device)
    with torch.cuda.device(counts_d.device):
        stream = torch.cuda.current_stream(counts_d.device).cuda_stream
        _cabi.check(_cabi.lib().hx_rebin_u32_dev(counts_d.data_ptr(), counts_d.numel(), int(current_rebin), out.data_ptr(),
                                                 stream))
    return out


def ffa_search(counts, dt, period_min, period_max):
    """Mirror of ``hendrics/ffa.py:366-395``: FFA at every integer bin period between ``period_min / dt`` and
    ``period_max / dt`` (rebinning the light curve by powers of two as the period grows); returns
    ``(bin_periods [s], stats)``.  Always Z^2_2, like the reference (``ffa`` is called without ``z_n_n``)."""
    counts_d = _counts_to_device(counts)
    pmin = np.floor(period_min / dt)
    pmax = np.ceil(period_max / dt)
    bin_periods = None
    stats = None

    current_rebin = 1
    rebinned_counts = counts_d
    for bin_period in np.arange(pmin, pmax + 1, dtype=int):
        # Only powers of two
        rebin = int(2 ** np.floor(np.log2(bin_period / pmin)))
        if rebin > current_rebin:
            current_rebin = rebin
            rebinned_counts = _quick_rebin(counts_d, current_rebin)

        if bin_period % current_rebin != 0:
            continue

        per, st = _ffa_device(rebinned_counts, bin_period // current_rebin, reuse_host_buffer=True)
        per = per * current_rebin

        if per[0] == 0:
            continue
        elif bin_periods is None:
            bin_periods = per[:-1] * dt
            stats = st[:-1].copy()  # st is a view of a reused pinned buffer
        else:
            bin_periods = np.concatenate((bin_periods, per[:-1] * dt))
            stats = np.concatenate((stats, st[:-1]))

    return bin_periods, stats
