"""The few general helpers of ``hendrics/base.py`` the search CLI needs, plus the detection-level
functions HENDRICS re-exports from ``stingray.stats`` (``hendrics/base.py:21-28``).

Host-side scalar code only; nothing here is on the GPU hot path."""

from __future__ import annotations

import os
import sys

import numpy as np

__all__ = [
    "hen_root", "check_negative_numbers_in_args", "assign_value_if_none", "find_peaks_in_image",
    "p_single_trial_from_p_multitrial", "fold_detection_level", "z2_n_detection_level",
    "a_from_ssig", "pf_from_a", "power_upper_limit", "amplitude_upper_limit", "pf_upper_limit",
    "power_confidence_limits", "HENDRICS_STAR_VALUE",
]

HENDRICS_STAR_VALUE = "**"


def splitext_improved(path):
    """hendrics/base.py: like os.path.splitext but aware of double extensions (.fits.gz ...)."""
    for ext in [".tar.gz", ".tar.bz2", ".fits.gz", ".fits.Z", ".evt.gz", ".evt.Z"]:
        if path.endswith(ext):
            return path[: -len(ext)], path[-len(ext):]
    return os.path.splitext(path)


def hen_root(filename):
    """Root file name without ``_ev``, ``_lc`` ... (hendrics/base.py:339-369).

    >>> hen_root("blabla_ev_calib.nc")
    'blabla'
    >>> hen_root("blablu_lc_asrd_ev_lc.nc")
    'blablu_lc_asrd'
    """
    fname, _ = splitext_improved(filename)
    todo = True
    while todo:
        todo = False
        for ending in ["_ev", "_lc", "_pds", "_cpds", "_calib"]:
            if fname.endswith(ending):
                fname = fname[: -len(ending)]
                todo = True
    return fname


def assign_value_if_none(value, default):
    return default if value is None else value


def check_negative_numbers_in_args(args):
    """Prepend a space to negative numbers so argparse accepts them (hendrics/base.py:523-547).

    >>> check_negative_numbers_in_args(['events.nc', '-f', '103', '--fdot', '-2e-10'])[4]
    ' -2e-10'
    """
    if args is None:
        args = sys.argv[1:]
    newargs = []
    for arg in args:
        arg = str(arg)
        try:
            assert -float(arg) > 0
        except (ValueError, AssertionError):
            newargs.append(arg)
            continue
        newargs.append(" " + arg)
    return newargs


def find_peaks_in_image(image, n=5, rough=False, **kwargs):
    """hendrics/base.py:1109-1148, the branch taken without skimage (``argpartition`` top-n sorted
    by value descending; ``threshold_abs`` is ignored there too).  Ties are broken by ascending flat
    index so that CPU and GPU candidate lists are comparable."""
    image = np.asarray(image)
    flat = image.ravel()
    n = min(n, flat.size)
    order = None
    if flat.size >= GPU_TOPK_MIN_SIZE and flat.dtype == np.float64:
        order = _gpu_topk(np.ascontiguousarray(flat), n)
    if order is None:
        order = np.lexsort((np.arange(flat.size), -flat))[:n]
    return np.asarray([np.unravel_index(i, image.shape) for i in order])


GPU_TOPK_MIN_SIZE = 1 << 20  # below this a host sort is faster than the H2D copy


def _gpu_topk(flat, n):
    """k streaming arg-max passes on the GPU (csrc/hx_topk.cu); None when no device is usable."""
    import ctypes

    from . import _cabi

    if _cabi.device_count() < 1:
        return None
    idx = np.empty(n, dtype=np.int64)
    val = np.empty(n, dtype=np.float64)
    n_out = ctypes.c_int64(0)
    import torch

    # the caller's CURRENT device (one rank per GPU: never cuda:0 by default); the entry point restores it on return
    device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    rc = _cabi.lib().hx_topk(_cabi.dptr(flat), flat.size, int(n), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                             _cabi.dptr(val), ctypes.byref(n_out), int(device))
    _cabi.check(rc)
    return idx[: n_out.value]


# ---- stingray.stats subset  [UPSTREAM-RECALLED: source not under /root/reference] -------------
def p_single_trial_from_p_multitrial(p1, n):
    """1 - (1 - p1)^(1/n), evaluated through log1p/expm1 for small p."""
    logp = np.log1p(-np.asarray(p1, dtype=np.float64)) / n
    return -np.expm1(logp)


def fold_detection_level(nbin, epsilon=0.01, ntrial=1):
    from scipy import stats

    epsilon = p_single_trial_from_p_multitrial(epsilon, ntrial)
    return stats.chi2.isf(np.double(epsilon), nbin - 1)


def z2_n_detection_level(n=2, epsilon=0.01, ntrial=1, n_summed_spectra=1):
    from scipy import stats

    epsilon = p_single_trial_from_p_multitrial(epsilon, ntrial)
    return stats.chi2.isf(np.double(epsilon), 2 * n_summed_spectra * n) / n_summed_spectra


def a_from_ssig(ssig, ncounts):
    return np.sqrt(2 * ssig / ncounts)


def pf_from_a(a):
    return 2 * a / (1 + a)


def power_upper_limit(pmeas, n=1, c=0.95):
    """Upper limit on the signal power given a measured power (Vaughan+94), via ncx2."""
    from scipy import optimize, stats

    def ppf(x):
        return stats.ncx2(2 * n, x).ppf(1 - c)

    def isf(x):
        return stats.ncx2(2 * n, x).ppf(c)

    initial = isf(pmeas)
    res = optimize.minimize(lambda x, xm: np.abs(ppf(x[0]) - xm), [initial], pmeas, bounds=[(0, initial * 2)])
    return res.x[0]


def amplitude_upper_limit(pmeas, counts, n=1, c=0.95):
    return a_from_ssig(power_upper_limit(pmeas, n, c), counts)


def pf_upper_limit(pmeas, counts, n=1, c=0.95):
    return pf_from_a(amplitude_upper_limit(pmeas, counts, n=n, c=c))


def power_confidence_limits(preal, n=1, c=0.95):
    from scipy import stats

    rv = stats.ncx2(2 * n, preal)
    return rv.ppf([1 - c, c])
