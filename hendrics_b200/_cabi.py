"""ctypes binding of ``libhxsearch.so`` (C ABI declared in ``include/hx_search.h``).

There is NO CPU fallback: if the shared library is missing or a CUDA call fails, an
exception is raised.  Nothing in this package imports ``oracle/``.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# HX_LIB_PATH: development hook (tools/variant_bench.py loads experimental builds of the same ABI)
LIB_PATH = os.environ.get("HX_LIB_PATH") or os.path.join(_PKG_DIR, "libhxsearch.so")

HX_OK = 0
HX_ERR_ARG, HX_ERR_CUDA, HX_ERR_NOMEM, HX_ERR_RANGE, HX_ERR_INTERNAL = -1, -2, -3, -4, -5
HX_FOLD_DEFAULT, HX_FOLD_FORCE_GENERIC, HX_FOLD_FORCE_SCRATCH, HX_FOLD_EXACT_PHASE = 0, 1, 2, 4

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_u32p = ctypes.POINTER(ctypes.c_uint32)
_c_i64p = ctypes.POINTER(ctypes.c_int64)
_i64, _int, _uint, _dbl, _vp = ctypes.c_int64, ctypes.c_int, ctypes.c_uint, ctypes.c_double, ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol of include/hx_search.h
SIGNATURES = {
    "hx_version": (_int, []),
    "hx_last_error": (ctypes.c_char_p, []),
    "hx_device_count": (_int, [ctypes.POINTER(_int)]),
    "hx_launch_count": (_i64, []),
    "hx_fold_search": (_int, [_c_dp, _i64, _c_dp, _i64, _c_dp, _i64, _int, _int, _uint, _c_dp, _c_u32p, _int]),
    "hx_fold_search_dev": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _int, _int, _uint, _vp, _vp, _vp]),
    "hx_fold_windows_dev": (_int, [_vp, _i64, _c_dp, _c_dp, _i64, _int, _dbl, _dbl, _vp, _i64, _vp, _i64, _int, _int, _vp, _vp,
                                   ctypes.POINTER(_int), _vp]),
    "hx_wfold_search": (_int, [_c_dp, _c_dp, _dbl, _i64, _c_dp, _i64, _c_dp, _i64, _int, _int, _c_dp, _i64, _dbl, _dbl,
                               _c_dp, _c_dp, _int]),
    "hx_wfold_search_dev": (_int, [_vp, _vp, _dbl, _i64, _vp, _i64, _vp, _i64, _int, _int, _c_dp, _i64, _dbl, _dbl,
                                   _vp, _vp, _vp]),
    "hx_wfold_last_kernel_ms": (_int, [_c_dp, ctypes.POINTER(_int), _c_i64p]),
    "hx_qffa_search": (_int, [_c_dp, _i64, _c_dp, _c_dp, _c_dp, _i64, _int, _int, _c_dp, _c_dp, _i64, _int, _c_dp,
                              _c_u32p, _int]),
    "hx_qffa_search_dev": (_int, [_vp, _i64, _dbl, _dbl, _c_dp, _c_dp, _c_dp, _i64, _int, _int, _c_dp, _c_dp, _i64,
                                  _int, _vp, _vp, _vp]),
    "hx_transient_search_dev": (_int, [_vp, _i64, _dbl, _dbl, _c_dp, _c_dp, _i64, _int, _int, _int, _int, _int, _vp, _vp]),
    "hx_orbital_search": (_int, [_c_dp, _i64, _dbl, _dbl, _dbl, _dbl, _c_dp, _i64, _int, _int, _c_dp, _c_u32p, _int]),
    "hx_orbital_search_dev": (_int, [_vp, _i64, _dbl, _dbl, _dbl, _dbl, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "hx_zn_unbinned": (_int, [_c_dp, _i64, _c_dp, _i64, _c_dp, _i64, _int, _c_dp, _int]),
    "hx_zn_unbinned_dev": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _int, _vp, _vp]),
    "hx_zn_last_kernel_ms": (_int, [_c_dp, ctypes.POINTER(_int)]),
    "hx_topk": (_int, [_c_dp, _i64, _int, _c_i64p, _c_dp, _c_i64p, _int]),
    "hx_topk_dev": (_int, [_vp, _i64, _int, _c_i64p, _c_dp, _c_i64p, _vp]),
    "hx_hist1d_dev": (_int, [_vp, _i64, _dbl, _dbl, _dbl, _i64, _vp, _vp]),
    "hx_rebin_u32_dev": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "hx_ffa": (_int, [_c_u32p, _i64, _i64, _int, _c_dp, _c_dp, _i64, _int]),
    "hx_ffa_dev": (_int, [_vp, _i64, _i64, _int, _c_dp, _vp, _i64, _vp]),
    "hx_ffa_last_kernel_ms": (_int, [_c_dp, _c_i64p, _c_i64p, _c_i64p]),
    "hx_ffa_set_tuning": (_int, [_int]),
    "hx_microbench": (_int, [_int, _int, _c_dp, _int]),
    "hx_fold_last_kernel_ms": (_int, [_c_dp, _c_i64p, _c_i64p, _c_i64p]),
    "hx_fold_last_kernel_shape": (_int, [ctypes.POINTER(_int)] * 4 + [_c_dp, _c_dp]),
    "hx_qffa_last_kernel_ms": (_int, [_c_dp, _c_i64p]),
    "hx_fold_set_tuning": (_int, [_int, _int]),
    "hx_qffa_set_tuning": (_int, [_int]),
    "hx_fold_set_packing": (_int, [_int]),
    "hx_fold_set_generic_fast": (_int, [_int]),
    "hx_wfold_set_tuning": (_int, [_int, _int, _int, _int]),
}


class HxError(RuntimeError):
    """A libhxsearch call failed (bad argument, CUDA error, no device, ...)."""

    def __init__(self, code, msg):
        super().__init__(f"libhxsearch error {code}: {msg}")
        self.code = code


class HxRangeError(HxError, ValueError):
    pass


_lib = None


def build_library(force=False, quiet=True):
    """(Re)build libhxsearch.so in-tree with nvcc for sm_100a."""
    csrc = os.path.join(_PKG_DIR, "csrc")
    cmd = ["make", "-C", csrc, "-j8"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL if quiet else None)
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C hendrics_b200/csrc`. hendrics_b200 has no CPU fallback."
            )
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != HX_OK:
        msg = lib().hx_last_error().decode("utf-8", "replace")
        if rc == HX_ERR_RANGE:
            raise HxRangeError(rc, msg)
        raise HxError(rc, msg)


def dptr(a, ctype=_c_dp):
    return a.ctypes.data_as(ctype)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    n = _int(0)
    rc = lib().hx_device_count(ctypes.byref(n))
    return n.value if rc == HX_OK else 0


def launch_count():
    return int(lib().hx_launch_count())


def fold_last_kernel():
    """(ms, grid, block, smem_bytes) of the most recent K1 launch."""
    ms, g, b, s = _dbl(0.0), _i64(0), _i64(0), _i64(0)
    check(lib().hx_fold_last_kernel_ms(ctypes.byref(ms), ctypes.byref(g), ctypes.byref(b), ctypes.byref(s)))
    return ms.value, g.value, b.value, s.value


def fold_last_kernel_shape():
    """dict(R, mode, rowfd, m2t, fp64_per_unit, smem_wavefronts_per_unit) of the last K1 launch."""
    iv = [_int() for _ in range(4)]
    a, b = _dbl(), _dbl()
    check(lib().hx_fold_last_kernel_shape(*[ctypes.byref(x) for x in iv], ctypes.byref(a), ctypes.byref(b)))
    return {"R": iv[0].value, "mode": iv[1].value, "rowfd": iv[2].value, "m2t": iv[3].value,
            "fp64_per_unit": a.value, "smem_wavefronts_per_unit": b.value}


def ffa_last_kernel():
    """(ms, passes over HBM, folding steps, steps fused per pass) of the last hx_ffa* call."""
    ms = _dbl()
    v = [_i64() for _ in range(3)]
    check(lib().hx_ffa_last_kernel_ms(ctypes.byref(ms), *[ctypes.byref(x) for x in v]))
    return ms.value, v[0].value, v[1].value, v[2].value


def wfold_last_kernel():
    """(ms, cell kind, events folded) of the K1w launches of the most recent weighted fold: 0 fp64 cells, 1 u32
    (integer weights), 2 u32 (no weight array); events folded < n when zero-weight entries were dropped."""
    ms, k, n = _dbl(0.0), _int(-1), _i64(0)
    check(lib().hx_wfold_last_kernel_ms(ctypes.byref(ms), ctypes.byref(k), ctypes.byref(n)))
    return ms.value, k.value, n.value


def zn_last_kernel():
    """(ms, FP64 instructions per event*trial) of the last K4 (unbinned Z^2_n) launch."""
    ms, ops = _dbl(0.0), _int(0)
    check(lib().hx_zn_last_kernel_ms(ctypes.byref(ms), ctypes.byref(ops)))
    return ms.value, ops.value


def qffa_last_kernel():
    """(summed ms, launches) of the K2 launches of the most recent --fast call."""
    ms, n = _dbl(0.0), _i64(0)
    check(lib().hx_qffa_last_kernel_ms(ctypes.byref(ms), ctypes.byref(n)))
    return ms.value, n.value


def microbench(kind, reps=5, device=0):
    out = _dbl(0.0)
    check(lib().hx_microbench(int(kind), int(reps), ctypes.byref(out), int(device)))
    return out.value
