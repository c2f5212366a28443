"""ctypes loader for oracle/libhxoracle.so (the plain-C restatement, oracle/hx_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/hx_oracle.py for the scope/parity statement."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    so = os.path.join(_HERE, "libhxoracle.so")
    src = os.path.join(_HERE, "hx_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libhxoracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.ora_fold_search.restype = ctypes.c_int
        _LIB.ora_fold_search.argtypes = [_dp, ctypes.c_int64, _dp, ctypes.c_int64, _dp, ctypes.c_int64,
                                         ctypes.c_int, ctypes.c_int, _dp, _i64p, ctypes.c_int]
        _LIB.ora_qffa_profiles.restype = ctypes.c_int
        _LIB.ora_qffa_profiles.argtypes = [_dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_int, ctypes.c_int, _u64p]
        _LIB.ora_fast_step.restype = ctypes.c_int
        _LIB.ora_fast_step.argtypes = [_u64p, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int64, ctypes.c_int,
                                       _dp, ctypes.c_int]
        _LIB.ora_max_threads.restype = ctypes.c_int
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return int(lib().ora_max_threads())


def fold_search(ts, freqs, fdots, nbin, nharm, want_profiles=False, nthreads=0):
    """stats (n_fd, n_f) [, profiles (n_fd, n_f, nbin) int64] for already-selected times."""
    ts, freqs, fdots = _d(ts), _d(np.atleast_1d(freqs)), _d(np.atleast_1d(fdots))
    stats = np.zeros((fdots.size, freqs.size))
    prof = np.zeros((fdots.size, freqs.size, nbin), dtype=np.int64) if want_profiles else None
    rc = lib().ora_fold_search(ts.ctypes.data_as(_dp), ts.size, freqs.ctypes.data_as(_dp), freqs.size,
                               fdots.ctypes.data_as(_dp), fdots.size, int(nbin), int(nharm),
                               stats.ctypes.data_as(_dp),
                               prof.ctypes.data_as(_i64p) if want_profiles else None, int(nthreads))
    if rc != 0:
        raise RuntimeError("ora_fold_search failed")
    return (stats, prof) if want_profiles else stats


def qffa_profiles(ts, mean_f, mean_fd, mean_fdd, nbin, nprof):
    ts = _d(ts)
    prof = np.zeros((nprof, nbin), dtype=np.uint64)
    lib().ora_qffa_profiles(ts.ctypes.data_as(_dp), ts.size, float(mean_f), float(mean_fd), float(mean_fdd),
                            int(nbin), int(nprof), prof.ctypes.data_as(_u64p))
    return prof


def fast_step(profiles, L, Q, nharm, nthreads=0):
    profiles = np.ascontiguousarray(profiles, dtype=np.uint64)
    L, Q = _d(L), _d(Q)
    stats = np.zeros(L.shape)
    lib().ora_fast_step(profiles.ctypes.data_as(_u64p), profiles.shape[0], profiles.shape[1],
                        L.ctypes.data_as(_dp), Q.ctypes.data_as(_dp), L.size, int(nharm),
                        stats.ctypes.data_as(_dp), int(nthreads))
    return stats
