"""Output container and file format of the search path.

Mirrors ``hendrics/io.py``: ``EFPeriodogram`` (:60-144), ``save_folding`` / ``load_folding``
(:656-696) with the pickle back-end (:894-906) that HENDRICS itself falls back to when netCDF4 is
absent (``HEN_FILE_EXTENSION = ".p"``, io.py:36-45).  A file written here is the same dict of
``EFPeriodogram.__dict__`` + ``"__sr__class__type__": "EFPeriodogram"`` and loads in HENDRICS.

Event input: the reference reads mission FITS / HENDRICS event files through stingray+astropy
(out of scope).  Here events arrive as fp64/long-double arrays:
``.npy`` (times), ``.npz`` (time, gti, mjdref, pi, energy) or a HENDRICS-style pickled dict.
"""

from __future__ import annotations

import copy
import glob
import os
import pickle

import numpy as np

from .base import find_peaks_in_image, fold_detection_level, z2_n_detection_level

HEN_FILE_EXTENSION = ".p"

__all__ = ["EFPeriodogram", "EventList", "Lightcurve", "save_folding", "load_folding", "load_events", "save_events",
           "save_lcurve", "load_lcurve",
           "get_file_type", "search_best_peaks", "HEN_FILE_EXTENSION"]


class EventList:
    """Minimal stand-in for ``stingray.EventList``: the attributes the search path reads
    (``time``, ``gti``, ``mjdref``, optional ``pi`` / ``energy``)."""

    def __init__(self, time=None, gti=None, mjdref=0, pi=None, energy=None, mission="", instr="", **_ignored):
        self.time = None if time is None else np.asarray(time)
        if gti is None and self.time is not None and self.time.size > 0:
            gti = np.array([[self.time[0], self.time[-1]]], dtype=self.time.dtype)
        self.gti = None if gti is None else np.asarray(gti)
        self.mjdref = mjdref
        self.pi = pi
        self.energy = energy
        self.mission = mission
        self.instr = instr


class Lightcurve:
    """Minimal stand-in for ``stingray.Lightcurve``: what ``folding_search`` reads from a light curve
    (hendrics/efsearch.py:1154-1157): bin ``time``s, ``counts`` per bin (the fold weights), ``gti``."""

    def __init__(self, time=None, counts=None, gti=None, dt=None, mjdref=0, mission="", instr="", **_ignored):
        self.time = np.asarray(time)
        self.counts = np.asarray(counts)
        if self.time.shape != self.counts.shape:
            raise ValueError("time and counts must have the same shape")
        if dt is None:
            dt = np.median(np.diff(self.time)) if self.time.size > 1 else 1.0
        self.dt = dt
        if gti is None and self.time.size > 0:
            gti = np.array([[self.time[0] - dt / 2, self.time[-1] + dt / 2]])
        self.gti = None if gti is None else np.asarray(gti)
        self.mjdref = mjdref
        self.mission = mission
        self.instr = instr


def save_lcurve(lcurve, fname):
    out = {"time": lcurve.time, "counts": lcurve.counts, "gti": lcurve.gti, "dt": lcurve.dt, "mjdref": lcurve.mjdref,
           "mission": lcurve.mission, "instr": lcurve.instr, "__sr__class__type__": "Lightcurve"}
    with open(fname, "wb") as fobj:
        pickle.dump(out, fobj)


def load_lcurve(fname):
    with open(fname, "rb") as fobj:
        out = pickle.load(fobj)
    return Lightcurve(**{k: v for k, v in out.items() if not k.startswith("__")})


def save_events(events, fname):
    out = {"time": events.time, "gti": events.gti, "mjdref": events.mjdref, "pi": events.pi,
           "energy": events.energy, "mission": events.mission, "instr": events.instr,
           "__sr__class__type__": "EventList"}
    if fname.endswith(".npz"):
        np.savez(fname, **{k: v for k, v in out.items() if v is not None and k != "__sr__class__type__"})
    else:
        with open(fname, "wb") as fobj:
            pickle.dump(out, fobj)


def load_events(fname):
    if fname.endswith(".npy"):
        return EventList(time=np.load(fname))
    if fname.endswith(".npz"):
        with np.load(fname, allow_pickle=False) as d:
            kw = {k: d[k] for k in d.files}
        for k in ("mjdref",):
            if k in kw:
                kw[k] = kw[k].item() if np.ndim(kw[k]) == 0 else kw[k]
        for k in ("mission", "instr"):
            if k in kw:
                kw[k] = str(kw[k])
        return EventList(**kw)
    with open(fname, "rb") as fobj:
        out = pickle.load(fobj)
    out = {k: v for k, v in out.items() if not k.startswith("__")}
    return EventList(**out)


def get_file_type(fname):
    """("events", EventList), ("lc", Lightcurve) or ("folding", EFPeriodogram) -- the types this path handles
    (hendrics/io.py:473-522)."""
    if fname.endswith((".npy", ".npz")):
        return "events", load_events(fname)
    with open(fname, "rb") as fobj:
        d = pickle.load(fobj)
    kind = d.get("__sr__class__type__", "EventList") if isinstance(d, dict) else None
    if kind == "EFPeriodogram":
        return "folding", load_folding(fname)
    if kind == "EventList":
        return "events", load_events(fname)
    if kind == "Lightcurve":
        return "lc", load_lcurve(fname)
    raise ValueError(f"File format of {fname} not understood (type {kind!r}); FITS / netCDF files are "
                     "outside the B200 search path")


def search_best_peaks(x, stat, threshold):
    """``stingray.pulse.search.search_best_peaks`` [UPSTREAM-RECALLED] (used at hendrics/io.py:127):
    contiguous regions with stat >= threshold -> (x, stat) at each region's arg-max, sorted by
    stat descending."""
    x = np.asarray(x)
    stat = np.asarray(stat)
    above = stat >= threshold
    best_x, best_stat = [], []
    i, n = 0, stat.size
    while i < n:
        if not above[i]:
            i += 1
            continue
        j = i
        while j < n and above[j]:
            j += 1
        m = i + int(np.argmax(stat[i:j]))
        best_x.append(x[m])
        best_stat.append(stat[m])
        i = j
    order = np.argsort(best_stat)[::-1]
    return [best_x[o] for o in order], [best_stat[o] for o in order]


class EFPeriodogram:
    """Same fields as hendrics/io.py:60-105."""

    def __init__(self, freq=None, stat=None, kind=None, nbin=None, N=None, oversample=None, M=None, pepoch=None,
                 mjdref=None, peaks=None, peak_stat=None, best_fits=None, fdots=0, fddots=0, segment_size=1e32,
                 filename="", parfile=None, emin=None, emax=None, ncounts=None, upperlim=None):
        self.freq = freq
        self.stat = stat
        self.kind = kind
        self.nbin = nbin
        self.oversample = oversample
        self.N = N
        self.peaks = peaks
        self.peak_stat = peak_stat
        self.best_fits = best_fits
        self.fdots = fdots
        self.fddots = fddots
        self.M = M
        self.segment_size = segment_size
        self.filename = filename
        self.parfile = parfile
        self.emin = emin
        self.emax = emax
        self.pepoch = pepoch
        self.mjdref = mjdref
        self.upperlim = upperlim
        self.ncounts = ncounts

    def find_peaks(self, conflevel=99.0):
        """hendrics/io.py:107-144."""
        ntrial = self.stat.size
        if hasattr(self, "oversample") and self.oversample is not None:
            ntrial /= self.oversample
            ntrial = int(ntrial)
        epsilon = 1 - conflevel / 100
        if self.kind == "Z2n":
            threshold = z2_n_detection_level(epsilon=epsilon, n=self.N, ntrial=ntrial, n_summed_spectra=int(self.M))
        else:
            threshold = fold_detection_level(nbin=int(self.nbin), epsilon=epsilon, ntrial=ntrial)
        if len(self.stat.shape) == 1:
            best_peaks, best_stat = search_best_peaks(self.freq, self.stat, threshold)
        else:
            best_cands = find_peaks_in_image(self.stat, n=10, threshold_abs=threshold)
            best_peaks, best_stat = [], []
            for idx in best_cands:
                best_peaks.append([self.freq[idx[0], idx[1]], self.fdots[idx[0], idx[1]]])
                best_stat.append(self.stat[idx[0], idx[1]])
        best_peaks = np.asarray(best_peaks)
        best_stat = np.asarray(best_stat)
        if len(best_peaks) > 0:
            self.peaks = best_peaks
            self.peak_stat = best_stat
        return best_peaks, best_stat


def save_folding(efperiodogram, fname):
    """hendrics/io.py:656-671 (pickle back-end; fitted models are pickled into side files)."""
    outdata = copy.copy(efperiodogram.__dict__)
    outdata["__sr__class__type__"] = "EFPeriodogram"
    if "best_fits" in outdata and efperiodogram.best_fits is not None:
        for i, b in enumerate(efperiodogram.best_fits):
            mfile = fname.replace(HEN_FILE_EXTENSION, f"__mod{i}__.p")
            with open(mfile, "wb") as fobj:
                pickle.dump(b, fobj)
        outdata.pop("best_fits")
    with open(fname, "wb") as fobj:
        pickle.dump(outdata, fobj)


def load_folding(fname):
    """hendrics/io.py:674-696."""
    with open(fname, "rb") as fobj:
        data = pickle.load(fobj)
    data.pop("__sr__class__type__")
    ef = EFPeriodogram()
    for key in data.keys():
        setattr(ef, key, data[key])
    modelfiles = sorted(glob.glob(fname.replace(HEN_FILE_EXTENSION, "__mod*__.p")))
    if len(modelfiles) >= 1:
        bmodels = []
        for mfile in modelfiles:
            if os.path.exists(mfile):
                with open(mfile, "rb") as fobj:
                    bmodels.append(pickle.load(fobj))
        ef.best_fits = bmodels
    if ef.peaks is not None and len(np.asarray(ef.peaks).shape) == 0:
        ef.peaks = [ef.peaks]
    return ef
