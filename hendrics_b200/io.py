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


# (field, default) in the order hendrics/io.py:85-105 assigns them; the saved file is the dict of these attributes, so
# the names are the file format
_EF_FIELDS = (
    ("freq", None), ("stat", None), ("kind", None), ("nbin", None), ("oversample", None), ("N", None), ("peaks", None),
    ("peak_stat", None), ("best_fits", None), ("fdots", 0), ("fddots", 0), ("M", None), ("segment_size", 1e32),
    ("filename", ""), ("parfile", None), ("emin", None), ("emax", None), ("pepoch", None), ("mjdref", None),
    ("upperlim", None), ("ncounts", None),
)
_EF_POSITIONAL = ("freq", "stat", "kind", "nbin", "N", "oversample", "M", "pepoch", "mjdref", "peaks", "peak_stat",
                  "best_fits", "fdots", "fddots", "segment_size", "filename", "parfile", "emin", "emax", "ncounts",
                  "upperlim")
_CLASS_TAG = "__sr__class__type__"


class EFPeriodogram:
    """Result of a folding search: the container of hendrics/io.py:60-105 (same constructor signature, positional
    order included, same attributes)."""

    def __init__(self, *args, **kwargs):
        if len(args) > len(_EF_POSITIONAL):
            raise TypeError(f"EFPeriodogram takes at most {len(_EF_POSITIONAL)} positional arguments")
        given = dict(zip(_EF_POSITIONAL, args))
        for k in kwargs:
            if k not in _EF_POSITIONAL:
                raise TypeError(f"EFPeriodogram got an unexpected keyword argument {k!r}")
            if k in given:
                raise TypeError(f"EFPeriodogram got multiple values for argument {k!r}")
        given.update(kwargs)
        for name, default in _EF_FIELDS:
            setattr(self, name, given.get(name, default))

    def detection_threshold(self, conflevel=99.0):
        """Statistic value exceeded by noise with probability 1 - conflevel/100 over the independent trials of the plane
        (trials / oversample), for this periodogram's statistic."""
        ntrial = self.stat.size
        if getattr(self, "oversample", None) is not None:
            ntrial = int(ntrial / self.oversample)
        epsilon = 1 - conflevel / 100
        if self.kind == "Z2n":
            return z2_n_detection_level(epsilon=epsilon, n=self.N, ntrial=ntrial, n_summed_spectra=int(self.M))
        return fold_detection_level(nbin=int(self.nbin), epsilon=epsilon, ntrial=ntrial)

    def find_peaks(self, conflevel=99.0):
        """Peaks above the ``conflevel`` threshold (hendrics/io.py:107-144): for a 1-D periodogram the maxima of the
        regions above it, for a plane the ten highest pixels as (f, fdot) pairs.  Stores them in ``peaks`` /
        ``peak_stat`` when there are any and returns ``(peaks, peak_stat)``."""
        threshold = self.detection_threshold(conflevel)
        stat = np.asarray(self.stat)
        if stat.ndim == 1:
            peaks, heights = search_best_peaks(self.freq, stat, threshold)
            peaks, heights = np.asarray(peaks), np.asarray(heights)
        else:
            at = find_peaks_in_image(stat, n=10, threshold_abs=threshold)
            rows, cols = (at[:, 0], at[:, 1]) if len(at) else (np.zeros(0, dtype=int),) * 2
            peaks = np.stack([np.asarray(self.freq)[rows, cols], np.asarray(self.fdots)[rows, cols]], axis=1)
            heights = stat[rows, cols]
        if len(peaks) > 0:
            self.peaks, self.peak_stat = peaks, heights
        return peaks, heights


def _model_file(fname, i):
    return fname.replace(HEN_FILE_EXTENSION, f"__mod{i}__.p")


def save_folding(efperiodogram, fname):
    """Pickle back-end of hendrics/io.py:656-671: ``fname`` holds the attribute dict plus the class tag; every entry of
    ``best_fits`` goes to its own ``<root>__mod<i>__.p`` side file instead."""
    content = dict(vars(efperiodogram))
    content[_CLASS_TAG] = "EFPeriodogram"
    fits = content.get("best_fits")
    if fits is not None:
        del content["best_fits"]
        for i, model in enumerate(fits):
            with open(_model_file(fname, i), "wb") as fobj:
                pickle.dump(model, fobj)
    with open(fname, "wb") as fobj:
        pickle.dump(content, fobj)


def load_folding(fname):
    """Inverse of :func:`save_folding` (hendrics/io.py:674-696): side files found next to ``fname`` become
    ``best_fits``; a scalar ``peaks`` is wrapped in a list."""
    with open(fname, "rb") as fobj:
        content = pickle.load(fobj)
    content.pop(_CLASS_TAG)
    ef = EFPeriodogram()
    vars(ef).update(content)
    side_files = sorted(glob.glob(_model_file(fname, "*")))
    if side_files:
        ef.best_fits = []
        for mfile in side_files:
            with open(mfile, "rb") as fobj:
                ef.best_fits.append(pickle.load(fobj))
    if ef.peaks is not None and np.ndim(ef.peaks) == 0:
        ef.peaks = [ef.peaks]
    return ef
