"""Candidate post-processing of a search plane: the part of ``hendrics/efsearch.py`` that runs AFTER the kernels
(SURVEY.md section 8, row f1).

    print_qffa_results             hendrics/efsearch.py:1246-1259
    get_xy_boundaries_from_level   hendrics/efsearch.py:1262-1296   (matplotlib contour + find_nearest_contour there;
                                                                     a marching-squares tracer here, no matplotlib)
    get_boundaries_from_level      hendrics/efsearch.py:1299-1332
    _analyze_qffa_results          hendrics/efsearch.py:1335-1545   (table, side files, table meta)
    analyze_qffa_results           hendrics/efsearch.py:1548-1565

astropy is not available here, so the reference's ``astropy.table.Table`` is replaced by :class:`CandTable`, which
offers exactly what the consumers of the table use (``hendrics/plot.py:411-711``: ``table.meta[...]``,
``table["f"][0]``, iteration over rows with ``row["f_idx"]``) and writes the two file flavours the reference writes:
the ``ascii`` (space separated) side files and the ``.csv`` candidate list, with the per-column formats applied.
"""

from __future__ import annotations

import os

import numpy as np

from .base import (
    a_from_ssig,
    find_peaks_in_image,
    fold_detection_level,
    power_confidence_limits,
    z2_n_detection_level,
)
from .io import HEN_FILE_EXTENSION, load_folding

__all__ = ["CandTable", "trace_level_curves", "get_xy_boundaries_from_level", "get_boundaries_from_level",
           "print_qffa_results", "analyze_qffa_results", "CAND_COLUMNS", "N_CANDIDATES"]

N_CANDIDATES = 5

# (name, python type, print/write format or None) -- names, order and formats of hendrics/efsearch.py:1379-1433
_CAND_SCHEMA = (
    ("fname", str, None),
    ("mjd", float, None),
    ("power", float, ".2f"),
    ("f", float, None),
    ("f_err_n", float, None),
    ("f_err_p", float, None),
    ("fdot", float, ".2e"),
    ("fdot_err_n", float, None),
    ("fdot_err_p", float, None),
    ("fddot", float, "g"),
    ("power_cl_0.9", float, ".2f"),
    ("pulse_amp", float, ".2f"),
    ("pulse_amp_err", float, ".2f"),
    ("pulse_amp_cl_0.1", float, ".2f"),
    ("pulse_amp_cl_0.9", float, ".2f"),
    ("pulse_amp_ul_0.9", float, ".2f"),
    ("f_idx", int, None),
    ("fdot_idx", int, None),
    ("fddot_idx", int, None),
)
CAND_COLUMNS = [c[0] for c in _CAND_SCHEMA]


class _Row(dict):
    """One table row: a mapping column -> value (what ``for row in table: row["f_idx"]`` needs)."""


class CandTable:
    """A column table with per-column formats and a ``meta`` dict; the subset of ``astropy.table.Table`` that the
    candidate list is used through."""

    def __init__(self, data=None, names=None, dtype=None, formats=None):
        self.meta = {}
        self._cols = {}
        self._types = {}
        self.formats = dict(formats or {})
        if data is not None:  # Table({"freq": ..., "stat": ...})
            for k, v in data.items():
                self._cols[k] = list(np.asarray(v).tolist())
                self._types[k] = None
        elif names is not None:
            dtype = dtype or [None] * len(names)
            for k, t in zip(names, dtype):
                self._cols[k] = []
                self._types[k] = t

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def for_candidates(cls):
        return cls(names=CAND_COLUMNS, dtype=[c[1] for c in _CAND_SCHEMA],
                   formats={c[0]: c[2] for c in _CAND_SCHEMA if c[2]})

    def add_row(self, values):
        if len(values) != len(self._cols):
            raise ValueError(f"row has {len(values)} values, table has {len(self._cols)} columns")
        for (k, col), v in zip(self._cols.items(), values):
            t = self._types[k]
            col.append(v if t is None or v is None else t(v))

    # -- access ---------------------------------------------------------------------------------
    @property
    def colnames(self):
        return list(self._cols)

    def __len__(self):
        return len(next(iter(self._cols.values()))) if self._cols else 0

    def __iter__(self):
        for i in range(len(self)):
            yield _Row((k, c[i]) for k, c in self._cols.items())

    def __getitem__(self, key):
        if isinstance(key, str):
            t = self._types[key]
            return np.asarray(self._cols[key], dtype=object if t is str else None)
        if isinstance(key, (tuple, list)) and all(isinstance(k, str) for k in key):
            return self._subset(cols=list(key))
        if isinstance(key, (int, np.integer)):
            return _Row((k, c[key]) for k, c in self._cols.items())
        sel = np.arange(len(self))[key]
        return self._subset(rows=np.atleast_1d(sel))

    def __setitem__(self, key, values):
        values = list(values)
        if self._cols and len(values) != len(self):
            raise ValueError("column length mismatch")
        self._cols[key] = values
        self._types[key] = None

    def _subset(self, rows=None, cols=None):
        out = CandTable()
        out.meta = dict(self.meta)
        out.formats = dict(self.formats)
        for k in cols if cols is not None else self._cols:
            c = self._cols[k]
            out._cols[k] = list(c) if rows is None else [c[i] for i in rows]
            out._types[k] = self._types[k]
        return out

    def copy(self):
        return self._subset()

    # -- output ---------------------------------------------------------------------------------
    def _cell(self, k, v):
        fmt = self.formats.get(k)
        if fmt and isinstance(v, (float, np.floating)):
            return format(v, fmt)
        if isinstance(v, (float, np.floating)):
            return repr(float(v))
        return str(v)

    def _lines(self, sep, quote):
        yield sep.join(self._cols)
        for i in range(len(self)):
            cells = [self._cell(k, c[i]) for k, c in self._cols.items()]
            yield sep.join(quote(c) for c in cells)

    def write(self, path, overwrite=False, format=None):
        """``format="ascii"``: header + rows, space separated (astropy ``ascii.basic``); a ``.csv`` name or
        ``format="csv"``: comma separated.  Column formats are applied like astropy applies ``info.format``."""
        if os.path.exists(path) and not overwrite:
            raise OSError(f"File exists: {path}")
        csv = format == "csv" or (format is None and str(path).endswith(".csv"))
        sep = "," if csv else " "

        def quote(c):
            if c == "" or any(ch in c for ch in (sep, '"', "\n")):
                return '"' + c.replace('"', '""') + '"'
            return c

        with open(path, "w") as fobj:
            for line in self._lines(sep, quote):
                fobj.write(line + "\n")

    @classmethod
    def read(cls, path):
        """Read back a file written by :meth:`write` (numbers become floats/ints where they parse)."""
        import csv

        with open(path, newline="") as fobj:
            first = fobj.readline()
            fobj.seek(0)
            rows = list(csv.reader(fobj, delimiter="," if "," in first else " "))

        def conv(s):
            for t in (int, float):
                try:
                    return t(s)
                except ValueError:
                    pass
            return s

        cols = list(zip(*[[conv(c) for c in r] for r in rows[1:]])) if len(rows) > 1 else [[] for _ in rows[0]]
        out = cls()
        for k, c in zip(rows[0], cols):
            out._cols[k] = list(c)
            out._types[k] = None
        return out

    def __str__(self):
        cells = [[k] + [self._cell(k, v) for v in c] for k, c in self._cols.items()]
        widths = [max(len(s) for s in col) for col in cells]
        lines = []
        for i in range(len(self) + 1):
            lines.append(" ".join(col[i].rjust(w) for col, w in zip(cells, widths)))
            if i == 0:
                lines.append(" ".join("-" * w for w in widths))
        return "\n".join(lines)


# ---------------------------------------------------------------------------------------------
# peak boundaries
# ---------------------------------------------------------------------------------------------
def get_boundaries_from_level(x, y, level, x0):
    """Extent of the region around ``x0`` where ``y`` stays above ``level`` (hendrics/efsearch.py:1299-1332): walk
    outwards from the sample nearest to ``x0`` while ``y > level``; the last sample still above it is the limit (the
    first sample is never tested on the way down, like in the reference's ``while idx > 0``).  Pinned to the reference's
    own function on 64 seeded cases (tests/golden/boundaries.npz).

    >>> x = np.linspace(-10, 10, 1000)
    >>> y = np.sinc(x)**2 + np.sinc((x - 5))**2
    >>> vals = get_boundaries_from_level(x, y, 0.5, 0)
    >>> assert np.allclose(np.abs(vals), 0.44, atol=0.1)
    """
    x, y = np.asarray(x), np.asarray(y)
    centre = int(np.argmin(np.abs(x - x0)))
    above = y > level
    lo = hi = x0
    below_left = np.flatnonzero(~above[1:centre + 1])  # index 0 is never tested by the reference's `idx > 0`
    first = below_left[-1] + 2 if below_left.size else 1
    if first <= centre:
        lo = x[first]
    below_right = np.flatnonzero(~above[centre:])
    last = centre + (below_right[0] if below_right.size else above.size - centre) - 1
    if last >= centre:
        hi = x[last]
    return lo, hi


# For each of the 16 corner patterns of a grid cell (bit 0: corner (i, j), bit 1: (i, j+1), bit 2: (i+1, j+1),
# bit 3: (i+1, j)): the pairs of cell edges a level curve joins.  Edges: 0 = (i, j)-(i, j+1), 1 = (i, j+1)-(i+1, j+1),
# 2 = (i+1, j)-(i+1, j+1), 3 = (i, j)-(i+1, j).  The two saddles (5, 10) are resolved by the cell-centre value.
_CELL_EDGES = {
    1: ((3, 0),), 2: ((0, 1),), 3: ((3, 1),), 4: ((1, 2),), 6: ((0, 2),), 7: ((3, 2),), 8: ((2, 3),), 9: ((0, 2),),
    11: ((1, 2),), 12: ((1, 3),), 13: ((0, 1),), 14: ((3, 0),),
    5: ((3, 0), (1, 2)), 10: ((0, 1), (2, 3)),           # centre below the level: corners stay separated
    -5: ((3, 2), (0, 1)), -10: ((3, 0), (1, 2)),          # centre above the level: the high corners are connected
}


def trace_level_curves(x, y, image, level):
    """Level curves of ``image`` (2-D) at ``level`` on the grid whose node coordinates are the 2-D arrays ``x``, ``y``
    (any grid pcolormesh accepts): marching squares with linear interpolation along the cell edges, the coordinates
    interpolated with the same weights.  Returns a list of ``(n_k, 2)`` vertex arrays, one per connected curve
    (vertex order along the curve is not reconstructed: only the vertex set and its segments are needed here) and
    a parallel list of ``(m_k, 2, 2)`` segment arrays."""
    z = np.asarray(image, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.ndim == 1 and y.ndim == 1:
        x, y = np.meshgrid(x, y)
    if z.ndim != 2 or x.shape != z.shape or y.shape != z.shape:
        raise ValueError("x, y and image must be 2-D arrays of the same shape")
    ni, nj = z.shape
    if ni < 2 or nj < 2:
        return [], []
    hi = z > level
    code = (hi[:-1, :-1].astype(np.int8) | (hi[:-1, 1:].astype(np.int8) << 1) | (hi[1:, 1:].astype(np.int8) << 2)
            | (hi[1:, :-1].astype(np.int8) << 3))
    ci, cj = np.nonzero((code != 0) & (code != 15))
    if ci.size == 0:
        return [], []
    codes = code[ci, cj].astype(np.int64)
    saddle = (codes == 5) | (codes == 10)
    if saddle.any():
        centre = 0.25 * (z[ci, cj] + z[ci, cj + 1] + z[ci + 1, cj + 1] + z[ci + 1, cj])
        codes = np.where(saddle & (centre > level), -codes, codes)

    # a crossing point is identified by the grid edge it lies on: horizontal edges (i, j)-(i, j+1) get the id
    # i*nj + j, vertical edges (i, j)-(i+1, j) the id ni*nj + i*nj + j
    def edge_id(e, i, j):
        return np.where(e == 0, i * nj + j, np.where(e == 2, (i + 1) * nj + j,
                        np.where(e == 3, ni * nj + i * nj + j, ni * nj + i * nj + j + 1)))

    seg_a, seg_b = [], []
    for c, pairs in _CELL_EDGES.items():
        m = codes == c
        if not m.any():
            continue
        i, j = ci[m], cj[m]
        for ea, eb in pairs:
            seg_a.append(edge_id(np.full(i.shape, ea), i, j))
            seg_b.append(edge_id(np.full(i.shape, eb), i, j))
    seg_a, seg_b = np.concatenate(seg_a), np.concatenate(seg_b)

    ids, inv = np.unique(np.concatenate([seg_a, seg_b]), return_inverse=True)
    a, b = inv[:seg_a.size], inv[seg_a.size:]
    # coordinates of every crossing point
    vert = ids >= ni * nj
    k = np.where(vert, ids - ni * nj, ids)
    i0, j0 = k // nj, k % nj
    i1, j1 = np.where(vert, i0 + 1, i0), np.where(vert, j0, j0 + 1)
    z0, z1 = z[i0, j0], z[i1, j1]
    w = (level - z0) / (z1 - z0)
    px = x[i0, j0] + w * (x[i1, j1] - x[i0, j0])
    py = y[i0, j0] + w * (y[i1, j1] - y[i0, j0])
    pts = np.stack([px, py], axis=1)

    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    n = ids.size
    graph = coo_matrix((np.ones(a.size, dtype=np.int8), (a, b)), shape=(n, n))
    n_comp, labels = connected_components(graph, directed=False)
    order = np.argsort(labels, kind="stable")
    bounds = np.searchsorted(labels[order], np.arange(n_comp + 1))
    seg_lab = labels[a]
    seg_order = np.argsort(seg_lab, kind="stable")
    seg_bounds = np.searchsorted(seg_lab[seg_order], np.arange(n_comp + 1))
    curves, segments = [], []
    for c in range(n_comp):
        curves.append(pts[order[bounds[c]:bounds[c + 1]]])
        s = seg_order[seg_bounds[c]:seg_bounds[c + 1]]
        segments.append(np.stack([pts[a[s]], pts[b[s]]], axis=1))
    return curves, segments


def _distance2_to_segments(segs, x0, y0):
    """Smallest squared distance from (x0, y0) to a set of line segments (m, 2, 2)."""
    p, q = segs[:, 0], segs[:, 1]
    d = q - p
    len2 = np.einsum("ij,ij->i", d, d)
    rel = np.array([x0, y0]) - p
    with np.errstate(invalid="ignore", divide="ignore"):
        t = np.clip(np.where(len2 > 0, np.einsum("ij,ij->i", rel, d) / len2, 0.0), 0.0, 1.0)
    near = p + t[:, None] * d
    diff = near - np.array([x0, y0])
    return float(np.min(np.einsum("ij,ij->i", diff, diff)))


def get_xy_boundaries_from_level(x, y, image, level, x0, y0):
    """Bounding box of the level curve of ``image`` at ``level`` that passes nearest to the local maximum (x0, y0)
    (hendrics/efsearch.py:1262-1296, where it is ``contour`` + ``find_nearest_contour(..., pixel=False)``).
    Returns ``min_x, max_x, min_y, max_y``.

    >>> x = np.linspace(-10, 10, 1000)
    >>> y = np.linspace(-10, 10, 1000)
    >>> X, Y = np.meshgrid(x, y)
    >>> Z = np.sinc(np.sqrt(X**2 + Y**2))**2 + np.sinc(np.sqrt((X - 5)**2 + Y**2))**2
    >>> vals = get_xy_boundaries_from_level(X, Y, Z, 0.5, 0, 0)
    >>> assert np.allclose(np.abs(vals), 0.44, atol=0.1)
    """
    curves, segments = trace_level_curves(x, y, image, level)
    if not curves:
        raise ValueError(f"no level curve at {level} in the image")
    d2 = [_distance2_to_segments(s, x0, y0) for s in segments]
    best = curves[int(np.argmin(d2))]
    return best[:, 0].min(), best[:, 0].max(), best[:, 1].min(), best[:, 1].max()


# ---------------------------------------------------------------------------------------------
# the candidate table
# ---------------------------------------------------------------------------------------------
def print_qffa_results(best_cand_table):
    """hendrics/efsearch.py:1246-1259: the detected candidates with "amp ± err", or -- when nothing was detected --
    the best one with its 90 % upper limit."""
    table = best_cand_table.copy()
    detected = ~np.isnan(np.asarray(table["pulse_amp"], dtype=np.float64)) if len(table) else np.zeros(0, dtype=bool)
    if detected.any():
        print("Best candidate(s):")
        table["Pulsed amplitude (%)"] = [f"{a:g} ± {e:g}" for a, e in zip(table["pulse_amp"], table["pulse_amp_err"])]
        rows = np.flatnonzero(detected)
    else:
        print("No pulsations found. Best candidate and upper limit:")
        table["Pulsed amplitude (%)"] = [f"<{a:g} (90%)" for a in table["pulse_amp_ul_0.9"]]
        rows = np.arange(min(1, len(table)))
    print(table["mjd", "f", "fdot", "fddot", "power", "Pulsed amplitude (%)"][rows])


def _thresholds(ef):
    """Detection level, degrees of freedom, plot binning and axis label of a periodogram
    (hendrics/efsearch.py:1347-1373)."""
    ntrial = ef.stat.size
    if getattr(ef, "oversample", None) is not None:
        ntrial = int(ntrial / ef.oversample)
    if ef.kind == "Z2n":
        return dict(
            ndof=ef.N - 1,
            detlev=z2_n_detection_level(epsilon=0.001, n=int(ef.N), ntrial=ntrial, n_summed_spectra=int(ef.M)),
            nbin=max(16, ef.N * 8, ef.nbin if ef.nbin is not None else 1),
            label="$" + "Z^2_{" + f"{ef.N}" + "}$",
        )
    return dict(
        ndof=ef.nbin,
        detlev=fold_detection_level(nbin=int(ef.nbin), epsilon=0.001, ntrial=ntrial),
        nbin=max(16, ef.nbin),
        label=rf"$\chi^2_{ef.nbin}$ Stat",
    )


class _Peak:
    """One peak of the plane with its 1-sigma box and the two cuts through it."""

    def __init__(self, ef, idx):
        stat = ef.stat
        self.fdot_idx = self.fddot_idx = 0
        if stat.ndim > 1 and stat.shape[0] > 1:
            # NB the reference names the FIRST index of the (n_fdot-like, n_f-like) plane "f_idx" (:1437) and cuts
            # freq[f_idx, :] / freq[:, fdot_idx]; kept, since plot.py indexes the plane with these columns
            self.f_idx, self.fdot_idx = int(idx[0]), int(idx[1])
            self.f, self.fdot = ef.freq[self.f_idx, self.fdot_idx], ef.fdots[self.f_idx, self.fdot_idx]
            self.power = stat[self.f_idx, self.fdot_idx]
            self.cut_f = (ef.freq[self.f_idx, :], stat[self.f_idx, :])
            self.cut_fdot = (ef.freq[:, self.fdot_idx], stat[:, self.fdot_idx])
            self.level, self.sig_e1 = power_confidence_limits(self.power, c=0.68, n=ef.N)
            if self.power < self.level:
                self.box = (np.nan,) * 4
            else:
                self.box = get_xy_boundaries_from_level(ef.freq, ef.fdots, stat, self.level, self.f, self.fdot)
        elif stat.ndim == 1:
            self.f_idx = int(np.ravel(idx)[0])
            self.f, self.fdot = ef.freq[self.f_idx], 0
            self.power = stat[self.f_idx]
            self.cut_f, self.cut_fdot = (ef.freq, stat), None
            self.level, self.sig_e1 = power_confidence_limits(self.power, c=0.68, n=ef.N)
            self.box = get_boundaries_from_level(ef.freq, stat, self.level, self.f) + (0, 0)
        else:
            raise ValueError("Did not understand stats shape.")

    def amplitudes(self, ncounts, detlev, n):
        """(cl_0.9 power, amp, amp_err, amp_cl_0.1, amp_cl_0.9, amp_ul_0.9), in percent; NaN where not applicable."""
        sig_0, sig_1 = power_confidence_limits(self.power, c=0.90, n=n)
        pct = lambda s: a_from_ssig(s, ncounts) * 100  # noqa: E731
        if self.power < detlev:
            return sig_0, np.nan, np.nan, np.nan, np.nan, pct(sig_1)
        amp = pct(self.power)
        return sig_0, amp, pct(self.sig_e1) - amp, pct(sig_0), pct(sig_1), np.nan


def _analyze_qffa_results(input_ef_periodogram, fname=None):
    """The best five candidates of a search plane as a :class:`CandTable` (hendrics/efsearch.py:1335-1545): for each,
    the power, its frequency / fdot with the 1-sigma error box (contour of the plane at the lower 68 % confidence
    limit of the power), pulsed amplitude or its upper limit.  For each DETECTED candidate, when ``fname`` is given,
    the cuts through the peak are written to ``<root>_cand_<k>_fdot<fdot>.csv`` and ``<root>_cand_<k>_f<f>.dat``
    (k counts down from 4).  A candidate below the detection level gets a row (the upper limit) but no files."""
    ef = input_ef_periodogram
    if getattr(ef, "M", None) is None:
        ef.M = 1
    info = _thresholds(ef)
    fddot = ef.fddots if getattr(ef, "fddots", None) is not None else 0
    table = CandTable.for_candidates()
    root = fname.replace(HEN_FILE_EXTENSION, "") if fname is not None else None

    for rank, idx in enumerate(find_peaks_in_image(ef.stat, n=N_CANDIDATES)):
        peak = _Peak(ef, idx)
        if ef.ncounts is None:
            continue
        fmin, fmax, fdotmin, fdotmax = peak.box
        sig_0, amp, amp_err, amp_0, amp_1, amp_ul = peak.amplitudes(ef.ncounts, info["detlev"], ef.N)
        table.add_row([
            ef.filename, ef.pepoch, peak.power, peak.f, fmin - peak.f, fmax - peak.f, peak.fdot, fdotmin - peak.fdot,
            fdotmax - peak.fdot, fddot, sig_0, amp, amp_err, amp_0, amp_1, amp_ul, peak.f_idx, peak.fdot_idx,
            peak.fddot_idx,
        ])
        if peak.power < info["detlev"] or root is None:
            continue
        k = N_CANDIDATES - rank - 1
        CandTable({"freq": peak.cut_f[0], "stat": peak.cut_f[1]}).write(
            f"{root}_cand_{k}_fdot{peak.fdot}.csv", overwrite=True, format="ascii")
        if peak.cut_fdot is not None:
            CandTable({"fdot": peak.cut_fdot[0], "stat": peak.cut_fdot[1]}).write(
                f"{root}_cand_{k}_f{peak.f}.dat", overwrite=True, format="ascii")

    print_qffa_results(table)
    table.meta.update(info, filename=None)
    if getattr(ef, "filename", None) is not None and os.path.exists(ef.filename):
        table.meta["filename"] = ef.filename
    return table


def analyze_qffa_results(fname):
    """hendrics/efsearch.py:1548-1565: load a saved periodogram, tabulate its candidates, write
    ``<fname>_best_cands.csv``.  Returns ``(EFPeriodogram, CandTable)``."""
    ef = load_folding(fname)
    table = _analyze_qffa_results(ef, fname=fname)
    table.write(fname + "_best_cands.csv", overwrite=True)
    return ef, table
