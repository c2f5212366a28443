"""Orbital-parameter trial search: the host side of ``hx_orbital_search_dev``.

    decide_binary_parameters   hendrics/efsearch.py:208-260   random (freq, fdot, X = a sin i / c, Porb) trial rows
    folding_orbital_search     hendrics/efsearch.py:263-310   per row: scan the epoch of ascending node T0

The reference demodulates the event times on the CPU (two fixed-point passes of ``t - X sin(2 pi (t - T0) / Porb)``)
and calls a single-trial search for every T0.  Here all T0s of a row are ONE kernel launch.  The random draws come from
numpy's global generator in the same order as the reference's, so a seeded run visits the same parameter rows.
"""

from __future__ import annotations

import logging
import os

import numpy as np

from .search import epoch_folding_search, z_n_search

log = logging.getLogger("hendrics_b200")

TWOPI = 2 * np.pi
D_OMEGA_FACTOR = 2 * np.sqrt(3)
PARAMETER_COLUMNS = ["freq", "fdot", "X", "Porb", "done", "max_stat", "min_stat", "best_T0"]


def _save_df_to_csv(df, csv_file, reset=False):
    """Append ``df`` to ``csv_file`` (header only when the file is created or ``reset``): hendrics/efsearch.py:189-196."""
    fresh = reset or not os.path.exists(csv_file)
    df.to_csv(csv_file, header=fresh, index=False, mode="w" if fresh else "a")


def _n_steps(value_range, resolution):
    return int(np.diff(value_range)[0] // resolution) + 1


def _draw_block(length, freq_range, fdot_range, asini_range, omega_range):
    """One (freq, fdot) draw with its X values and, per X, its orbital angular frequencies; the sampling densities
    are the reference's: dX = 1 / (2 pi f), dOmega = 2 sqrt(3) / (2 pi f X length)."""
    freq = np.random.uniform(*freq_range)
    fdot = np.random.uniform(*fdot_range)
    rows = []
    for X in np.random.uniform(asini_range[0], asini_range[1], _n_steps(asini_range, 1 / (TWOPI * freq))):
        d_omega = D_OMEGA_FACTOR / (TWOPI * freq * X * length)
        for omega in np.random.uniform(omega_range[0], omega_range[1], _n_steps(omega_range, d_omega)):
            rows.append([freq, fdot, X, TWOPI / omega, False, 0.0, 0.0, 0.0])
    return rows


def decide_binary_parameters(length, freq_range, porb_range, asini_range, fdot_range=[0, 0], NMAX=10,
                             csv_file="db.csv", reset=False):
    """Write ``NMAX`` blocks of random orbital trial rows to ``csv_file`` (hendrics/efsearch.py:208-260)."""
    import pandas as pd

    omega_range = [1 / porb_range[1], 1 / porb_range[0]]
    log.info(f"Recommended frequency steps: {_n_steps(freq_range, 1 / length)}")
    for block in range(NMAX):
        rows = _draw_block(length, freq_range, fdot_range, asini_range, omega_range)
        # only the first block may truncate an existing file
        _save_df_to_csv(pd.DataFrame(rows, columns=PARAMETER_COLUMNS), csv_file, reset=reset and block == 0)
    return csv_file


def _t0_statistics(times, row, T0s, fun, fun_kwargs):
    """Statistic of the search at (freq, fdot) for every trial epoch T0 of one parameter row."""
    from .search import orbital_fold_stats

    freq, fdot, X, Porb = row
    if fun in (epoch_folding_search, z_n_search) and set(fun_kwargs) <= {"nbin", "nharm"}:
        nharm = fun_kwargs.get("nharm", 4) if fun is z_n_search else 0
        return orbital_fold_stats(times, freq, fdot, X, Porb, T0s, nbin=fun_kwargs.get("nbin", 128), nharm=nharm)
    # any other search function: T0 by T0, as the reference does
    out = []
    for T0 in T0s:
        demod = times - X * np.sin(TWOPI * (times - T0) / Porb)
        demod = demod - X * np.sin(TWOPI * (demod - T0) / Porb)
        out.append(fun(demod, np.array([freq]), fdots=fdot, **fun_kwargs)[1][0])
    return np.asarray(out)


def _extremes(T0s, stats):
    """The reference's running max / min (:298-302): start values 0 and 1e32, strict comparisons, NaNs never win."""
    stats = np.asarray(stats, dtype=np.float64)
    finite = ~np.isnan(stats)
    best = (0, None)
    if np.any(finite & (stats > 0)):
        i = int(np.nanargmax(np.where(finite, stats, -np.inf)))
        best = (stats[i], T0s[i])
    low = 1e32
    if np.any(finite & (stats < 1e32)):
        low = float(np.nanmin(stats))
    return best[0], low, best[1]


def folding_orbital_search(events, parameter_csv_file, chunksize=100, outfile="out.csv", fun=epoch_folding_search,
                           **fun_kwargs):
    """For every row of ``parameter_csv_file`` not yet done, scan T0 over one orbit in steps of
    ``min(Porb / (4 pi^2 f X), Porb / 10)`` (drawn uniformly at random, like the reference) and record the largest and
    smallest statistic and the best T0 in ``outfile`` (hendrics/efsearch.py:263-310).

    As in the reference, chunks whose index does not start at 0 are skipped (``chunk["done"][0]`` raises for them,
    :274-277)."""
    import pandas as pd

    times = (events.time - events.gti[0, 0]).astype(np.float64)
    for chunk in pd.read_csv(parameter_csv_file, chunksize=chunksize):
        if 0 not in chunk.index:
            continue
        col = {name: chunk.columns.get_loc(name) for name in ("max_stat", "min_stat", "best_T0", "done")}
        for pos in range(len(chunk)):
            if chunk["done"][pos]:
                continue
            row = chunk.iloc[pos][["freq", "fdot", "X", "Porb"]].to_numpy(dtype=np.float64)
            freq, _, X, Porb = row
            dT0 = min(1 / (TWOPI**2 * freq) * Porb / X, Porb / 10)
            T0s = np.random.uniform(0, Porb, int(Porb // dT0 + 1))
            hi, lo, best_T0 = _extremes(T0s, _t0_statistics(times, row, T0s, fun, fun_kwargs))
            at = chunk.index[pos]
            chunk.iloc[at, col["max_stat"]] = hi
            chunk.iloc[at, col["min_stat"]] = lo
            chunk.iloc[at, col["best_T0"]] = best_T0
            chunk.iloc[at, col["done"]] = True
        _save_df_to_csv(chunk, outfile)
