"""Synthetic HENfake-style event lists for benchmarks and tests.

The reference's simulator (``hendrics/fake.py:623-766``, ``EventList.simulate_times``) needs
stingray + astropy FITS; this restates its recipe on plain arrays: N sorted fp64 arrival times on
[tstart, tstop) drawn from a rate proportional to 1 + amp*cos(2 pi f0 t) by thinning (the
fixture of ``hendrics/tests/test_efsearch.py:46-62`` uses 400 + 100 cos, i.e. amp = 0.25)."""

from __future__ import annotations

import numpy as np

PULSE_FREQUENCY = 1 / 0.101
MET_OFFSET = 8e7  # NuSTAR-like absolute mission elapsed time, removed by folding_search


def simulate_events(n_events, tstop, f0=PULSE_FREQUENCY, amp=0.25, seed=1, tstart=0.0):
    rng = np.random.default_rng(seed)
    chunks, have = [], 0
    while have < n_events:
        m = int((n_events - have) * (1 + amp) * 1.05) + 16
        t = rng.uniform(tstart, tstop, m)
        if amp > 0:
            keep = rng.uniform(0, 1 + amp, m) < 1 + amp * np.cos(2 * np.pi * f0 * t)
            t = t[keep]
        chunks.append(t)
        have += t.size
    out = np.concatenate(chunks)[:n_events]
    out.sort()
    return out
