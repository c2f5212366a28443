"""How far is the orbital-demodulation fold from the oracle?  Prints, per test case, the moved counts per profile and
the largest relative difference of the statistics, for the oracle with numpy's sin and with a correctly rounded sin."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import hx_oracle as O
from hendrics_b200.search import orbital_fold_stats

f0, X, Porb, T0 = 1.2345, 3.0, 5000.0, 1234.0
for nbin, nharm, fdot in [(16, 0, 0.0), (64, 2, 0.0), (32, 1, 2e-9), (20, 0, -1e-9)]:
    te = O.simulate_events(40001, 20000.0, f0=f0, amp=0.8, seed=nbin)
    ts = np.sort(te + X * np.sin(2 * np.pi * (te - T0) / Porb))
    T0s = np.concatenate([[T0], np.random.default_rng(1).uniform(0, Porb, 70)])
    stats, prof = orbital_fold_stats(ts, f0, fdot, X, Porb, T0s, nbin=nbin, nharm=nharm, return_profiles=True)
    fun = O.epoch_folding_search if nharm == 0 else O.z_n_search
    kw = dict(nbin=nbin) if nharm == 0 else dict(nbin=nbin, nharm=nharm)
    for name, sin in (("np.sin", np.sin), ("CR sin", O.sin_correctly_rounded)):
        ref_s, ref_p = O.orbital_stats(ts, f0, fdot, X, Porb, T0s, fun=fun, return_profiles=True, sin=sin, **kw)
        moved = np.abs(prof - ref_p).sum(axis=1)
        print(f"nbin={nbin} nharm={nharm} fdot={fdot}: vs {name}: moved per profile max {moved.max()} total {moved.sum()} "
              f"profiles touched {np.count_nonzero(moved)}/{len(T0s)}; stats max rel diff {np.max(np.abs(stats / ref_s - 1)):.3g}")
        bad = np.flatnonzero(moved)
        for k in bad[:3]:
            d = prof[k].astype(np.int64) - ref_p[k]
            print("   T0 idx", k, "diff bins", np.flatnonzero(d), d[np.flatnonzero(d)])
