"""One GPU's share of config 3 through the public API: Z^2_4, nbin=128, 1e7 events x 125000 frequencies x
64 fdots (= 1/8 of the 1e6-frequency grid), followed by the top-10 candidate search on the 8e6 statistics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import z_n_search, _cabi
from hendrics_b200.base import find_peaks_in_image
from hendrics_b200.fake import simulate_events

times = simulate_events(10_000_000, 1e5, seed=3) + 8e7
n_f = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
freqs = 9.85 + 5e-6 * np.arange(n_f)          # the slice that contains the injected pulse (9.90099 Hz)
fdots = -1.6e-9 + 5e-11 * np.arange(64)
torch.cuda.synchronize()
t0 = time.perf_counter()
fg, fdg, stats = z_n_search(times, freqs, nharm=4, nbin=128, fdots=fdots)
torch.cuda.synchronize()
t1 = time.perf_counter()
peaks = find_peaks_in_image(stats, n=10)
t2 = time.perf_counter()
evtr = (times.size - 1) * freqs.size * fdots.size
print(f"C3 share: {times.size} events x {freqs.size} f x {fdots.size} fdot = {evtr:.3e} event*trials")
print(f"search {t1-t0:.2f} s -> {evtr/(t1-t0):.3e} event*trials/s end to end (kernel {_cabi.fold_last_kernel()[0]/1e3:.2f} s)")
print(f"top-10 of {stats.size} statistics in {t2-t1:.3f} s; best Z^2_4 = {stats[tuple(peaks[0])]:.1f} at f = {fg[tuple(peaks[0])]:.6f} Hz, fdot = {fdg[tuple(peaks[0])]:.2e}")
print("stats finite:", bool(np.isfinite(stats).all()), " mean over noise trials ~ 2n = 8:", float(np.median(stats)))
