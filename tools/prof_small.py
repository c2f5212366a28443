"""Host-side profile of a small search (C1 shape: 1e5 events x 1e4 frequencies, Z^2_2 nbin=128) through the public API."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi, z_n_search
from hendrics_b200.fake import simulate_events
ts = simulate_events(100_000, 1e5, seed=1) + 8e7
freqs = 9.876 + 5e-6 * np.arange(10_000)
for _ in range(3):
    z_n_search(ts, freqs, nharm=2, nbin=128)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    z_n_search(ts, freqs, nharm=2, nbin=128)
torch.cuda.synchronize()
print(f"wall per search {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms; kernel {_cabi.fold_last_kernel()[0]:.3f} ms; launches per search {_cabi.launch_count()}")
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    z_n_search(ts, freqs, nharm=2, nbin=128)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
