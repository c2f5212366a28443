"""Where does the end-to-end (host arrays in/out) time go?  cProfile + kernel time."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi, epoch_folding_search
from hendrics_b200.fake import simulate_events

times = simulate_events(1_000_000, 1e5, seed=2)
freqs = 7.4 + 5e-6 * np.arange(1_000_000)
for _ in range(2):
    epoch_folding_search(times, freqs, nbin=16)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(3):
    epoch_folding_search(times, freqs, nbin=16)
    print("kernel ms", _cabi.fold_last_kernel())
pr.disable()
torch.cuda.synchronize()
print("wall per call", (time.perf_counter() - t0) / 3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
