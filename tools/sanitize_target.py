"""Small end-to-end invocations of every kernel: target for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hendrics_b200 import fold_profiles, z_n_search, epoch_folding_search, z_n_search_unbinned, base
from hendrics_b200.efsearch import search_with_qffa, search_with_qffa_step
from hendrics_b200.fake import simulate_events

ts = simulate_events(9001, 300.0, seed=1)
f = 9.85 + 0.1 * np.arange(70) / 70
for nbin, fd in ((16, [0.0]), (128, [0.0, 1e-6]), (10, [1e-6]), (256, [0.0]), (64, [0.0, 2e-6, 3e-6])):
    fold_profiles(ts, f, fd, nbin=nbin, nharm=2)
    z_n_search(ts, f, nharm=2, nbin=nbin, fdots=fd)
fold_profiles(ts, f, [1e-6], nbin=128, nharm=2, exact_phase=True)
epoch_folding_search(ts, 5.0 + 10 * np.arange(30011) / 30011, nbin=16)      # fused epilogue, ragged
epoch_folding_search(ts[:1025], f, nbin=32, segment_size=100.0)
z_n_search_unbinned(ts, f, nharm=3, fdots=[0.0, 1e-6])
np.random.seed(0)
search_with_qffa(ts + 8e7, 9.85, 9.95, nbin=16, n=2, silent=True)
tc = ts - (ts.max() + ts.min()) / 2
search_with_qffa_step(tc, 9.9, 1e-7, 1e-11, nbin=64, nprof=100, n=3)
search_with_qffa_step(tc, 3.3, nbin=10, nprof=37, n=1)
base.GPU_TOPK_MIN_SIZE = 10
base.find_peaks_in_image(np.random.default_rng(0).normal(size=(30, 40)), n=5)
print("sanitize target ok")
