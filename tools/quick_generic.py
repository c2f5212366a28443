"""How much slower is the generic (non power-of-two nbin) path of K1 than the pre-scaled fast path?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device
rng = np.random.default_rng(0)
ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, 2_000_000))).cuda()
for nbin, n_f, n_fd, nharm in ((16, 100000, 1, 0), (20, 100000, 1, 0), (24, 100000, 1, 0), (32, 100000, 1, 0), (50, 50000, 1, 2), (64, 50000, 1, 2),
                               (100, 20000, 1, 2), (128, 20000, 1, 2), (100, 2048, 16, 2), (128, 2048, 16, 2)):
    f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
    fd = torch.from_numpy((-1.6e-9 + 5e-11 * np.arange(n_fd)) if n_fd > 1 else np.zeros(1)).cuda()
    fold_stats_device(ts, f, fd, nbin, nharm)
    best = 1e9
    for _ in range(3):
        fold_stats_device(ts, f, fd, nbin, nharm); best = min(best, _cabi.fold_last_kernel()[0])
    k = _cabi.fold_last_kernel(); sh = _cabi.fold_last_kernel_shape()
    print(f"nbin={nbin:4d} nf={n_f:6d} nfd={n_fd:3d}: {best:8.2f} ms {2e6*n_f*n_fd/best/1e-3:.3e} ev*tr/s  mode={sh['mode']} R={sh['R']} block={k[2]} smem={k[3]}", flush=True)
