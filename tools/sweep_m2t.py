"""(R=2, W) sweep of the (f, fdot) kernel with the per-CTA fdot-term tile (M2T) at nbin = 64, 128."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device
lib = _cabi.lib()
rng = np.random.default_rng(0)
ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, 2_000_000))).cuda()
for nbin in (64, 128):
    for n_f, n_fd in ((4096, 32), (2048, 64), (20000, 8)):
        f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
        fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd)).cuda()
        for R, W in ((0, 0), (2, 2), (2, 3), (2, 4), (2, 6), (2, 8), (2, 12)):
            lib.hx_fold_set_tuning(R, W)
            try:
                fold_stats_device(ts, f, fd, nbin, 2)
                best = 1e9
                for _ in range(2):
                    fold_stats_device(ts, f, fd, nbin, 2); best = min(best, _cabi.fold_last_kernel()[0])
                k = _cabi.fold_last_kernel()
                print(f"nbin={nbin} nf={n_f} nfd={n_fd} R={R} W={W}: {best:8.2f} ms {2e6*n_f*n_fd/best/1e-3:.3e}  block={k[2]} smem={k[3]}", flush=True)
            except Exception as e:
                print(f"nbin={nbin} R={R} W={W}: {e}")
lib.hx_fold_set_tuning(0, 0)
