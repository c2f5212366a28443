"""K1 (R, W, packing) sweep at nbin = 16, 32, 64: are the table entries / the packing threshold right?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device
lib = _cabi.lib()
rng = np.random.default_rng(0)
ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, 1_000_000))).cuda()
for nbin in (16, 32, 64):
    for n_f, n_fd in ((200_000, 1), (4096, 32)):
        f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
        fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd) if n_fd > 1 else np.zeros(1)).cuda()
        for pack in ((0, 1) if nbin >= 32 else (0,)):
            for R, W in ((0, 0), (2, 2), (2, 4), (2, 6), (2, 8), (4, 2), (4, 3), (4, 4), (4, 6), (4, 8), (8, 2), (8, 4)):
                lib.hx_fold_set_packing(pack); lib.hx_fold_set_tuning(R, W)
                try:
                    fold_stats_device(ts, f, fd, nbin, 0)
                    best = 1e9
                    for _ in range(2):
                        fold_stats_device(ts, f, fd, nbin, 0); best = min(best, _cabi.fold_last_kernel()[0])
                    print(f"nbin={nbin} nf={n_f} nfd={n_fd} pack={pack} R={R} W={W}: {best:8.2f} ms {1e6*n_f*n_fd/best/1e-3:.3e}", flush=True)
                except Exception as e:
                    print(f"nbin={nbin} pack={pack} R={R} W={W}: {e}")
lib.hx_fold_set_packing(-1); lib.hx_fold_set_tuning(0, 0)
