"""Exhaustive (R, W) sweep per (nbin, has_fd): prints the best few configurations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device

lib = _cabi.lib()
rng = np.random.default_rng(0)
n_ev = 400_000
ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, n_ev))).cuda()
import sys as _s
pack = int(_s.argv[1]) if len(_s.argv) > 1 else -1
lib.hx_fold_set_packing(pack)
for nbin in ((128, 256, 512) if pack == 1 else (8, 16, 32, 64, 128, 256)):
    for has_fd in (0, 1):
        n_f, n_fd = (262144, 1) if not has_fd else (4096, 64)
        f = torch.from_numpy(7.4 + 5e-6 * np.arange(n_f)).cuda()
        fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd) if n_fd > 1 else np.zeros(1)).cuda()
        res = []
        for R in (1, 2, 4, 8):
            for W in (2, 3, 4, 6, 8, 10, 12, 13, 16):
                lib.hx_fold_set_tuning(R, W)
                try:
                    fold_stats_device(ts, f, fd, nbin, 2)
                    fold_stats_device(ts, f, fd, nbin, 2)
                except Exception:
                    continue
                ms, grid, block, smem = _cabi.fold_last_kernel()
                if block != W * 32:
                    continue
                res.append((n_ev * n_f * n_fd / ms * 1e3 / 148 / 1.965e9, R, W, grid, smem))
        res.sort(reverse=True)
        print(f"nbin={nbin} fd={has_fd}: " + "  ".join(f"[R{r} W{w} g{g} {v:.1f}]" for v, r, w, g, s in res[:7]), flush=True)
lib.hx_fold_set_tuning(0, 0)
