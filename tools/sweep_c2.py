"""(R, W) sweep of the f-only kernel on the C2 shape (nbin = 16) and nbin = 32."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device
lib = _cabi.lib()
rng = np.random.default_rng(2)
ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, 1_000_000))).cuda()
fd = torch.zeros(1, dtype=torch.float64).cuda()
for nbin, n_f in ((16, 400_000), (32, 200_000)):
    f = torch.from_numpy(7.4 + 5e-6 * np.arange(n_f)).cuda()
    ref = None
    for R, W in ((0, 0), (4, 4), (4, 6), (4, 8), (4, 10), (4, 12), (8, 2), (8, 3), (8, 4), (8, 5), (8, 6), (8, 8), (2, 8), (2, 12)):
        lib.hx_fold_set_tuning(R, W)
        try:
            st = fold_stats_device(ts, f, fd, nbin, 0)
            best = 1e9
            for _ in range(2):
                st = fold_stats_device(ts, f, fd, nbin, 0); best = min(best, _cabi.fold_last_kernel()[0])
            k = _cabi.fold_last_kernel(); sh = _cabi.fold_last_kernel_shape()
            crc = zlib.crc32(st.cpu().numpy().tobytes()); ref = ref or crc
            print(f"nbin={nbin} R={R} W={W}: {best:8.2f} ms {1e6*n_f/best/1e-3:.3e}  block={k[2]} smem={k[3]} R={sh['R']} {'ok' if crc == ref else 'MISMATCH'}", flush=True)
        except Exception as e:
            print(f"nbin={nbin} R={R} W={W}: {e}", flush=True)
lib.hx_fold_set_tuning(0, 0)
