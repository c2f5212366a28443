"""(R, W) sweep of the (f, fdot) kernel, nbin = 32 / 64 / 128; statistics checked against the first combo.

Written for the 3-trials-per-thread experiment (profiles/r01_sweep_r3.txt): the R = 3 instantiations were removed again
(never better than R = 2), so `hx_fold_set_tuning(3, W)` now means "default R"; R = 2 / 4 / 8 combinations still sweep."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device
lib = _cabi.lib()
rng = np.random.default_rng(0)
ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, 2_000_000))).cuda()
combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(0, 0), (2, 4), (3, 2), (3, 3), (3, 4), (3, 8), (4, 2), (4, 3)]
for nbin, nharm in ((128, 4), (64, 2), (32, 2)):
    for n_f, n_fd in ((2048, 64), (20000, 8)):
        f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
        fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd)).cuda()
        ref = None
        for R, W in combos:
            lib.hx_fold_set_tuning(R, W)
            try:
                st = fold_stats_device(ts, f, fd, nbin, nharm)
                best = 1e9
                for _ in range(2):
                    st = fold_stats_device(ts, f, fd, nbin, nharm)
                    best = min(best, _cabi.fold_last_kernel()[0])
                k = _cabi.fold_last_kernel()
                sh = _cabi.fold_last_kernel_shape()
                crc = zlib.crc32(st.cpu().numpy().tobytes())
                ref = ref or crc
                print(f"nbin={nbin} nf={n_f} nfd={n_fd} R={R} W={W}: {best:8.2f} ms {2e6*n_f*n_fd/best/1e-3:.3e}  grid={k[1]} "
                      f"block={k[2]} smem={k[3]} R={sh['R']} m2t={sh['m2t']} {'ok' if crc == ref else 'MISMATCH'}", flush=True)
            except Exception as e:
                print(f"nbin={nbin} R={R} W={W}: {e}", flush=True)
lib.hx_fold_set_tuning(0, 0)
