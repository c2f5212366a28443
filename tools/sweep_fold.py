"""Sweep (R, W) of K1 on the benchmark shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device

def run(n_ev, n_f, n_fd, nbin, nharm, R, W, reps=2):
    lib = _cabi.lib()
    lib.hx_fold_set_tuning(R, W)
    rng = np.random.default_rng(0)
    ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, n_ev))).cuda()
    f = torch.from_numpy(7.4 + 5e-6 * np.arange(n_f)).cuda()
    fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd) if n_fd > 1 else np.zeros(1)).cuda()
    try:
        fold_stats_device(ts, f, fd, nbin, nharm)
    except Exception as e:
        print(f"R={R} W={W}: {e}"); return
    best = 1e30
    for _ in range(reps):
        fold_stats_device(ts, f, fd, nbin, nharm)
        ms, grid, block, smem = _cabi.fold_last_kernel()
        best = min(best, ms)
    evtr = n_ev * n_f * n_fd
    print(f"N={n_ev:.0e} nf={n_f} nfd={n_fd} nbin={nbin} R={R} W={W:2d} block={block} grid={grid} smem={smem}: "
          f"{best:8.2f} ms {evtr/best*1e3:.3e} ev*tr/s ({evtr/best*1e3/148/1.965e9:.2f}/SM/clk)", flush=True)

which = sys.argv[1]
if which == "c2":
    for R, Ws in ((8, (0, 4, 6, 8, 13)), (4, (0, 4, 8, 12, 16)), (2, (0, 8, 16))):
        for W in Ws:
            run(1_000_000, 1_000_000, 1, 16, 0, R, W, reps=1)
elif which == "c3":
    for R, Ws in ((1, (0, 13, 12, 8, 6)), (2, (0, 6, 4, 3)), (4, (0, 3))):
        for W in Ws:
            run(2_000_000, 2048, 64, 128, 4, R, W)
elif which == "mid":
    for nbin in (32, 64):
        for R in (1, 2, 4, 8):
            run(1_000_000, 100_000, 1, nbin, 2, R, 0)
            run(1_000_000, 2048, 32, nbin, 2, R, 0)
    for R in (1, 2):
        run(1_000_000, 100_000, 1, 128, 2, R, 0)
    for R in (4, 8):
        run(1_000_000, 2048, 64, 16, 2, R, 0)
