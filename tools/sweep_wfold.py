"""Sweep K1w's (trials/thread, events/group, warps/CTA) on B200."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import wfold_stats_device

def run(n_ev, n_f, n_fd, nbin, weights, R, U, W, gen=0, reps=2):
    lib = _cabi.lib(); lib.hx_wfold_set_tuning(R, U, W, gen)
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 1e5, n_ev))
    ts = torch.from_numpy(t).cuda()
    w = None
    if weights == "int": w = torch.from_numpy(rng.poisson(3.0, n_ev).astype(np.float64)).cuda()
    if weights == "float": w = torch.from_numpy(rng.uniform(0.5, 1.5, n_ev)).cuda()
    f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
    fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd) if n_fd > 1 else np.zeros(1)).cuda()
    wfold_stats_device(ts, w, 1.0, f, fd, nbin, 2)
    best = 1e30
    for _ in range(reps):
        wfold_stats_device(ts, w, 1.0, f, fd, nbin, 2)
        ms, cell, _ = _cabi.wfold_last_kernel(); best = min(best, ms*1e-3)
    u = n_ev*n_f*n_fd
    print(f"nbin={nbin} nfd={n_fd} w={weights} cell={cell} R={R} U={U} W={W} gen={gen}: {best*1e3:8.2f} ms {u/best:.3e} ev*tr/s ({u/best/148/1.965e9:.2f}/SM/clk)", flush=True)

for nbin in (16, 64, 128):
    for R, U in ((1, 2), (1, 4), (2, 2), (2, 4)):
        run(1_000_000, 20_000, 1, nbin, "float", R, U, 0)
for nbin in (16, 64, 128, 256):
    for R in (1, 2, 4):
        run(1_000_000, 50_000, 1, nbin, "int", R, 0, 0)
        run(1_000_000, 50_000, 1, nbin, None, R, 0, 0)
run(1_000_000, 20_000, 1, 64, "float", 0, 0, 0, gen=1)
run(1_000_000, 20_000, 1, 64, "int", 0, 0, 0, gen=1)
run(1_000_000, 20_000, 1, 100, "float", 0, 0, 0)
run(1_000_000, 20_000, 1, 100, "int", 0, 0, 0)
run(1_000_000, 2_000, 8, 64, "float", 0, 0, 0)
run(1_000_000, 2_000, 8, 64, "int", 0, 0, 0)
for W in (2, 4, 8):
    run(1_000_000, 50_000, 1, 16, "int", 4, 0, W)
    run(1_000_000, 50_000, 1, 64, "int", 2, 0, W)
