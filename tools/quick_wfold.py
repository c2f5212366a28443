"""Developer timing of K1w (weighted / exposure-corrected fold) on a few shapes; CUDA events, device-resident inputs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from hendrics_b200 import _cabi
from hendrics_b200.search import wfold_stats_device


def run(n_ev, n_f, n_fd, nbin, nharm, weighted=True, n_gti=0, reps=3):
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 1e5, n_ev))
    ts = torch.from_numpy(t).cuda()
    w = torch.from_numpy(rng.poisson(3.0, n_ev).astype(np.float64)).cuda() if weighted else None
    f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
    fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd) if n_fd > 1 else np.zeros(1)).cuda()
    gti = None
    if n_gti:
        e = np.linspace(0, 1e5, n_gti + 1)
        gti = np.stack([e[:-1], e[1:] - 1.0], axis=1)
    kw = dict(gti=gti, start_time=t[0], stop_time=t[-1] + abs(t[1] - t[0]))
    wfold_stats_device(ts, w, 1.0, f, fd, nbin, nharm, **kw)
    torch.cuda.synchronize()
    best, bestk = 1e30, 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        wfold_stats_device(ts, w, 1.0, f, fd, nbin, nharm, **kw)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
        bestk = min(bestk, _cabi.wfold_last_kernel()[0] * 1e-3)
    evtr = n_ev * n_f * n_fd
    print(f"N={n_ev:.0e} nf={n_f} nfd={n_fd} nbin={nbin} nharm={nharm} w={weighted} gti={n_gti}: {best*1e3:9.2f} ms "
          f"(fold kernel {bestk*1e3:9.2f} ms)  {evtr/best:.3e} ev*tr/s  ({evtr/bestk/148/1.965e9:.2f} /SM/clk kernel)", flush=True)


if __name__ == "__main__":
    run(1_000_000, 20_000, 1, 16, 0)
    run(1_000_000, 20_000, 1, 64, 2)
    run(1_000_000, 20_000, 1, 64, 2, weighted=False, n_gti=10)
    run(1_000_000, 20_000, 1, 128, 2)
    run(1_000_000, 2_000, 8, 64, 2)
    run(1_000_000, 20_000, 1, 100, 2)
    run(100_000, 100_000, 1, 64, 0, weighted=False, n_gti=100)
    # sparse light curve: 1e7 bins, 1 % occupied
    rng = np.random.default_rng(1)
    n = 10_000_000
    t = (np.arange(n) + 0.5) * 1e-2
    w = (rng.random(n) < 0.01).astype(np.float64)
    ts, wd = torch.from_numpy(t).cuda(), torch.from_numpy(w).cuda()
    f = torch.from_numpy(9.0 + 5e-6 * np.arange(20000)).cuda(); fd = torch.zeros(1, dtype=torch.float64).cuda()
    for keep in (0, 4):
        _cabi.lib().hx_wfold_set_tuning(0, 0, 0, keep)
        wfold_stats_device(ts, wd, 1.0, f, fd, 64, 2)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        wfold_stats_device(ts, wd, 1.0, f, fd, 64, 2)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"sparse light curve 1e7 bins x 2e4 f, 1% occupied, keep_zeros={keep != 0}: {dt*1e3:8.2f} ms, folded {_cabi.wfold_last_kernel()[2]} events", flush=True)
    _cabi.lib().hx_wfold_set_tuning(0, 0, 0, 0)
