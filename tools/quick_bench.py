"""Developer timing of K1 on a few shapes (CUDA events, device-resident inputs)."""
import ctypes
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device


def run(n_ev, n_f, n_fd, nbin, nharm, reps=3, R=0, W=0):
    lib = _cabi.lib()
    lib.hx_fold_set_tuning(R, W)
    rng = np.random.default_rng(0)
    ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, n_ev))).cuda()
    f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
    fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd) if n_fd > 1 else np.zeros(1)).cuda()
    fold_stats_device(ts, f, fd, nbin, nharm)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fold_stats_device(ts, f, fd, nbin, nharm)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    evtr = n_ev * n_f * n_fd
    print(f"N={n_ev:.0e} nf={n_f} nfd={n_fd} nbin={nbin} nharm={nharm} R={R} W={W}: {best*1e3:9.2f} ms  "
          f"{evtr/best:.3e} ev*tr/s  ({evtr/best/148/1.965e9:.2f} /SM/clk@max)", flush=True)


if __name__ == "__main__":
    lib = _cabi.lib()
    lib.hx_fold_set_tuning.argtypes = [ctypes.c_int, ctypes.c_int]
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "c2"):
        for R in (1, 2, 4):
            run(1_000_000, 200_000, 1, 16, 0, R=R)
    if which in ("all", "c3"):
        for R in (1, 2, 4):
            run(1_000_000, 4096, 64, 128, 4, R=R)
    if which in ("all", "c1"):
        run(100_000, 10_000, 1, 128, 2)
        run(1_000_000, 100_000, 1, 128, 2)
        run(1_000_000, 100_000, 1, 64, 2)
        run(1_000_000, 100_000, 1, 32, 2)
