"""K1 on the C5 shape (1e8 events, Z^2_2, nbin 128, f only) for the per-GPU frequency counts of 1/2/4/8-way strong scaling:
kernel time, launch shape and efficiency relative to the 1e5-frequency run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.fake import simulate_events
from hendrics_b200.search import fold_stats_device

n_ev = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
ts = simulate_events(n_ev, 1e8, seed=5)
ts = ts - ts[0]; ts = ts[ts < ts[-1]]
td = torch.from_numpy(ts).cuda()
fd = torch.zeros(1, dtype=torch.float64, device="cuda")
base = None
for n_f in (100_000, 50_000, 25_000, 12_500, 12_544, 6_250):
    f = torch.from_numpy(9.9009 + 5e-9 * np.arange(n_f)).cuda()
    for _ in range(2):
        fold_stats_device(td, f, fd, 128, 2)
    ms, grid, block, smem = _cabi.fold_last_kernel()
    rate = ts.size * n_f / (ms * 1e-3)
    base = base or rate
    print(f"n_f={n_f:7d}: K1 {ms:9.2f} ms  {rate:.3e} ev*tr/s ({rate/base:.3f} of the 1e5-f rate)  grid={grid} block={block} smem={smem} shape={_cabi.fold_last_kernel_shape()}", flush=True)
