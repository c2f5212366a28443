"""Time hx_ffa_dev on a C4-sized light curve (T = 1e5 s, f ~ 10 Hz, nbin = 128 -> 1.28e8 bins, 2^20 tables) for
several steps-per-pass settings; prints ms, passes and the algorithmic / streamed bandwidth."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.ffa import _ffa_device
lib = _cabi.lib()
n_raw = int(float(sys.argv[1])) if len(sys.argv) > 1 else 127_000_000
g = torch.Generator(device="cuda"); g.manual_seed(0)
counts = (torch.rand(n_raw, device="cuda", generator=g) < 0.1).to(torch.int32)
ref = None
for P in (128, 129, 250):
    for k in (0, 2, 3, 4, 5, 6, 7, 8, -1):
        lib.hx_ffa_set_tuning(k)
        _ffa_device(counts, P)
        best = 1e9
        for _ in range(3):
            per, st = _ffa_device(counts, P); best = min(best, _cabi.ffa_last_kernel()[0])
        ms, passes, steps, kk = _cabi.ffa_last_kernel()
        nel = per.size * P
        crc = float(np.nansum(st))
        if k == 0: ref = crc
        print(f"P={P} k={k:2d} (used {kk}): {best:8.3f} ms  passes={passes} steps={steps}  algorithmic {8*nel*steps/best/1e6:9.1f} GB/s  "
              f"streamed ~{8*nel*passes/best/1e6:8.1f} GB/s  {'ok' if abs(crc-ref) <= 1e-9*abs(ref) else 'MISMATCH'}", flush=True)
lib.hx_ffa_set_tuning(0)
