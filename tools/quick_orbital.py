"""Throughput of the orbital-demodulation fold (hx_orbital_search_dev): one parameter row, all its trial epochs T0 in
one launch.  Unit = event * T0 trial (two fp64 sin(), one IEEE division each, + one generic fold)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.fake import simulate_events

n_ev = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
n_t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 4736  # 148 SMs x 32 lanes
ts = simulate_events(n_ev, 1e5, seed=6)
td = torch.from_numpy(ts).cuda()
t0 = torch.from_numpy(np.random.default_rng(0).uniform(0, 86400.0, n_t0)).cuda()
stats = torch.empty(n_t0, dtype=torch.float64, device="cuda")
lib = _cabi.lib()
def run():
    _cabi.check(lib.hx_orbital_search_dev(td.data_ptr(), td.numel(), 9.9, 1e-10, 3.0, 86400.0, t0.data_ptr(), n_t0, 64, 2,
                                          stats.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
for _ in range(2):
    run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
peak = _cabi.microbench(0, reps=3, device=0)
print(f"orbital fold: {n_ev} events x {n_t0} T0 trials, Z^2_2 nbin 64: {ms:.2f} ms per row -> {n_ev * n_t0 / ms / 1e-3:.3e} event*T0/s "
      f"(FP64 issue rate {peak:.3e} inst/s => {peak / (n_ev * n_t0 / ms / 1e-3):.0f} FP64 issue slots available per unit)")
