"""Run the --fast kernels once (after a warm-up) on the C4 shape: target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.efsearch import _qffa_device_call, _shift_grids, qffa_centres
n_ev = int(float(sys.argv[1])); cmax = int(sys.argv[2]); nbin = int(sys.argv[3]) if len(sys.argv) > 3 else 128
rng = np.random.default_rng(4)
ts = np.sort(rng.uniform(-5e4, 5e4, n_ev))
td = torch.from_numpy(ts).cuda()
cents, step = qffa_centres(9.85, 9.95, 1e5, 2)
L, Q = _shift_grids(nbin, 2, 8, True)
_cabi.lib().hx_qffa_set_tuning(cmax)
for _ in range(2):
    _qffa_device_call(td, ts[0], ts[-1], cents[:cmax], nbin, 16 * nbin, L, Q, 2)
torch.cuda.synchronize()
x = td.sum(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); x = td.sum(); e1.record(); torch.cuda.synchronize()
print("torch sum over events: %.1f us -> %.0f GB/s" % (e0.elapsed_time(e1) * 1e3, 8 * n_ev / e0.elapsed_time(e1) / 1e6))
