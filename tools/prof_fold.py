"""Run K1 once (after one warm-up) on a given shape: target for ncu."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device

n_ev, n_f, n_fd, nbin, nharm = [int(float(x)) for x in sys.argv[1:6]]
R = int(sys.argv[6]) if len(sys.argv) > 6 else 0
W = int(sys.argv[7]) if len(sys.argv) > 7 else 0
lib = _cabi.lib()
lib.hx_fold_set_tuning.argtypes = [ctypes.c_int, ctypes.c_int]
lib.hx_fold_set_tuning(R, W)
rng = np.random.default_rng(0)
ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, n_ev))).cuda()
f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd) if n_fd > 1 else np.zeros(1)).cuda()
for _ in range(2):
    fold_stats_device(ts, f, fd, nbin, nharm)
torch.cuda.synchronize()
print("ok")
