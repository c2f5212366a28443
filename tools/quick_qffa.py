"""Developer timing of the --fast path (K2 + K3) on the C4 shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.efsearch import _qffa_device_call, _shift_grids, qffa_centres

n_ev = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
ncent = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(4)
ts = np.sort(rng.uniform(-5e4, 5e4, n_ev))
td = torch.from_numpy(ts).cuda()
cents, step = qffa_centres(9.85, 9.95, 1e5, 2)
cents = cents[:ncent]
lib = _cabi.lib()
for nbin, n in ((128, 2), (16, 1)):
    nprof = 16 * nbin
    for sfd in (True, False):
        L, Q = _shift_grids(nbin, 2, 8 if n == 2 else 8, sfd)
        for cmax in (1008, 8):
            lib.hx_qffa_set_tuning(cmax)
            _qffa_device_call(td, ts[0], ts[-1], cents, nbin, nprof, L, Q, n)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _qffa_device_call(td, ts[0], ts[-1], cents, nbin, nprof, L, Q, n)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            per = dt / len(cents)
            print(f"N={n_ev:.0e} nbin={nbin} n={n} nLQ={L.size} cmax={cmax}: {per*1e6:8.1f} us/centre "
                  f"=> event-bytes {8*n_ev/per/1e9:8.1f} GB/s-equivalent per centre, "
                  f"K2 kernels {_cabi.qffa_last_kernel()[0]/len(cents)*1e3:7.1f} us/centre ({'lane=event' if cmax >= 1000 else 'lane=centre'})", flush=True)
lib.hx_qffa_set_tuning(0)
