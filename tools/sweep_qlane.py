"""Sweep K2-lane's (centres/thread, warps/CTA) on the C4 shape (HX_QLANE_RW hook)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.efsearch import _qffa_device_call, _shift_grids, qffa_centres

n_ev = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(4)
ts = np.sort(rng.uniform(-5e4, 5e4, n_ev))
td = torch.from_numpy(ts).cuda()
cents, step = qffa_centres(9.85, 9.95, 1e5, 2)
for nbin, n in ((128, 2), (64, 2), (16, 1)):
    nprof = 16 * nbin
    L, Q = _shift_grids(nbin, 2, 8, True)
    for rw in ("", "1,2", "1,4", "1,8", "2,1", "2,2", "2,3", "2,4", "2,6", "4,1", "4,2", "4,3"):
        if rw: os.environ["HX_QLANE_RW"] = rw
        else: os.environ.pop("HX_QLANE_RW", None)
        try:
            _qffa_device_call(td, ts[0], ts[-1], cents, nbin, nprof, L, Q, n)
            torch.cuda.synchronize()
            _qffa_device_call(td, ts[0], ts[-1], cents, nbin, nprof, L, Q, n)
            ms, nl = _cabi.qffa_last_kernel()
            print(f"nbin={nbin} RW={rw or 'auto':5s}: K2 {ms:8.2f} ms over {nl} launches -> {n_ev*len(cents)/(ms*1e-3):.3e} ev*centres/s", flush=True)
        except Exception as ex:
            print(f"nbin={nbin} RW={rw}: {ex}", flush=True)
