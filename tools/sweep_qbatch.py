"""K2/K3 on the full C4 shape (1e8 events, 1252 centres) for several profile-scratch batch sizes (HX_QFFA_BATCH_MIB)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.efsearch import _qffa_device_call, _shift_grids, qffa_centres
n_ev = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
g = torch.Generator(device="cuda"); g.manual_seed(4)
td = torch.sort(torch.rand(n_ev, device="cuda", dtype=torch.float64, generator=g) * 1e5 - 5e4).values
t0, t1 = float(td[0]), float(td[-1])
cents, step = qffa_centres(9.85, 9.95, 1e5, 2)
nbin, n = 128, 2
L, Q = _shift_grids(nbin, 2, 8, True)
ref = None
for mib in (512, 256, 1024, 2048, 4096):
    os.environ["HX_QFFA_BATCH_MIB"] = str(mib)
    st = _qffa_device_call(td, t0, t1, cents, nbin, 16 * nbin, L, Q, n)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a = time.perf_counter()
        st = _qffa_device_call(td, t0, t1, cents, nbin, 16 * nbin, L, Q, n)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - a)
    ms, nl = _cabi.qffa_last_kernel()
    chk = float(np.sum(st))
    ref = ref or chk
    print(f"batch {mib:5d} MiB: search {best*1e3:7.2f} ms, K2 {ms:6.2f} ms over {nl} launches -> {n_ev*len(cents)/best:.3e} ev*centres/s  {'ok' if chk == ref else 'DIFF %r' % (chk-ref)}", flush=True)
