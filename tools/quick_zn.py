"""Developer timing of K4 (unbinned Z^2_n)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hendrics_b200 import _cabi
lib = _cabi.lib()
rng = np.random.default_rng(0)
for n_ev, n_f, n_fd, nh in ((1_000_000, 100_000, 1, 2), (1_000_000, 2048, 64, 2), (1_000_000, 100_000, 1, 4), (100_000, 10_000, 1, 2)):
    ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, n_ev))).cuda()
    f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
    fd = torch.from_numpy(-1.6e-9 + 5e-11 * np.arange(n_fd) if n_fd > 1 else np.zeros(1)).cuda()
    out = torch.empty(n_f * n_fd, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    best = 1e30
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.check(lib.hx_zn_unbinned_dev(ts.data_ptr(), n_ev, f.data_ptr(), n_f, fd.data_ptr(), n_fd, nh, out.data_ptr(), st))
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    evtr = n_ev * n_f * n_fd
    print(f"unbinned N={n_ev:.0e} nf={n_f} nfd={n_fd} nharm={nh}: {best*1e3:8.2f} ms {evtr/best:.3e} ev*tr/s ({evtr/best/148/1.965e9:.2f}/SM/clk)", flush=True)
