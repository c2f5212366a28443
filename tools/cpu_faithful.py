"""The reference's own code structure on this box's CPU: serial Python-level trial loop over numba-JIT'd
per-trial kernels (oracle.folding_search_core, parallel=False), 1 thread; and the same with prange."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import hx_oracle as O, c_oracle as C
from hendrics_b200.fake import simulate_events
for name, n_ev, nbin, nharm, n_f in (("C1-like", 100_000, 128, 2, 2000), ("C2-like", 1_000_000, 16, 0, 200)):
    ts = simulate_events(n_ev, 1e5, seed=1)
    f = 9.876 + 5e-6 * np.arange(n_f)
    O.folding_search_core(ts, f[:4], nbin=nbin, nharm=nharm)              # JIT warm-up
    O.folding_search_core(ts, f[:4], nbin=nbin, nharm=nharm, parallel=True)
    t0 = time.perf_counter(); O.folding_search_core(ts, f, nbin=nbin, nharm=nharm); t1 = time.perf_counter()
    O.folding_search_core(ts, f, nbin=nbin, nharm=nharm, parallel=True); t2 = time.perf_counter()
    tt = (ts - ts[0]); tt = tt[tt < tt[-1]]
    C.fold_search(tt, f, [0.0], nbin, nharm); t3 = time.perf_counter()
    u = (n_ev - 1) * n_f
    print(f"{name}: numba serial (reference structure, 1 thread) {u/(t1-t0):.3e} ev*tr/s; numba prange ({os.cpu_count()} cpus) {u/(t2-t1):.3e}; "
          f"C port ({C.max_threads()} threads) {u/(t3-t2):.3e}")
