"""Config 5 accuracy report: reference-parity fp64 phase vs the exact-phase extension on a long
baseline, judged against exact rational arithmetic (oracle.exact_bins) and the reference's own
criterion check_phase_error_after_casting_to_double (hendrics/efsearch.py:199-205)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hendrics_b200 import fold_profiles, z_n_search
from hendrics_b200.efsearch import check_phase_error_after_casting_to_double
from hendrics_b200.fake import simulate_events
from oracle import hx_oracle as O

nbin, n_sample = 128, 20000
np.random.seed(0)
print("C5 accuracy: T = 1e8 s, nbin = 128; bin flips w.r.t. exact rational arithmetic on a %d-event sample" % n_sample)
for T, f0 in ((1e8, 9.9009900990099), (1e8, 700.123456789), (1e9, 700.123456789)):
    ts = simulate_events(n_sample, T, f0=f0, seed=5)
    ts = ts - ts[0]
    for fd in (0.0, -2e-15):
        exact = O.exact_bins(ts, f0, fd, nbin)
        ph = np.empty_like(ts); O.pulse_phase_fast(ts, f0, fd, ph)
        b64 = np.minimum(np.floor(ph * nbin).astype(np.int64), nbin - 1)
        s_par, p_par = fold_profiles(ts, [f0], [fd], nbin=nbin, nharm=2)
        s_ex, p_ex = fold_profiles(ts, [f0], [fd], nbin=nbin, nharm=2, exact_phase=True)
        pe = np.bincount(exact, minlength=nbin)
        err = check_phase_error_after_casting_to_double(T, f0, fd)
        print(f"T={T:.0e} f={f0:.6f} fd={fd:+.1e}: |t f|max={T*f0:.2e} cycles; fp64 flips {int(np.sum(b64 != exact)):5d}/{n_sample} "
              f"(GPU parity profile == reference fp64: {np.array_equal(p_par[0,0], np.bincount(b64, minlength=nbin))}); "
              f"exact-phase profile == rational: {np.array_equal(p_ex[0,0], pe)}; "
              f"Z2 parity {s_par[0,0]:.6f} vs exact {s_ex[0,0]:.6f} (rel {abs(s_par[0,0]-s_ex[0,0])/s_ex[0,0]:.1e}); "
              f"reference long-double criterion {float(err):.2e} cycles (warn if > {1/nbin/10:.1e})")
