import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import c_oracle as C, hx_oracle as O
from hendrics_b200 import _cabi
sys.path.insert(0, "tests")
from test_qffa_gpu import _host_qffa, _centred
nbin, nprof, n = 128, 2048, 2
ts = _centred(40000, 2000.0, seed=nbin)
cents = [(9.88 + 0.004 * i, 0.0, 0.0) for i in range(11)]
lin = np.linspace(-nbin * 2, nbin * 2, 16)
L, Q = np.meshgrid(lin, lin, indexing="ij")
stats, prof = _host_qffa(ts, cents, nbin, nprof, L, Q, n)
for ic, (f, fd, fdd) in enumerate(cents[:3]):
    ref = C.qffa_profiles(ts, f, fd, fdd, nbin, nprof).astype(np.int64)
    got = prof[ic].astype(np.int64)
    d = got - ref
    rows = np.nonzero(np.any(d != 0, axis=1))[0]
    print("centre", ic, "sum got", got.sum(), "ref", ref.sum(), "n diff cells", np.count_nonzero(d), "rows", rows[:20], "...", rows[-5:] if rows.size else "")
    if rows.size:
        r = rows[0]
        print(" row", r, "got", np.nonzero(got[r])[0], got[r][got[r] != 0], "ref", np.nonzero(ref[r])[0], ref[r][ref[r] != 0])
