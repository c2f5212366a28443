"""Generate golden vectors by EXECUTING THE REAL REFERENCE (read-only tree at /root/reference).

Runs only in the build container (the GPU box has no /root/reference); the .npz outputs are
committed and used by tests/test_golden.py (CPU: oracle vs golden) and tests/test_golden_gpu.py
(GPU: CUDA path vs golden).

HENDRICS imports stingray/astropy at module import time and neither is installable here, so
those third-party packages are replaced by inert stub modules; every function called below is
IN-TREE reference code (numpy/numba only) and runs unmodified:

    hendrics.efsearch: _fast_phase, _fast_phase_fdot, _fast_phase_fddot, shift_and_sum, _fast_step,
                       search_with_qffa_step, search_with_qffa, z_n_fast, folding_search,
                       _average_and_z_sub_search, transient_search,
                       check_phase_error_after_casting_to_double
    hendrics.ffa:      _z_n_fast_cached, z_n_fast_cached, z_n_fast_cached_all, h_test, ffa, ffa_search
    hendrics.efsearch: search_with_ffa (-> hendrics.base.histogram / hist1d_numba_seq)
    hendrics.base:     hist2d_numba_seq / histogram2d, find_peaks_in_image(rough=True)

The stingray functions (z_n_search, epoch_folding_search) are NOT available: that boundary stays
"parity unpinned" (see oracle/hx_oracle.py).
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


class StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("stingray", "astropy", "matplotlib", "pint", "netCDF4", "h5py", "skimage", "imageio", "tqdm")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        m.__version__ = "2.3.2"
        return m

    def exec_module(self, module):
        pass


def main():
    sys.meta_path.insert(0, StubFinder())
    sys.path.insert(0, "/root/reference")
    from hendrics import base as hbase
    from hendrics import efsearch as hef
    from hendrics import ffa as hffa

    from hendrics_b200.fake import simulate_events

    # tqdm is one of the stubbed third-party modules: make the progress-bar wrapper the identity
    hef.show_progress = lambda x: x

    out = {}

    # ---- 1. single --fast steps: phases, 2-D profiles, shift-and-sum statistics -----------------
    cases = [
        # name, n_ev, T, seed, mean_f, fd, fdd, nbin, nprof, n, search_fdot
        ("step_f", 20000, 1000.0, 11, 9.9012, 0.0, 0.0, 16, 128, 1, True),
        ("step_fd", 20000, 1000.0, 12, 9.8931, 3e-7, 0.0, 32, 64, 2, True),
        ("step_fdd", 15000, 2000.0, 13, 9.9077, -2e-7, 4e-11, 64, 256, 3, True),
        ("step_n128", 30000, 1000.0, 14, 9.9, 0.0, 0.0, 128, 2048, 2, True),
        ("step_nofd", 8000, 300.0, 15, 9.91, 0.0, 0.0, 16, 128, 2, False),
        ("step_nbin10", 8000, 300.0, 16, 3.377, 1e-6, 0.0, 10, 37, 1, True),
    ]
    for name, n_ev, T, seed, mf, fd, fdd, nbin, nprof, n, sfd in cases:
        ts = simulate_events(n_ev, T, seed=seed)
        tc = (ts - (ts.max() + ts.min()) / 2).astype(np.float64)
        if fdd != 0:
            ph = hef._fast_phase_fddot(tc, mf, fd, fdd)
        elif fd != 0:
            ph = hef._fast_phase_fdot(tc, mf, fd)
        else:
            ph = hef._fast_phase(tc, mf)
        prof = hbase.histogram2d(ph, tc, range=[[0, 1], [tc[0], tc[-1]]], bins=(nbin, nprof)).T
        fgrid, fdgrid, stats = hef.search_with_qffa_step(
            tc, np.double(mf), mean_fdot=np.double(fd), mean_fddot=np.double(fdd), nbin=nbin, nprof=nprof, npfact=2,
            oversample=8, n=n, search_fdot=sfd)
        out[name] = dict(times_c=tc, params=np.array([mf, fd, fdd, nbin, nprof, n, int(sfd)], dtype=np.float64),
                         profiles=prof.astype(np.uint32), fgrid=fgrid, fdotgrid=fdgrid, stats=stats,
                         phases_head=ph[:64])

    # ---- 2. a complete search_with_qffa --------------------------------------------------------
    ts = simulate_events(12000, 200.0, seed=21) + 8e7
    np.random.seed(1234)
    res = hef.search_with_qffa(ts, 9.85, 9.95, nbin=16, n=2, oversample=8, silent=True)
    out["qffa_full"] = dict(times=ts, fgrid=np.asarray(res[0]), fdotgrid=np.asarray(res[1]), stats=np.asarray(res[2]),
                            scalars=np.array([res[3], res[4], res[5]], dtype=np.float64))
    np.random.seed(1234)
    res = hef.search_with_qffa(ts, 9.85, 9.95, nbin=16, n=1, oversample=4, search_fdot=False, silent=True)
    out["qffa_full_nofd"] = dict(times=ts, fgrid=np.asarray(res[0]), stats=np.asarray(res[1]),
                                 scalars=np.array([res[2], res[3]], dtype=np.float64))

    # ---- 3. statistic formulas -----------------------------------------------------------------
    rng = np.random.default_rng(31)
    prof = rng.poisson(200, size=(6, 64)).astype(np.float64)
    out["zn_cached"] = dict(
        profiles=prof,
        z2=np.array([hffa.z_n_fast_cached(p, n=2) for p in prof]),
        z4=np.array([hffa.z_n_fast_cached(p, n=4) for p in prof]),
        zall=np.array([hffa.z_n_fast_cached_all(p, nmax=6)[1] for p in prof]),
        htest=np.array([hffa.h_test(p, nmax=6) for p in prof], dtype=np.float64),
    )
    ph = 2 * np.pi * rng.uniform(0, 1, 5000)
    out["zn_fast"] = dict(phase=ph, z=np.array([hef.z_n_fast(ph, np.ones_like(ph), n=k) for k in (1, 2, 3, 4)]))

    # ---- 4. shift_and_sum ----------------------------------------------------------------------
    profs = rng.poisson(30, size=(128, 32)).astype(np.float64)
    rep = np.hstack((profs, profs, profs))
    base_shift = np.linspace(-1, 1, 128)
    splats = []
    lq = [(-64.0, 17.5), (3.3, -41.0), (0.0, 0.0), (63.99, 64.0), (-12.5, 12.5)]
    for l, q in lq:
        splats.append(np.array(hef.shift_and_sum(rep, l, q, np.zeros(32), base_shift, base_shift**2)))
    out["shift_and_sum"] = dict(profiles=profs, lq=np.array(lq), splats=np.array(splats))

    # ---- 5. histogram2d / find_peaks_in_image ----------------------------------------------------
    x = rng.uniform(0, 1, 4000)
    y = np.sort(rng.uniform(2, 3, 4000))
    H = hbase.histogram2d(x, y, range=[[0, 1], [y[0], y[-1]]], bins=(12, 9))
    out["hist2d"] = dict(x=x, y=y, H=np.asarray(H, dtype=np.uint64))
    img = rng.normal(0, 1, (40, 30))
    out["peaks"] = dict(image=img, idx=np.asarray(hbase.find_peaks_in_image(img, n=5, rough=True)))

    # ---- 6. folding_search trial grids (in-tree grid construction, stingray func replaced) -------
    class Ev:
        pass

    grids = []
    for (T, fmin, fmax, os_, fdmin, fdmax) in [(25.25, 9.85, 9.95, 2, 0, 0), (1000.0, 9.89, 9.91, 2, -1e-6, 1e-6),
                                                 (100000.0, 9.876, 9.926, 2, 0, 0), (300.0, 5.0, 5.0, 4, 0, 2e-5)]:
        ev = Ev()
        ev.time = np.array([8e7, 8e7 + T])
        ev.gti = np.array([[8e7, 8e7 + T]])
        cap = {}

        def func(times, freqs, fdots=0, **kw):
            cap["f"] = np.array(freqs)
            cap["fd"] = np.array(fdots)
            return freqs, np.zeros_like(freqs)

        r = hef.folding_search(ev, fmin, fmax, func=func, oversample=os_, fdotmin=fdmin, fdotmax=fdmax)
        grids.append(dict(T=T, fmin=fmin, fmax=fmax, oversample=os_, fdmin=fdmin, fdmax=fdmax, freqs=cap["f"],
                          fdots=cap["fd"], step=r[2], length=r[3]))
    for i, g in enumerate(grids):
        out[f"grid{i}"] = {k: np.asarray(v) for k, v in g.items()}

    # ---- 7. transient search (in-tree, efsearch.py:407-650) ---------------------------------------
    prof = rng.poisson(20, size=(16, 32)).astype(np.float64)
    prof[5] = 0
    nave, res = hef._average_and_z_sub_search(prof, n=2)
    out["transient_sub"] = dict(profiles=prof, nave=np.asarray(nave), results=np.asarray(res))
    ts = simulate_events(15000, 128.0, seed=41) + 8e7
    np.random.seed(7)
    tr = hef.transient_search(ts, 9.85, 9.95, nbin=16, n=2, oversample=4)
    out["transient_full"] = dict(times=ts, nave=np.asarray(tr.nave), freqs=np.asarray(tr.freqs), tgrid=np.asarray(tr.times),
                                 stats=np.asarray(tr.stats))
    tr = hef.transient_search(ts, 9.88, 9.92, fdot=1e-6, nbin=32, nprof=96, n=1, oversample=2)
    out["transient_np96"] = dict(times=ts, nave=np.asarray(tr.nave), freqs=np.asarray(tr.freqs), tgrid=np.asarray(tr.times),
                                 stats=np.asarray(tr.stats))
    # nprof = 100: `int(2 ** np.log2(100))` is 99 (floating-point round trip), so the reference histograms 100 time
    # slices and averages the first 99 -- results have 99 columns and int(99 / 2^l) groups per level
    tr = hef.transient_search(ts[:6000], 9.88, 9.92, nbin=16, nprof=100, n=2, oversample=2)
    out["transient_np100"] = dict(times=ts[:6000], nave=np.asarray(tr.nave), freqs=np.asarray(tr.freqs),
                                  tgrid=np.asarray(tr.times), stats=np.asarray(tr.stats))

    # ---- 8. fast folding algorithm, --ffa (in-tree: ffa.py:210-395, efsearch.py:1092-1138) -------
    hffa.show_progress = lambda x: x
    rng = np.random.default_rng(8)
    counts = rng.poisson(2.0, 6000).astype(np.float64)
    counts[1000:1300] = 0  # a data gap
    # 49: np.arange(0, 1, 1/49) has 50 elements, so the reference's tiled harmonic tables are skewed for k = 2
    for bp in (16, 37, 49, 128):
        per, st = hffa.ffa(counts, bp)
        out[f"ffa_single_{bp}"] = dict(counts=counts, bin_period=np.array([bp]), periods=per, stats=st)
    per, st = hffa.ffa_search(counts, 0.01, 0.33, 0.7)  # period range > 2x: the rebinning branch runs
    out["ffa_search"] = dict(counts=counts, params=np.array([0.01, 0.33, 0.7]), periods=per, stats=st)
    ts = simulate_events(20000, 200.0, seed=21) + 8e7
    res = hef.search_with_ffa(ts, 9.85, 9.95, nbin=16, n=2)
    out["ffa_events"] = dict(times=ts, params=np.array([9.85, 9.95, 16, 2]), freqs=np.asarray(res[0]), stats=np.asarray(res[1]),
                             length=np.array([res[3]]))
    res = hef.search_with_ffa(ts[:6000], 3.1, 3.4, nbin=24, n=1)
    out["ffa_events_nbin24"] = dict(times=ts[:6000], params=np.array([3.1, 3.4, 24, 1]), freqs=np.asarray(res[0]),
                                    stats=np.asarray(res[1]), length=np.array([res[3]]))

    # ---- 9. candidate post-processing (in-tree, numpy only: efsearch.py:1299-1332, 208-260) -------------------
    rng = np.random.default_rng(9)
    bx, by, blevel, bx0, bout = [], [], [], [], []
    for k in range(64):
        n = int(rng.integers(1, 60))
        x = np.sort(rng.uniform(-5, 5, n))
        y = rng.uniform(0, 1, n) if k % 4 else np.sinc(x) ** 2 + np.sinc(x - 2.5) ** 2
        level, x0 = float(rng.uniform(0, 1)), float(rng.uniform(-6, 6))
        lo, hi = hef.get_boundaries_from_level(x, y, level, x0)
        bx.append(np.pad(x, (0, 60 - n))), by.append(np.pad(y, (0, 60 - n)))
        blevel.append([n, level, x0]), bout.append([lo, hi])
    out["boundaries"] = dict(x=np.array(bx), y=np.array(by), params=np.array(blevel), out=np.array(bout))
    import tempfile

    np.random.seed(9)
    csv = os.path.join(tempfile.mkdtemp(), "db.csv")
    hef.decide_binary_parameters(137430, [0.03, 0.035], [2.0 * 86400, 2.5 * 86400], [0.0, 1.0], fdot_range=[0, 5e-10],
                                 NMAX=3, csv_file=csv, reset=True)
    import pandas as pd

    df = pd.read_csv(csv)
    out["binary_parameters"] = dict(table=df[["freq", "fdot", "X", "Porb"]].to_numpy(dtype=np.float64))

    only = sys.argv[1] if len(sys.argv) > 1 else ""  # `make_golden.py ffa`: (re)write only the fixtures with that prefix
    for name, d in out.items():
        if name.startswith(only):
            np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **d)
    total = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print(f"wrote {len(out)} fixtures, {total / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
