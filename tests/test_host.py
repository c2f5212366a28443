"""CPU: host-side logic of the product (no GPU, no compute calls) and the C-ABI surface."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from hendrics_b200 import _cabi

    hdr = open(os.path.join(ROOT, "include", "hx_search.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char \*)\s*\*?\s*(hx_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hx_search.h but not exported"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    assert _cabi.lib().hx_version() >= 1000


def test_no_gpu_means_loud_failure_not_fallback():
    from hendrics_b200 import _cabi

    if _cabi.device_count() > 0:
        pytest.skip("a GPU is present")
    import hendrics_b200

    with pytest.raises(_cabi.HxError):
        hendrics_b200.z_n_search(np.arange(10.0), np.array([1.0]))
    # host-pointer C entry points fail with HX_ERR_CUDA as well
    t = np.arange(10.0)
    f = np.array([1.0])
    s = np.zeros(1)
    rc = _cabi.lib().hx_fold_search(_cabi.dptr(t), 10, _cabi.dptr(f), 1, _cabi.dptr(f), 1, 16, 0, 0, _cabi.dptr(s),
                                    None, 0)
    assert rc == _cabi.HX_ERR_CUDA and b"" != _cabi.lib().hx_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hendrics_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
                assert "libhxoracle" not in src, fn


def test_argument_validation_without_gpu():
    from hendrics_b200 import _cabi

    lib = _cabi.lib()
    t = np.arange(10.0)
    s = np.zeros(1)
    assert lib.hx_fold_search(None, 10, _cabi.dptr(t), 1, _cabi.dptr(t), 1, 16, 0, 0, _cabi.dptr(s), None, 0) == _cabi.HX_ERR_ARG
    assert lib.hx_fold_search(_cabi.dptr(t), -1, _cabi.dptr(t), 1, _cabi.dptr(t), 1, 16, 0, 0, _cabi.dptr(s), None, 0) == _cabi.HX_ERR_ARG
    out = ctypes.c_double(0)
    assert lib.hx_microbench(99, 1, ctypes.byref(out), 0) == _cabi.HX_ERR_ARG
    assert b"bad arguments" in lib.hx_last_error()


def test_hen_root_and_negative_args():
    from hendrics_b200.base import check_negative_numbers_in_args, hen_root

    assert hen_root("blabla_ev_calib.nc") == "blabla"
    assert hen_root("blablu_ev_bli.fits.gz") == "blablu_ev_bli"
    assert hen_root("blablu_lc_asrd_ev_lc.nc") == "blablu_lc_asrd"
    args = ["events.nc", "-f", "103", "--fdot", "-2e-10"]
    new = check_negative_numbers_in_args(args)
    assert args[:4] == new[:4] and new[4] == " -2e-10"


def test_parser_flag_surface():
    from hendrics_b200.efsearch import _common_parser

    a = _common_parser(["x.p", "-f", "1", "-F", "2"])
    assert (a.nbin, a.N, a.segment_size, a.dynstep, a.npfact, a.conflevel, a.curve) == (128, 2, 1e32, 128, 2, 99, "sinc")
    assert a.oversample is None and a.step is None and not a.fast and a.fdotmin is None and a.mean_fdot == 0
    a = _common_parser(["x.p", "-f", "1", "-F", "2", "--fdotmin", "-1e-9", "--fdotmax", "1e-9", "--fast", "-N", "3",
                        "-n", "64", "--oversample", "8", "--mean-fddot", "-1e-20", "--force-memmap"])
    assert a.fdotmin == -1e-9 and a.fast and a.N == 3 and a.nbin == 64 and a.oversample == 8 and a.force_memmap


def test_efperiodogram_roundtrip_and_peaks(tmp_path):
    from hendrics_b200.io import EFPeriodogram, load_folding, save_folding

    freq = np.linspace(9.85, 9.95, 201)
    stat = np.random.default_rng(0).chisquare(4, 201)
    stat[100] = 400.0
    stat[99] = 300.0
    stat[30] = 200.0
    ef = EFPeriodogram(freq, stat, "Z2n", 64, 2, M=1, oversample=2, ncounts=1000, filename="x", mjdref=56000, pepoch=56000.5)
    peaks, pstat = ef.find_peaks(conflevel=99)
    assert np.isclose(peaks[0], freq[100]) and np.isclose(peaks[1], freq[30]) and len(peaks) == 2
    ef.best_fits = [{"model": "sinc", "mean": 9.9}]
    fn = str(tmp_path / "a_Z22.p")
    save_folding(ef, fn)
    ef2 = load_folding(fn)
    assert ef2.kind == "Z2n" and ef2.N == 2 and ef2.nbin == 64 and ef2.ncounts == 1000
    assert np.array_equal(ef2.freq, freq) and np.array_equal(ef2.stat, stat)
    assert ef2.best_fits == [{"model": "sinc", "mean": 9.9}]
    import pickle

    raw = pickle.load(open(fn, "rb"))
    assert raw["__sr__class__type__"] == "EFPeriodogram" and "best_fits" not in raw
    for key in ("freq", "stat", "kind", "nbin", "N", "oversample", "M", "pepoch", "mjdref", "peaks", "peak_stat", "fdots",
                "fddots", "segment_size", "filename", "parfile", "emin", "emax", "upperlim", "ncounts"):
        assert key in raw, key


def test_detection_levels_sane():
    from scipy import stats

    from hendrics_b200.base import fold_detection_level, pf_upper_limit, z2_n_detection_level

    assert np.isclose(z2_n_detection_level(n=2, epsilon=0.01, ntrial=1), stats.chi2.isf(0.01, 4))
    assert z2_n_detection_level(n=2, epsilon=0.01, ntrial=1000) > z2_n_detection_level(n=2, epsilon=0.01, ntrial=1)
    assert np.isclose(fold_detection_level(16, 0.01), stats.chi2.isf(0.01, 15))
    assert 0 < pf_upper_limit(40.0, 10000, n=2) < 1


def test_qffa_centres_match_oracle():
    from hendrics_b200.efsearch import qffa_centres
    from oracle import hx_oracle as O

    for args in [(9.85, 9.95, 200.0, 2), (9.0, 9.0, 1000.0, 2), (5.0, 5.5, 33.3, 3)]:
        a, sa = qffa_centres(*args)
        b, sb = O.qffa_centres(*args)
        assert sa == sb and np.array_equal(np.array(a), np.array(b))


def test_shard_bounds():
    from hendrics_b200.dist import shard_bounds

    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
dist.init_process_group("gloo")
from hendrics_b200 import dist as hd
from oracle import hx_oracle as O
ts = O.simulate_events(3000, 200.0, seed=1)
freqs = 9.85 + 0.1 * np.arange(37) / 37
full = O.z_n_search(ts, freqs, nharm=2, nbin=32)
got = hd.sharded_search(O.z_n_search, ts, freqs, fdots=0, nharm=2, nbin=32)
assert len(got) == 2 and np.array_equal(got[0], full[0]) and np.allclose(got[1], full[1], rtol=1e-13)
fd = np.array([0.0, 1e-6, 2e-6])
full3 = O.epoch_folding_search(ts, freqs, nbin=16, fdots=fd)
got3 = hd.sharded_search(O.epoch_folding_search, ts, freqs, fdots=fd, nbin=16)
assert got3[2].shape == (3, 37) and np.allclose(got3[2], full3[2], rtol=1e-13)
assert np.array_equal(got3[0], full3[0]) and np.array_equal(got3[1], full3[1])
rows = np.arange(11 * 3, dtype=np.float64).reshape(11, 3)
out = hd.sharded_rows(lambda r: r[:, :1] * np.ones((1, 4)), rows)
assert out.shape == (11, 4) and np.array_equal(out[:, 0], rows[:, 0])
# a failure on one rank is raised on every rank instead of dead-locking the all-gather
def bad(times, f, fdots=0, **kw):
    if dist.get_rank() == 1:
        raise ValueError("boom")
    return O.z_n_search(times, f, fdots=fdots, **kw)
try:
    hd.sharded_search(bad, ts, freqs, fdots=0, nharm=2, nbin=32)
    raise SystemExit("no error raised")
except (ValueError, RuntimeError) as e:
    assert ("boom" in str(e)) == (dist.get_rank() == 1), str(e)
# torch tensors (the device-resident path of the CUDA searches) gather like numpy arrays
import torch
r = dist.get_rank()
loc = torch.arange(3 * (4 + r), dtype=torch.float64).reshape(3, 4 + r) + 100 * r
out = hd.all_gather_concat(loc, [4, 5], axis=-1)
assert out.shape == (3, 9) and out[0, 0] == 0 and out[0, 4] == 100 and out[2, 8] == 100 + 14
print("rank", dist.get_rank(), "ok")
dist.destroy_process_group()
"""


def test_sharded_search_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29581", str(script), ROOT]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("ok") == 2


def test_lightcurve_container_and_file_type(tmp_path):
    """light-curve input of the search path (hendrics/efsearch.py:1154-1157): time / counts / gti round trip."""
    from hendrics_b200.io import EventList, Lightcurve, get_file_type, save_events, save_lcurve

    t = np.arange(0.5, 100.5)
    lc = Lightcurve(time=t, counts=np.arange(100), mjdref=55000)
    assert lc.dt == 1.0 and np.allclose(lc.gti, [[0.0, 100.0]])
    fn = str(tmp_path / "x_lc.p")
    save_lcurve(lc, fn)
    kind, back = get_file_type(fn)
    assert kind == "lc" and np.array_equal(back.counts, lc.counts) and np.array_equal(back.time, t)
    assert hasattr(back, "counts") and back.mjdref == 55000
    fn2 = str(tmp_path / "x_ev.p")
    save_events(EventList(time=t), fn2)
    kind, ev = get_file_type(fn2)
    assert kind == "events" and not hasattr(ev, "counts")
    with pytest.raises(ValueError):
        Lightcurve(time=t, counts=np.arange(3))


def test_slow_branch_selection_mirrors_upstream():
    """expocorr or array weights select the fold_events branch; scalar weights do not; expocorr needs the GTIs."""
    from hendrics_b200 import epoch_folding_search, z_n_search
    from hendrics_b200.search import _is_slow_branch

    assert _is_slow_branch(True, 1) and _is_slow_branch(False, np.ones(3)) and _is_slow_branch(False, [1, 2])
    assert not _is_slow_branch(False, 1) and not _is_slow_branch(False, 2.5)
    for fn in (epoch_folding_search, z_n_search):
        with pytest.raises(ValueError, match="specify the GTIs"):
            fn(np.arange(10.0), [1.0], expocorr=True)
