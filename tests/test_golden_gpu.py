"""GPU: the CUDA `--fast` path against golden vectors produced by the REAL reference code."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
STEP_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "step_*.npz")))
RTOL = 1e-9


def load(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as d:
        return {k: d[k] for k in d.files}


@pytest.mark.parametrize("name", STEP_CASES)
def test_qffa_step_cuda_vs_reference(name):
    from hendrics_b200.efsearch import search_with_qffa_step

    g = load(name)
    mf, fd, fdd, nbin, nprof, n, sfd = g["params"]
    r = search_with_qffa_step(g["times_c"], mf, fd, fdd, nbin=int(nbin), nprof=int(nprof), npfact=2, oversample=8,
                              n=int(n), search_fdot=bool(sfd), return_profiles=True)
    assert np.array_equal(r[3].astype(np.uint32), g["profiles"]), "2-D profile counts must be bit-exact"
    assert np.array_equal(r[0], g["fgrid"]) and np.array_equal(r[1], g["fdotgrid"])
    np.testing.assert_allclose(r[2], g["stats"], rtol=RTOL)


def test_search_with_qffa_cuda_vs_reference():
    from hendrics_b200.efsearch import search_with_qffa

    g = load("qffa_full")
    np.random.seed(1234)
    r = search_with_qffa(g["times"], 9.85, 9.95, nbin=16, n=2, oversample=8, silent=True)
    assert np.array_equal(r[0], g["fgrid"]) and np.array_equal(r[1], g["fdotgrid"])
    np.testing.assert_allclose(r[2], g["stats"], rtol=RTOL)
    assert np.array_equal(np.array([r[3], r[4], r[5]]), g["scalars"])
    # identical top-5 candidate list
    from oracle import hx_oracle as O

    assert np.array_equal(O.topk_flat(r[2], 5)[0], O.topk_flat(g["stats"], 5)[0])
    g = load("qffa_full_nofd")
    r = search_with_qffa(g["times"], 9.85, 9.95, nbin=16, n=1, oversample=4, search_fdot=False, silent=True)
    assert np.array_equal(r[0], g["fgrid"])
    np.testing.assert_allclose(r[1], g["stats"], rtol=RTOL)


def test_binned_z_of_reference_profiles():
    """K1's fused epilogue reproduces hendrics/ffa.py z_n_fast_cached on the golden profiles: fold
    synthetic events that land exactly p_j times in bin j."""
    from hendrics_b200 import fold_profiles

    g = load("zn_cached")
    nbin = g["profiles"].shape[1]
    for i, p in enumerate(g["profiles"][:3]):
        p = p.astype(np.int64)
        # events at integer + (j + 0.5)/nbin with f = 1 fall in bin j
        ts = np.concatenate([np.arange(c, dtype=np.float64) + (j + 0.5) / nbin for j, c in enumerate(p)])
        stats, prof = fold_profiles(np.sort(ts), [1.0], [0.0], nbin=nbin, nharm=2)
        assert np.array_equal(prof[0, 0], p)
        assert np.isclose(stats[0, 0], g["z2"][i], rtol=RTOL)
        stats4, _ = fold_profiles(np.sort(ts), [1.0], [0.0], nbin=nbin, nharm=4)
        assert np.isclose(stats4[0, 0], g["z4"][i], rtol=RTOL)


def test_transient_search_cuda_vs_reference():
    from hendrics_b200.efsearch import transient_search

    for name, kw in (("transient_full", dict(nbin=16, n=2, oversample=4, f0=9.85, f1=9.95)),
                     ("transient_np96", dict(nbin=32, nprof=96, n=1, oversample=2, f0=9.88, f1=9.92, fdot=1e-6)),
                     # nprof = 100 -> int(2 ** np.log2(100)) = 99 slices are averaged (the last one is dropped)
                     ("transient_np100", dict(nbin=16, nprof=100, n=2, oversample=2, f0=9.88, f1=9.92))):
        g = load(name)
        kw = dict(kw)
        f0, f1 = kw.pop("f0"), kw.pop("f1")
        np.random.seed(7)
        tr = transient_search(g["times"], f0, f1, **kw)
        assert np.array_equal(tr.nave, g["nave"]) and np.array_equal(tr.freqs, g["freqs"])
        assert np.array_equal(tr.times, g["tgrid"])
        assert tr.stats.shape == g["stats"].shape
        np.testing.assert_allclose(tr.stats, g["stats"], rtol=RTOL)
