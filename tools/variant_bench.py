"""Time K1 on C3-like shapes for experimental builds (hendrics_b200/variants/libhx_*.so, `make variant`).

Each variant runs in its own process (HX_LIB_PATH); results are checked against the stock library's statistics.
usage: python tools/variant_bench.py [name ...]      (no names: every variant found, stock library first)
"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, zlib
sys.path.insert(0, %r)
import numpy as np, torch
from hendrics_b200 import _cabi
from hendrics_b200.search import fold_stats_device
rng = np.random.default_rng(0)
ts = torch.from_numpy(np.sort(rng.uniform(0, 1e5, 2_000_000))).cuda()
for nbin, n_f, n_fd, nharm in ((128, 2048, 64, 4), (64, 2048, 64, 2), (32, 4096, 32, 2), (16, 100000, 1, 0), (128, 20000, 1, 2)):
    f = torch.from_numpy(9.0 + 5e-6 * np.arange(n_f)).cuda()
    fd = torch.from_numpy((-1.6e-9 + 5e-11 * np.arange(n_fd)) if n_fd > 1 else np.zeros(1)).cuda()
    st = fold_stats_device(ts, f, fd, nbin, nharm)
    best = 1e9
    for _ in range(3):
        st = fold_stats_device(ts, f, fd, nbin, nharm); best = min(best, _cabi.fold_last_kernel()[0])
    k = _cabi.fold_last_kernel()
    crc = zlib.crc32(st.cpu().numpy().tobytes())
    print(f"  nbin={nbin:4d} nf={n_f:6d} nfd={n_fd:3d}: {best:8.2f} ms {2e6*n_f*n_fd/best/1e-3:.3e} ev*tr/s  block={k[2]} smem={k[3]} crc={crc:08x}", flush=True)
''' % ROOT
names = sys.argv[1:]
libs = [("stock", None)]
for p in sorted(glob.glob(os.path.join(ROOT, "hendrics_b200", "variants", "libhx_*.so"))):
    n = os.path.basename(p)[6:-3]
    if not names or n in names:
        libs.append((n, p))
for n, p in libs:
    env = dict(os.environ)
    if p:
        env["HX_LIB_PATH"] = p
    print(f"== {n}", flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=env, timeout=600)
