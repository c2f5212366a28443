"""Where does the wall time of one hendrics_b200.ffa._ffa_device call go (host side)?"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hendrics_b200 import _cabi
from hendrics_b200.ffa import _ffa_device
n_raw = 127_000_000
counts = (torch.rand(n_raw, device="cuda") < 0.1).to(torch.int32)
for _ in range(2):
    _ffa_device(counts, 128)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    _ffa_device(counts, 128)
torch.cuda.synchronize()
print(f"wall per call {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms, kernels {_cabi.ffa_last_kernel()[0]:.2f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    _ffa_device(counts, 128)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
