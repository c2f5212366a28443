"""One FFA on a C4-sized light curve (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hendrics_b200 import _cabi
from hendrics_b200.ffa import _ffa_device
n_raw = 127_000_000
g = torch.Generator(device="cuda"); g.manual_seed(0)
counts = (torch.rand(n_raw, device="cuda", generator=g) < 0.1).to(torch.int32)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for _ in range(2):
    per, st = _ffa_device(counts, P)
print(_cabi.ffa_last_kernel())
