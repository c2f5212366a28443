"""All hx_microbench kinds on this GPU (issue-rate ceilings used in DESIGN.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hendrics_b200 import _cabi
names = {0: "DFMA inst/s", 1: "DMUL+DADD inst/s", 2: "LDS+IADD+STS rmw/s", 3: "RED.shared atom/s",
         4: "mix (f,fdot): 4 DP + LOP3 + ATOMS units/s", 5: "mix f-only: 2 DP + LOP3 + ATOMS units/s",
         6: "F2I.S64.F64.FLOOR inst/s", 7: "mix (f,fdot), 2/8 floors via F2I units/s", 8: "mix (f,fdot), 4/8 floors via F2I units/s",
         9: "mix (f,fdot), 6/8 floors via F2I units/s", 10: "mix (f,fdot), 8/8 floors via F2I units/s",
         11: "K1 (f,fdot) loop @12 warps/SM, m2 per thread units/s", 12: "K1 (f,fdot) loop @12 warps/SM, m2 from smem tile units/s",
         13: "K1 f-only loop @12 warps/SM, events from smem (LDS.128) units/s", 14: "K1 f-only loop @12 warps/SM, events from __constant__ (LDC) units/s",
         15: "K1 (f,fdot) loop R=4 @6 warps/SM, m2 per thread units/s", 16: "K1 (f,fdot) loop R=4 @6 warps/SM, m2 from smem tile units/s"}
for k, n in names.items():
    v = _cabi.microbench(k, reps=5)
    print(f"{n:48s} {v:.3e}  per SM per clk @1.965GHz: {v/148/1.965e9:.2f}", flush=True)
