#!/bin/bash
# first GPU contact: parity tests + microbenchmarks
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests/test_fold_gpu.py -x -q -m gpu 2>&1 | tail -30
python - <<'PY' 2>&1 | tee gpurun_out/microbench.txt
from hendrics_b200 import _cabi
names = ["DFMA inst/s", "DMUL+DADD inst/s", "LDS+IADD+STS rmw/s", "RED.shared atom/s"]
for k in range(4):
    v = _cabi.microbench(k, reps=5)
    print(names[k], "%.4g" % v, "per SM per clk @1.965GHz: %.2f" % (v / 148 / 1.965e9))
PY
python tools/quick_bench.py all 2>&1 | tee gpurun_out/quick_bench.txt
