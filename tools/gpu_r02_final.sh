#!/bin/bash
# what the driver runs at round end, on one GPU: the -m gpu suite, smoke(), the default bench line and its reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python bench.py --impl reference > gpurun_out/bench_ref_default.json 2> gpurun_out/bench_ref_default.err; echo "ref rc=$?"
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_default.json") if l.startswith("{")][-1])
r=d["roofline"]
print("value", f'{d["value"]:.4e}', "e2e", f'{d["e2e"]["value"]:.4e}', "ms/step", round(d["ms_per_step"],1), "frac", round(r["frac"],3), "mix", round(r["frac_of_mix_ceiling"],3), "launches", d["gpu_launches"], "clocks", d["clocks"])
print("cpu", d["cpu_baseline"])
ref=json.loads([l for l in open("gpurun_out/bench_ref_default.json") if l.startswith("{")][-1])
print("ref", f'{ref["value"]:.4e}', "same config:", ref["config"]==d["config"])
PY
