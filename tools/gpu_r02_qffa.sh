#!/bin/bash
# round 2: fused K2 -> harmonic coefficients (--fast) A/B, K4 after the constant-bank change
mkdir -p gpurun_out
python -m pytest tests/test_qffa_gpu.py tests/test_golden_gpu.py tests/test_zn_topk_gpu.py tests/test_cli_gpu.py -x -q > gpurun_out/pytest_qffa.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_qffa.log
python -m pytest tests/test_fullsize_gpu.py -x -q -k c4 > gpurun_out/pytest_c4.log 2>&1; echo "pytest c4 rc=$?"; tail -n 2 gpurun_out/pytest_c4.log
HX_QFFA_FUSE=0 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_unfused.json 2> gpurun_out/bench_c4_unfused.err; echo "unfused rc=$?"
python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "fused rc=$?"
for f in gpurun_out/bench_c4_unfused.json gpurun_out/bench_c4.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"],2), "value", f'{d["value"]:.3e}', "e2e", f'{d["e2e"]["value"]:.3e}', "K2 ms", round(r["kernel_ms"],2), "mix frac", round(r["frac_of_mix_ceiling"],3), "launches", d["gpu_launches"])
PY
done
python bench.py --workload zn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zn_b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_zn_b.json')); print('zn', f'{d[\"value\"]:.3e}', d['roofline']['frac'])"
bash tools/gpu_ncu_one.sh c4 qffa_lane_kernel
