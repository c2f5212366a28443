#!/bin/bash
# round 2: whole GPU suite, smoke(), C5 (full config) and K4 bench lines at N = 1, ncu of K4
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; echo "c5 rc=$?"; cut -c1-300 gpurun_out/bench_c5_1gpu.json
python bench.py --workload zn --steps 5 --warmup 3 > gpurun_out/bench_zn.json 2> gpurun_out/bench_zn.err; echo "zn rc=$?"
python bench.py --workload zn4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zn4.json 2>/dev/null
python bench.py --workload zn1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zn1.json 2>/dev/null
python bench.py --impl reference --workload zn --steps 2 --warmup 1 > gpurun_out/bench_ref_zn.json 2>/dev/null
bash tools/gpu_ncu_one.sh zn zn_unbinned_kernel
