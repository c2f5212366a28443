#!/bin/bash
# round 2, first GPU trip: the whole -m gpu suite, the default bench line (C3 slab, strong scaling, N = 1), the
# reference arm, then ncu on the per-GPU shape of the N = 8 run (2048 f x 64 fdots).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3_strong_1gpu.json 2> gpurun_out/bench_c3_strong_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c3.json 2>/dev/null
bash tools/gpu_ncu_one.sh c3 fold_kernel "--scaling weak"
