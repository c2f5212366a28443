#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29531 tools/dist_check.py 2>&1 | grep "rank\|Error\|error" | head -10
$TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_8gpu.json 2> gpurun_out/bench_c2_8gpu.err
$TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 3 --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_8gpu.json 2> gpurun_out/bench_c3_8gpu.err
$TR --nproc-per-node 4 --master-port 29534 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_4gpu.json 2> gpurun_out/bench_c2_4gpu.err
for f in gpurun_out/bench_c2_8gpu gpurun_out/bench_c3_8gpu gpurun_out/bench_c2_4gpu; do tail -n 2 $f.err | cut -c1-200; cut -c1-330 $f.json | tail -n 1; done
