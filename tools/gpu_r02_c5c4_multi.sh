#!/bin/bash
# usage: gpu_r02_c5c4_multi.sh N   -- C5 (whole config, strong scaling) and C4 (--fast, centres sharded, weak) on N GPUs
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
$TR --master-port $((29560+N)) bench.py --gpus $N --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_${N}gpu.json 2> gpurun_out/bench_c5_${N}gpu.err; echo "c5 x$N rc=$?"
$TR --master-port $((29570+N)) bench.py --gpus $N --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_${N}gpu.json 2> gpurun_out/bench_c4_${N}gpu.err; echo "c4 x$N rc=$?"
for w in c5 c4; do grep "^{" gpurun_out/bench_${w}_${N}gpu.json | cut -c1-250; done
