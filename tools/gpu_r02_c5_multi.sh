#!/bin/bash
# usage: gpu_r02_c5_multi.sh N -- the whole config 5 (strong scaling) on N GPUs
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port $((29580+N)) bench.py --gpus $N --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_${N}gpu.json 2> gpurun_out/bench_c5_${N}gpu.err; echo "c5 x$N rc=$?"
grep "^{" gpurun_out/bench_c5_${N}gpu.json | cut -c1-250
