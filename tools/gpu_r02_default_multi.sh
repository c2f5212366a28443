#!/bin/bash
# usage: gpu_r02_default_multi.sh N -- the default bench line (C3 slab, strong scaling) on N GPUs
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port $((29620+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_strong_${N}gpu.json 2> gpurun_out/bench_c3_strong_${N}gpu.err; echo "rc=$?"
grep "^{" gpurun_out/bench_c3_strong_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('N', d['n_gpus'], 'value %.4e' % d['value'], 'e2e %.4e' % d['e2e']['value'], 'ms/step %.1f' % d['ms_per_step'])"
