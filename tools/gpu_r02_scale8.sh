#!/bin/bash
# round 2, 8 GPUs: the WHOLE config 3 through efsearch.folding_search, then the default bench (C3 slab, strong scaling) at 8 / 4 / 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29541 tools/run_c3_full.py > gpurun_out/c3_full_8gpu.log 2> gpurun_out/c3_full_8gpu.err; echo "c3 full rc=$?"; tail -c 1500 gpurun_out/c3_full_8gpu.log
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port $((29550+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_strong_${n}gpu.json 2> gpurun_out/bench_c3_strong_${n}gpu.err; echo "bench $n rc=$?"
  cut -c1-260 gpurun_out/bench_c3_strong_${n}gpu.json
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_strong_1gpu.json 2> gpurun_out/bench_c3_strong_1gpu.err; cut -c1-260 gpurun_out/bench_c3_strong_1gpu.json
