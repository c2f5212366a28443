#!/bin/bash
# ncu evidence for ONE bench workload: launch list + one full capture of the dominant kernel.
# usage: gpu_ncu_one.sh c3 [kernel-regex] [extra bench.py args, e.g. "--scaling weak"]
W=${1:-c3}; K=${2:-fold_kernel}; EXTRA=${3:-}
mkdir -p gpurun_out
CMD="python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline $EXTRA"
$CMD > gpurun_out/plain_$W.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$W.csv $CMD > gpurun_out/ncu_l_$W.log 2>&1
$CMD > gpurun_out/plain2_$W.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/bench_${W}_top -f $CMD > gpurun_out/ncu_f_$W.log 2>&1
tail -n 1 gpurun_out/ncu_f_$W.log; tail -c 600 gpurun_out/plain2_$W.log
ls -la gpurun_out/*.ncu-rep
