#!/bin/bash
# ncu evidence for the bench commands themselves: launch lists + one full capture of the dominant kernel
mkdir -p gpurun_out
for W in c2 c3 c4; do
  CMD="python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline"
  K=fold_kernel; [ $W = c4 ] && K=qffa_lane_kernel
  $CMD > gpurun_out/plain_$W.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$W.csv $CMD > gpurun_out/ncu_l_$W.log 2>&1
  $CMD > gpurun_out/plain2_$W.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/bench_${W}_top -f $CMD > gpurun_out/ncu_f_$W.log 2>&1
  tail -n 1 gpurun_out/ncu_f_$W.log
done
ls -la gpurun_out/*.ncu-rep
