#!/bin/bash
# bench lines + ncu evidence of the K1w (weighted fold) workloads
mkdir -p gpurun_out
for W in w1 w2; do
  python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -c 2500 gpurun_out/bench_$W.json; tail -3 gpurun_out/bench_$W.err
  CMD="python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain_$W.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$W.csv $CMD > gpurun_out/ncu_l_$W.log 2>&1
  $CMD > gpurun_out/plain2_$W.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:wfold_kernel -s 3 -c 1 -o gpurun_out/bench_${W}_top -f $CMD > gpurun_out/ncu_f_$W.log 2>&1
  tail -n 1 gpurun_out/ncu_f_$W.log
done
python bench.py --workload w1 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_w1.json 2>&1; tail -c 800 gpurun_out/bench_ref_w1.json
ls -la gpurun_out/*.ncu-rep
