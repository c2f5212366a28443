#!/bin/bash
# ncu captures of the K1 kernel: nbin=128 (f,fd) and nbin=16 f-only
mkdir -p gpurun_out
A="200000 4096 64 128 4 1"
B="500000 100000 1 16 0 4"
python tools/prof_fold.py $A > gpurun_out/plain_a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fold_kernel -s 1 -c 1 -o gpurun_out/fold_c3 -f python tools/prof_fold.py $A > gpurun_out/ncu_a.log 2>&1
python tools/prof_fold.py $B > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fold_kernel -s 1 -c 1 -o gpurun_out/fold_c2 -f python tools/prof_fold.py $B > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
ls -la gpurun_out
