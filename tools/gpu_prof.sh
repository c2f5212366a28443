#!/bin/bash
# ncu capture of K1 on the nbin=128 f-only shape (C1/C5 class)
mkdir -p gpurun_out
A="400000 262144 1 128 2"
python tools/prof_fold.py $A > gpurun_out/plain_a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fold_kernel -s 1 -c 1 -o gpurun_out/fold_n128f -f python tools/prof_fold.py $A > gpurun_out/ncu_a.log 2>&1
tail -n 2 gpurun_out/ncu_a.log
