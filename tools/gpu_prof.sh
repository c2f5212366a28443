#!/bin/bash
# ncu captures of the K1 kernel on the two bench shapes (reduced event counts: ncu replays ~40x)
mkdir -p gpurun_out
A="300000 2048 64 128 4"       # C3 shape (f, fdot), nbin=128 -> packed counters
B="300000 400000 1 16 0"       # C2 shape, f only, nbin=16
python tools/prof_fold.py $A > gpurun_out/plain_a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fold_kernel -s 1 -c 1 -o gpurun_out/fold_c3 -f python tools/prof_fold.py $A > gpurun_out/ncu_a.log 2>&1
python tools/prof_fold.py $B > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fold_kernel -s 1 -c 1 -o gpurun_out/fold_c2 -f python tools/prof_fold.py $B > gpurun_out/ncu_b.log 2>&1
tail -n 3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
