#!/bin/bash
mkdir -p gpurun_out
python tools/prof_qffa.py 4e7 1 > gpurun_out/plain_q1.log 2>&1 && cat gpurun_out/plain_q1.log &&
ncu --set full --clock-control none --import-source on -k regex:qffa_hist -s 1 -c 1 -o gpurun_out/qffa_c1 -f python tools/prof_qffa.py 4e7 1 > gpurun_out/ncu_q1.log 2>&1
python tools/prof_qffa.py 4e7 8 > gpurun_out/plain_q8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qffa_hist -s 1 -c 1 -o gpurun_out/qffa_c8 -f python tools/prof_qffa.py 4e7 8 > gpurun_out/ncu_q8.log 2>&1
tail -n 2 gpurun_out/ncu_q1.log gpurun_out/ncu_q8.log
