#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_c3_fused -s 2 -c 1 -f -o gpurun_out/q_tail \
   python scripts/bench_tail.py > gpurun_out/q_ncu.log 2>&1; tail -2 gpurun_out/q_ncu.log
