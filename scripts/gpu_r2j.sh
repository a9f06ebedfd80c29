#!/bin/bash
mkdir -p gpurun_out
for R in 40960 2560; do
  ROWS=$R timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 1 -f -o gpurun_out/j_pass2_$R \
     python scripts/prof_vocab_pass2.py > gpurun_out/j_pass2_$R.log 2>&1; tail -2 gpurun_out/j_pass2_$R.log
done
