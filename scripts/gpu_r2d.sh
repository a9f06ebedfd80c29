#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resize.py -m gpu -q > gpurun_out/d_resize.log 2>&1; tail -5 gpurun_out/d_resize.log
# ncu --set full of the selection + beam-step launch and of the stacked LSTM step / vocab passes inside the cfg2 decode loop
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:select_beam_kernel -s 10 -c 1 -f \
   -o gpurun_out/d_select python bench.py --workload cfg2 --profile-mode --no-cpu-baseline > gpurun_out/d_ncu_select.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 95 -c 5 -f \
   -o gpurun_out/d_gemm python bench.py --workload cfg2 --profile-mode --no-cpu-baseline > gpurun_out/d_ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/d_ncu_select.log; tail -3 gpurun_out/d_ncu_gemm.log
