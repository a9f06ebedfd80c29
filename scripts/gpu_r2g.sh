#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 1 -f \
   -o gpurun_out/g_ln python scripts/bench_ln.py > gpurun_out/g_ncu_ln.log 2>&1; tail -3 gpurun_out/g_ncu_ln.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:select_beam_kernel -s 10 -c 1 -f \
   -o gpurun_out/g_select python bench.py --workload cfg2 --profile-mode --no-cpu-baseline > gpurun_out/g_ncu_select.log 2>&1
ls -la gpurun_out/g_*.ncu-rep
