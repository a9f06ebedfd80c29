#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "embed" > gpurun_out/i_pytest.log 2>&1; tail -3 gpurun_out/i_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 8 -c 1 -f \
   -o gpurun_out/i_ln python scripts/bench_ln.py > gpurun_out/i_ncu_ln.log 2>&1; tail -2 gpurun_out/i_ncu_ln.log
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/i_kernels_cfg5.txt 2>&1; grep -A30 "idle time" gpurun_out/i_kernels_cfg5.txt | cut -c1-160
