#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_ln.py > gpurun_out/f_bench_ln.txt 2>&1; cat gpurun_out/f_bench_ln.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tc_kernel<256, 1, 5>' -s 3 -c 1 -f \
   -o gpurun_out/f_ln python scripts/bench_ln.py > gpurun_out/f_ncu_ln.log 2>&1; tail -3 gpurun_out/f_ncu_ln.log
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py -m gpu -q -x -k "pool or bf16_mode or batch1 or shard" > gpurun_out/f_pytest.log 2>&1; tail -5 gpurun_out/f_pytest.log
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/f_kernels_cfg5.txt 2>&1; head -8 gpurun_out/f_kernels_cfg5.txt | cut -c1-150
