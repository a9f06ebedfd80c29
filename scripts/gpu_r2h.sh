#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_ln.py > gpurun_out/h_bench_ln.txt 2>&1; cat gpurun_out/h_bench_ln.txt
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/h_pytest.log; tail -4 gpurun_out/h_pytest.log
timeout 600 python scripts/kernel_times.py cfg2 > gpurun_out/h_kernels_cfg2.txt 2>&1; sed -n 3,12p gpurun_out/h_kernels_cfg2.txt | cut -c1-150
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/h_kernels_cfg5.txt 2>&1; sed -n 3,14p gpurun_out/h_kernels_cfg5.txt | cut -c1-150
