#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c_pytest.log
timeout 600 python scripts/kernel_times.py cfg2 > gpurun_out/c_kernels_cfg2.txt 2>&1
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/c_kernels_cfg5.txt 2>&1
tail -8 gpurun_out/c_pytest.log; head -40 gpurun_out/c_kernels_cfg2.txt; head -45 gpurun_out/c_kernels_cfg5.txt
