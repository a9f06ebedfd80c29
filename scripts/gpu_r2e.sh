#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/e_pytest.log
timeout 600 python scripts/kernel_times.py cfg2 > gpurun_out/e_kernels_cfg2.txt 2>&1
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/e_kernels_cfg5.txt 2>&1
tail -12 gpurun_out/e_pytest.log; head -12 gpurun_out/e_kernels_cfg2.txt | cut -c1-150; head -16 gpurun_out/e_kernels_cfg5.txt | cut -c1-150
