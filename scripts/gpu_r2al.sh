#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/al_kernels_cfg5.txt 2>&1; head -22 gpurun_out/al_kernels_cfg5.txt | cut -c1-160
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/al_pytest.log 2>&1; tail -4 gpurun_out/al_pytest.log
