#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "lstm" > gpurun_out/z_pytest.log 2>&1; tail -3 gpurun_out/z_pytest.log
STEPS=600 timeout 600 python scripts/stress_lstm_stack.py > gpurun_out/z_stress.txt 2>&1; tail -2 gpurun_out/z_stress.txt
timeout 600 python scripts/kernel_times.py cfg2 > gpurun_out/z_kernels_cfg2.txt 2>&1; sed -n 3,9p gpurun_out/z_kernels_cfg2.txt | cut -c1-150
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity.py -m gpu -q -x -k "lstm" > gpurun_out/z_pytest2.log 2>&1; tail -3 gpurun_out/z_pytest2.log
