#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -q -x -k "attention or attn or xfmr or bf16_mode" > gpurun_out/w_pytest.log 2>&1; tail -4 gpurun_out/w_pytest.log
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/w_kernels_cfg5.txt 2>&1; sed -n 3,16p gpurun_out/w_kernels_cfg5.txt | cut -c1-150
