#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py tests/test_gpu_models.py -m gpu -q -x -k "select or beam or fused_vocab or shard or batch1" > gpurun_out/af_pytest.log 2>&1; tail -4 gpurun_out/af_pytest.log
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/af_kernels_cfg5.txt 2>&1; grep -E "select_beam|step " gpurun_out/af_kernels_cfg5.txt | cut -c1-150
timeout 600 python scripts/kernel_times.py cfg2 > gpurun_out/af_kernels_cfg2.txt 2>&1; grep -E "select_beam|step " gpurun_out/af_kernels_cfg2.txt | cut -c1-150
