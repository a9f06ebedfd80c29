#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "headline_config" > gpurun_out/x_pytest.log 2>&1; tail -6 gpurun_out/x_pytest.log
