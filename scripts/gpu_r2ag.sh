#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resize.py -m gpu -q > gpurun_out/ag_pytest.log 2>&1; tail -5 gpurun_out/ag_pytest.log
