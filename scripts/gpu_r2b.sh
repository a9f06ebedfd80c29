#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/debug_parity.py lstm_labels,xfmr_base > gpurun_out/b_debug.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "pool" > gpurun_out/b_pool.log 2>&1
tail -40 gpurun_out/b_debug.txt; tail -15 gpurun_out/b_pool.log
