#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "bottleneck_tail" > gpurun_out/p_pytest.log 2>&1; tail -4 gpurun_out/p_pytest.log
timeout 300 python scripts/bench_tail.py > gpurun_out/p_tail.txt 2>&1; cat gpurun_out/p_tail.txt
CHUNKS=0 timeout 600 python scripts/bench_trunk.py > gpurun_out/p_trunk_fused.txt 2>&1; cat gpurun_out/p_trunk_fused.txt
