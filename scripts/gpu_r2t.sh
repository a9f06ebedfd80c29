#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py -m gpu -q -x -k "stem or uint8 or bf16_mode or trunk_pass" > gpurun_out/t_pytest.log 2>&1; tail -4 gpurun_out/t_pytest.log
timeout 300 python scripts/bench_stem.py > gpurun_out/t_stem.txt 2>&1; tail -6 gpurun_out/t_stem.txt
CHUNKS=0 timeout 600 python scripts/bench_trunk.py > gpurun_out/t_trunk.txt 2>&1; cat gpurun_out/t_trunk.txt
