#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "bottleneck_tail" > gpurun_out/n_pytest.log 2>&1; tail -8 gpurun_out/n_pytest.log
CHUNKS=0 timeout 600 python scripts/bench_trunk.py > gpurun_out/n_trunk_fused.txt 2>&1; cat gpurun_out/n_trunk_fused.txt
DH_NO_FUSED_TAIL=1 CHUNKS=0 timeout 600 python scripts/bench_trunk.py > gpurun_out/n_trunk_unfused.txt 2>&1; cat gpurun_out/n_trunk_unfused.txt
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "bf16_mode or pool or trunk_pass or shard" > gpurun_out/n_pytest2.log 2>&1; tail -4 gpurun_out/n_pytest2.log
