#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "chain" > gpurun_out/ao_pytest.log 2>&1; tail -3 gpurun_out/ao_pytest.log
for i in 1 2; do
CHUNKS=0 timeout 300 python scripts/bench_trunk.py 2>&1 | tail -2
CHUNKS=0 DH_CHAIN_LAYERS=1 timeout 300 python scripts/bench_trunk.py 2>&1 | tail -2
CHUNKS=0 DH_NO_CHAIN=1 timeout 300 python scripts/bench_trunk.py 2>&1 | tail -2
done
