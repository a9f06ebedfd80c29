#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "chain" > gpurun_out/ao_pytest.log 2>&1; tail -3 gpurun_out/ao_pytest.log
CHUNKS=0 timeout 300 python scripts/bench_trunk.py > gpurun_out/ao_trunk_chain.txt 2>&1; cat gpurun_out/ao_trunk_chain.txt | tail -3
CHUNKS=0 DH_NO_CHAIN=1 timeout 300 python scripts/bench_trunk.py > gpurun_out/ao_trunk_nochain.txt 2>&1; cat gpurun_out/ao_trunk_nochain.txt | tail -3
