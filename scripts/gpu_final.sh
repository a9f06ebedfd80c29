#!/bin/bash
# end-of-round validation: GPU tests, smoke, default bench (cfg5 + cfg2), reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log; tail -4 gpurun_out/final_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -3 gpurun_out/final_smoke.log
( time timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -4 gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.json
