#!/bin/bash
# validation of the session's kernels: sanitizer on the new epilogue / LayerNorm-split / chain paths, full GPU suite, smoke, default bench
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -m gpu -q -x \
   -k "chain or (layernorm_epilogue and not 40960) or conv1x1_dual or gemm_bf16_fp32_out" > gpurun_out/ar_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/ar_memcheck.log; tail -4 gpurun_out/ar_memcheck.log
timeout 1200 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -m gpu -q -x \
   -k "(chain and 56-56-64-False-256-64) or (layernorm_epilogue and 777)" > gpurun_out/ar_racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/ar_racecheck.log; tail -4 gpurun_out/ar_racecheck.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/ar_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/ar_pytest.log; tail -4 gpurun_out/ar_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/ar_smoke.log 2>&1; tail -3 gpurun_out/ar_smoke.log
( time timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/ar_bench.json 2> gpurun_out/ar_bench.err; tail -4 gpurun_out/ar_bench.err; cut -c1-300 gpurun_out/ar_bench.json
