#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_tc.py -m gpu -q -x \
  -k "fused_stem_pool or lstm_stack_one_launch and not 2560 or uint8_pixels_is_bit" > gpurun_out/aa_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/aa_memcheck.log; tail -5 gpurun_out/aa_memcheck.log
timeout 1500 $CS --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_tc.py -m gpu -q -x \
  -k "fused_stem_pool and float16" > gpurun_out/aa_racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/aa_racecheck.log; tail -5 gpurun_out/aa_racecheck.log
