#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_tc.py -m gpu -q -x \
  -k "test_fused_stem_pool" > gpurun_out/ab_racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/ab_racecheck.log; tail -12 gpurun_out/ab_racecheck.log
