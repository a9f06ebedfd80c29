#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 2400 $CS --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_models.py -m gpu -q -x \
  -k "small and (fp32_mode or bf16_mode or perplexity)" > gpurun_out/ad_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/ad_memcheck.log; tail -6 gpurun_out/ad_memcheck.log
timeout 1200 $CS --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_models.py -m gpu -q -x \
  -k "path_level and xfmr_base or char_level and xfmr" > gpurun_out/ad_memcheck2.log 2>&1; echo "memcheck rc $?" >> gpurun_out/ad_memcheck2.log; tail -6 gpurun_out/ad_memcheck2.log
