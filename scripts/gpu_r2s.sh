#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "path_level" > gpurun_out/s_pytest.log 2>&1; tail -4 gpurun_out/s_pytest.log
# compute-sanitizer memcheck over the kernels added this round (small shapes; the tool slows kernels ~50x)
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_tc.py tests/test_resize.py tests/test_gpu_ops.py -m gpu -q -x \
  -k "pool_equals or layernorm_epilogue and not 40960 or bottleneck_tail or cuda_resize_equals or embed_vectorised" > gpurun_out/s_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/s_memcheck.log; tail -6 gpurun_out/s_memcheck.log
timeout 1200 $CS --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_models.py -m gpu -q -x \
  -k "fused_vocab_path and lstm_labels or batch1" > gpurun_out/s_memcheck2.log 2>&1; echo "memcheck rc $?" >> gpurun_out/s_memcheck2.log; tail -6 gpurun_out/s_memcheck2.log
