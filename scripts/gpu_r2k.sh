#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_models.py tests/test_gpu_ops.py -m gpu -q -x -k "vocab or fused or select or perplexity or sampled" > gpurun_out/k_pytest.log 2>&1; tail -4 gpurun_out/k_pytest.log
for R in 2560 40960; do
  ROWS=$R timeout 300 python scripts/bench_vocab2.py > gpurun_out/k_vocab_$R.txt 2>&1; cat gpurun_out/k_vocab_$R.txt
done
