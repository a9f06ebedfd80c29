#!/bin/bash
# session 1: new parity tests, vocab microbench (random vs correlated rows, strides), cfg5 / cfg2 bench with stage ranges
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/s1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s1_pytest.log
for st in 1 4 8; do
  STRIDE=$st timeout 300 python scripts/bench_vocab.py > gpurun_out/s1_vocab_rand_s$st.txt 2>&1
  CORR=1 STRIDE=$st timeout 300 python scripts/bench_vocab.py > gpurun_out/s1_vocab_corr_s$st.txt 2>&1
done
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_cfg2.json 2> gpurun_out/s1_bench_cfg2.err
DH_VOCAB_STRIDE=1 timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_cfg2_stride1.json 2>> gpurun_out/s1_bench_cfg2.err
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s1_bench_cfg5.json 2> gpurun_out/s1_bench_cfg5.err
tail -5 gpurun_out/s1_pytest.log
