#!/bin/bash
# round 2 session A: GPU tests, default bench (cfg5 headline), cfg2 bench, vocab microbench random/correlated
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_cfg2.json 2> gpurun_out/a_bench_cfg2.err
for st in 1 8; do
  STRIDE=$st timeout 300 python scripts/bench_vocab.py > gpurun_out/a_vocab_rand_s$st.txt 2>&1
  CORR=1 STRIDE=$st timeout 300 python scripts/bench_vocab.py > gpurun_out/a_vocab_corr_s$st.txt 2>&1
done
tail -15 gpurun_out/a_pytest.log; tail -3 gpurun_out/a_bench.err; cat gpurun_out/a_bench.json | cut -c1-600
