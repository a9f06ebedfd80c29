#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s2_pytest.log
for st in 1 4 8; do
  STRIDE=$st timeout 300 python scripts/bench_vocab.py > gpurun_out/s2_vocab_rand_s$st.txt 2>&1
done
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
grep -c . gpurun_out/s2_bench.json; tail -3 gpurun_out/s2_bench.err
tail -15 gpurun_out/s2_pytest.log
