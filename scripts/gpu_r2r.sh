#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; tail -4 gpurun_out/r_bench.err
for w in cfg1 cfg3 cfg4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --quick > gpurun_out/r_bench_$w.json 2> gpurun_out/r_bench_$w.err; cut -c1-250 gpurun_out/r_bench_$w.json
done
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r_bench_ref.json 2> gpurun_out/r_bench_ref.err; cut -c1-400 gpurun_out/r_bench_ref.json
