#!/bin/bash
# final evidence of the round: in-situ kernel tables, ncu launch lists (final dispatch), the other configs' bench lines
mkdir -p gpurun_out
timeout 600 python scripts/kernel_times.py cfg5 8192 > gpurun_out/as_kernels_cfg5.txt 2>&1; head -4 gpurun_out/as_kernels_cfg5.txt | cut -c1-150
timeout 600 python scripts/kernel_times.py cfg2 > gpurun_out/as_kernels_cfg2.txt 2>&1; head -4 gpurun_out/as_kernels_cfg2.txt | cut -c1-150
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
   --log-file gpurun_out/as_launches_cfg5.csv python bench.py --workload cfg5 --batch 8192 --profile-mode --no-cpu-baseline > gpurun_out/as_cfg5.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
   --log-file gpurun_out/as_launches_cfg2.csv python bench.py --workload cfg2 --profile-mode --no-cpu-baseline > gpurun_out/as_cfg2.log 2>&1
for w in cfg1 cfg3 cfg4; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/as_bench_$w.json 2> gpurun_out/as_bench_$w.err; cut -c1-200 gpurun_out/as_bench_$w.json
done
