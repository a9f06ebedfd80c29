#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
   --log-file gpurun_out/aq_launches_cfg5.csv python bench.py --workload cfg5 --batch 8192 --profile-mode --no-cpu-baseline > gpurun_out/aq_cfg5.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
   --log-file gpurun_out/aq_launches_cfg2.csv python bench.py --workload cfg2 --profile-mode --no-cpu-baseline > gpurun_out/aq_cfg2.log 2>&1
DH_CHAIN_LAYERS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
   --log-file gpurun_out/aq_launches_cfg2_l1.csv python bench.py --workload cfg2 --profile-mode --no-cpu-baseline > gpurun_out/aq_cfg2_l1.log 2>&1
wc -l gpurun_out/aq_launches_*.csv
