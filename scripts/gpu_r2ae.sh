#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:select_beam_kernel -s 20 -c 1 -f \
   -o gpurun_out/ae_select python bench.py --workload cfg5 --batch 8192 --profile-mode --no-cpu-baseline > gpurun_out/ae_ncu.log 2>&1; tail -2 gpurun_out/ae_ncu.log
