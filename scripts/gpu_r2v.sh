#!/bin/bash
mkdir -p gpurun_out
# attn_row_kernel inside the cfg5 decode loop (8192 images): the 60th launch is layer 0 of beam step ~20
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_row_kernel -s 60 -c 1 -f \
   -o gpurun_out/v_attn python bench.py --workload cfg5 --batch 8192 --profile-mode --no-cpu-baseline > gpurun_out/v_ncu.log 2>&1; tail -2 gpurun_out/v_ncu.log
