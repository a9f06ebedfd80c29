#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_pool_kernel -s 2 -c 1 -f -o gpurun_out/u_stem \
   python scripts/bench_stem.py > gpurun_out/u_ncu.log 2>&1; tail -2 gpurun_out/u_ncu.log
