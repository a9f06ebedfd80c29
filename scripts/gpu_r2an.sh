#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel --launch-skip 2 -c 1 -f -o gpurun_out/an_ln python scripts/prof_ln.py > gpurun_out/an_ncu.log 2>&1
tail -2 gpurun_out/an_ncu.log
