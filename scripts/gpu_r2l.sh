#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_trunk.py > gpurun_out/l_trunk.txt 2>&1; cat gpurun_out/l_trunk.txt
