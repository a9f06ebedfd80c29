#!/bin/bash
mkdir -p gpurun_out
CHUNKS=0 timeout 600 python scripts/bench_trunk.py > gpurun_out/ac_trunk_fwd.txt 2>&1; cat gpurun_out/ac_trunk_fwd.txt
DH_TC_ALTERNATE=1 CHUNKS=0 timeout 600 python scripts/bench_trunk.py > gpurun_out/ac_trunk_alt.txt 2>&1; cat gpurun_out/ac_trunk_alt.txt
CHUNKS=0 timeout 600 python scripts/bench_trunk.py > gpurun_out/ac_trunk_fwd2.txt 2>&1; cat gpurun_out/ac_trunk_fwd2.txt
DH_TC_ALTERNATE=1 CHUNKS=0 timeout 600 python scripts/bench_trunk.py > gpurun_out/ac_trunk_alt2.txt 2>&1; cat gpurun_out/ac_trunk_alt2.txt
