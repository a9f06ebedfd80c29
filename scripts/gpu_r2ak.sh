#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x > gpurun_out/ak_pytest.log 2>&1; tail -3 gpurun_out/ak_pytest.log
timeout 300 python scripts/bench_conv.py 256 l1 > gpurun_out/ak_conv_auto.txt 2>&1
DH_TC_EPI_GROUPS=1 timeout 300 python scripts/bench_conv.py 256 l1 > gpurun_out/ak_conv_one.txt 2>&1
DH_TC_EPI_GROUPS=3 timeout 300 python scripts/bench_conv.py 256 l1 > gpurun_out/ak_conv_three.txt 2>&1
echo "auto | one | three"
paste -d'|' <(cut -c1-62 gpurun_out/ak_conv_auto.txt) <(cut -c40-62 gpurun_out/ak_conv_one.txt) <(cut -c40-62 gpurun_out/ak_conv_three.txt)
timeout 300 python scripts/bench_trunk.py 512 > gpurun_out/ak_trunk.txt 2>&1; tail -3 gpurun_out/ak_trunk.txt
DH_TC_EPI_GROUPS=1 timeout 300 python scripts/bench_trunk.py 512 > gpurun_out/ak_trunk_one.txt 2>&1; tail -3 gpurun_out/ak_trunk_one.txt
