#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x > gpurun_out/aj_pytest.log 2>&1; tail -3 gpurun_out/aj_pytest.log
timeout 300 python scripts/bench_conv.py 256 > gpurun_out/aj_conv_auto.txt 2>&1
DH_TC_EPI_GROUPS=1 timeout 300 python scripts/bench_conv.py 256 > gpurun_out/aj_conv_one.txt 2>&1
DH_TC_EPI_GROUPS=2 timeout 300 python scripts/bench_conv.py 256 > gpurun_out/aj_conv_two.txt 2>&1
timeout 300 python scripts/bench_dec_gemm.py > gpurun_out/aj_dec_auto.txt 2>&1
DH_TC_EPI_GROUPS=1 timeout 300 python scripts/bench_dec_gemm.py > gpurun_out/aj_dec_one.txt 2>&1
DH_TC_EPI_GROUPS=2 timeout 300 python scripts/bench_dec_gemm.py > gpurun_out/aj_dec_two.txt 2>&1
echo "auto | one | two"
paste -d'|' <(cut -c1-62 gpurun_out/aj_conv_auto.txt) <(cut -c40-62 gpurun_out/aj_conv_one.txt) <(cut -c40-62 gpurun_out/aj_conv_two.txt)
paste -d'|' gpurun_out/aj_dec_auto.txt <(cut -c38- gpurun_out/aj_dec_one.txt) <(cut -c38- gpurun_out/aj_dec_two.txt)
