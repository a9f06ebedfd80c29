#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "layernorm" > gpurun_out/am_pytest.log 2>&1; tail -5 gpurun_out/am_pytest.log
timeout 300 python scripts/bench_ln.py > gpurun_out/am_ln_split.txt 2>&1; cat gpurun_out/am_ln_split.txt
DH_TC_LN_NO_PAIR=1 timeout 300 python scripts/bench_ln.py
ROWS=10240 timeout 300 python scripts/bench_ln.py
ROWS=2560 timeout 300 python scripts/bench_ln.py
ROWS=2560 DH_TC_LN_UNSPLIT=1 timeout 300 python scripts/bench_ln.py
