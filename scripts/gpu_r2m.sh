#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/m_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/m_pytest.log; tail -5 gpurun_out/m_pytest.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --quick > gpurun_out/m_bench_cfg3.json 2> gpurun_out/m_bench_cfg3.err; cut -c1-300 gpurun_out/m_bench_cfg3.json
