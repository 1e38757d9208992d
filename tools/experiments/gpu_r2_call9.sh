#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c9_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c9_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c9_layers.json > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err
echo "bench exit $?"; cut -c1-200 gpurun_out/r2c9_bench.json; tail -2 gpurun_out/r2c9_bench.err
SAG_UMMA_TRACE=256 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c9_trace_256.err
grep "umma trace" gpurun_out/r2c9_trace_256.err
