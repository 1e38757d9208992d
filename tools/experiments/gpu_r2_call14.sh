#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c14_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c14_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c14_layers.json > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err
echo "bench exit $?"; cut -c1-200 gpurun_out/r2c14_bench.json; tail -2 gpurun_out/r2c14_bench.err | cut -c1-300
timeout 300 python bench.py --precision mixed --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2c14_bench_mixed.json 2> gpurun_out/r2c14_bench_mixed.err
echo "bench mixed exit $?"; cut -c1-200 gpurun_out/r2c14_bench_mixed.json
