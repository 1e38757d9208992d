#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c8_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c8_pytest.log | cut -c1-300
for ov in 1 0; do
SAG_OVERLAP=$ov timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2c8_bench_ov$ov.json 2> gpurun_out/r2c8_bench_ov$ov.err
echo "bench overlap=$ov exit $?"; cut -c1-200 gpurun_out/r2c8_bench_ov$ov.json; tail -2 gpurun_out/r2c8_bench_ov$ov.err
done
timeout 300 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c8_bench_c3.json 2> gpurun_out/r2c8_bench_c3.err
echo "config 3 exit $?"; cut -c1-200 gpurun_out/r2c8_bench_c3.json
timeout 300 python bench.py --config 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c8_bench_c1.json 2> gpurun_out/r2c8_bench_c1.err
echo "config 1 exit $?"; cut -c1-200 gpurun_out/r2c8_bench_c1.json
