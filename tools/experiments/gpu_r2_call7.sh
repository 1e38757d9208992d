#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c7_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c7_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c7_layers.json > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
echo "bench exit $?"; cut -c1-200 gpurun_out/r2c7_bench.json; tail -2 gpurun_out/r2c7_bench.err
timeout 300 python bench.py --config 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c7_bench_c1.json 2> gpurun_out/r2c7_bench_c1.err
echo "config 1 exit $?"; cut -c1-200 gpurun_out/r2c7_bench_c1.json; tail -2 gpurun_out/r2c7_bench_c1.err
timeout 300 python bench.py --batch 10 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c7_bench_b10.json 2> gpurun_out/r2c7_bench_b10.err
echo "B=10 exit $?"; cut -c1-200 gpurun_out/r2c7_bench_b10.json; tail -2 gpurun_out/r2c7_bench_b10.err
SAG_BENCH_GRAPH=0 timeout 300 python bench.py --batch 10 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c7_bench_b10_eager.json 2> gpurun_out/r2c7_bench_b10_eager.err
echo "B=10 eager exit $?"; cut -c1-200 gpurun_out/r2c7_bench_b10_eager.json
SAG_UMMA_TRACE=256 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c7_trace_256.err
grep "umma trace" gpurun_out/r2c7_trace_256.err
SAG_UMMA_TMA_STORE=0 SAG_UMMA_TRACE=256 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c7_trace_256_lsu.err
grep "umma trace" gpurun_out/r2c7_trace_256_lsu.err
