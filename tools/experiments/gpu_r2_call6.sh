#!/bin/bash
# round 2, call 6: TMA-store / bulk-run epilogue + pairs on long 256-wide tiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c6_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c6_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c6_layers.json > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
echo "bench exit $?"; cut -c1-200 gpurun_out/r2c6_bench.json; tail -2 gpurun_out/r2c6_bench.err
SAG_UMMA_TMA_STORE=0 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c6_layers_lsu.json > gpurun_out/r2c6_bench_lsu.json 2> gpurun_out/r2c6_bench_lsu.err
echo "bench lsu exit $?"; cut -c1-200 gpurun_out/r2c6_bench_lsu.json
for mt in 6272 1568; do
SAG_UMMA_TRACE=$mt timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c6_trace_$mt.err
grep "umma trace" gpurun_out/r2c6_trace_$mt.err
done
