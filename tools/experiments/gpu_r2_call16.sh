#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c16_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c16_pytest.log | cut -c1-300
for sc in 1 0 1 0; do
SAG_OVERLAP_SC=$sc timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_bench_sc$sc.json 2> gpurun_out/r2c16_bench_sc$sc.err
echo "bench sc=$sc exit $?"; cut -c1-160 gpurun_out/r2c16_bench_sc$sc.json
done
