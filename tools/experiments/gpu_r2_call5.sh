#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cta_pair or tma_gather or reproducible" > gpurun_out/r2c5_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2c5_pytest.log | cut -c1-300
SAG_UMMA_PAIR=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c5_layers_pair1.json > gpurun_out/r2c5_bench_pair1.json 2> gpurun_out/r2c5_bench_pair1.err
echo "bench pair=1 exit $?"; cut -c1-200 gpurun_out/r2c5_bench_pair1.json
for mt in 1568; do
SAG_UMMA_PAIR=1 SAG_UMMA_TRACE=$mt timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c5_trace_pair_$mt.err
grep "umma trace" gpurun_out/r2c5_trace_pair_$mt.err
done
