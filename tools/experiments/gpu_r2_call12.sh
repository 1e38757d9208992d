#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "halo" > gpurun_out/r2c12_pytest_halo.log 2>&1
echo "halo test exit $?"; tail -15 gpurun_out/r2c12_pytest_halo.log | cut -c1-300
for halo in 1 0; do
SAG_UMMA_HALO=$halo timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c12_layers_halo$halo.json > gpurun_out/r2c12_bench_halo$halo.json 2> gpurun_out/r2c12_bench_halo$halo.err
echo "bench halo=$halo exit $?"; cut -c1-200 gpurun_out/r2c12_bench_halo$halo.json; tail -2 gpurun_out/r2c12_bench_halo$halo.err | cut -c1-300
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c12_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c12_pytest.log | cut -c1-300
