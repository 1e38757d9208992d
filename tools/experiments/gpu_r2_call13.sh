#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "halo" > gpurun_out/r2c13_pytest_halo.log 2>&1
echo "halo test exit $?"; tail -12 gpurun_out/r2c13_pytest_halo.log | cut -c1-300
for pair in -1 0; do
SAG_UMMA_PAIR=$pair timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c13_layers_pair$pair.json > gpurun_out/r2c13_bench_pair$pair.json 2> gpurun_out/r2c13_bench_pair$pair.err
echo "bench pair=$pair exit $?"; cut -c1-200 gpurun_out/r2c13_bench_pair$pair.json; tail -2 gpurun_out/r2c13_bench_pair$pair.err | cut -c1-300
done
SAG_UMMA_HALO=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c13_layers_halo_forced.json > gpurun_out/r2c13_bench_halo_forced.json 2> gpurun_out/r2c13_bench_halo_forced.err
echo "bench halo forced exit $?"; cut -c1-200 gpurun_out/r2c13_bench_halo_forced.json
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c13_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c13_pytest.log | cut -c1-300
