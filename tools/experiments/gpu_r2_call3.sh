#!/bin/bash
# round 2, call 3: fixed-point BN sums (deterministic, no tail), CTA pairs with direct 2-SM TMA signalling
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c3_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2c3_pytest.log | cut -c1-300
for pair in 0 1; do
  SAG_UMMA_PAIR=$pair timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c3_layers_pair$pair.json > gpurun_out/r2c3_bench_pair$pair.json 2> gpurun_out/r2c3_bench_pair$pair.err
  echo "bench pair=$pair exit $?"; cut -c1-200 gpurun_out/r2c3_bench_pair$pair.json; tail -3 gpurun_out/r2c3_bench_pair$pair.err
done
