#!/bin/bash
# re-entry check: whole -m gpu suite, smoke(), one bench line with the per-layer table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c34_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2c34_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c34_smoke.log 2>&1
echo "smoke exit $?"; tail -6 gpurun_out/r2c34_smoke.log | cut -c1-200
timeout 300 python bench.py --steps 30 --warmup 3 --layer-table gpurun_out/r2c34_layers.json > gpurun_out/r2c34_bench.json 2> gpurun_out/r2c34_bench.err
echo "bench exit $?"; cut -c1-1500 gpurun_out/r2c34_bench.json; tail -2 gpurun_out/r2c34_bench.err
