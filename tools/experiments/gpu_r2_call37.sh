#!/bin/bash
# lanes by default (3) in bench and inference_stream / deploy_stream: whole GPU suite + default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c37_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2c37_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 60 --warmup 3 > gpurun_out/r2c37_bench.json 2> gpurun_out/r2c37_bench.err
echo "bench exit $?"; cut -c1-300 gpurun_out/r2c37_bench.json; tail -2 gpurun_out/r2c37_bench.err
