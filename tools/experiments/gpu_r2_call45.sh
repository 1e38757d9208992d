#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c45_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2c45_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r2c45_bench.json 2> gpurun_out/r2c45_bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c45_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('steps'))"; tail -1 gpurun_out/r2c45_bench.err
