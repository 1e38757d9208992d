#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
SAG_E2E_REPEAT=6 timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r2c48_bench.json 2> gpurun_out/r2c48_bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c48_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('steps'))"; grep "e2e repeat" gpurun_out/r2c48_bench.err | tr '\n' ' '; echo
done
