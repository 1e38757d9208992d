#!/bin/bash
mkdir -p gpurun_out
for n in 3 1 3 1; do
timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --lanes $n > gpurun_out/r2c46_bench.json 2> gpurun_out/r2c46_bench.err
echo "bench lanes $n exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c46_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('steps'))"; tail -1 gpurun_out/r2c46_bench.err
done
