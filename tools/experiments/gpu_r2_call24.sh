#!/bin/bash
# twelve-warp TMA-mode layout + deferred statistics on 64-wide tiles: full GPU suite, bench, layer table
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200; done
timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c24_layers.json > /dev/null 2>&1
python - <<'P'
import json
a=json.load(open('gpurun_out/r2c24_layers.json'))['layers']
for x in a:
    if x['us']>25: print('%-36s %6.1f %s'%(x['name'], x['us'], x['tile']))
P
} > gpurun_out/r2c24.txt 2>&1
