#!/bin/bash
# is conv1 bound by the bytes its tiles pull through L2?  per-layer times at bf16x3 (mixed), plain bf16 (half the operand bytes), and with
# the halo kernel forced on (conv1: 4x fewer activation bytes; conv3_x: padded tiles)
mkdir -p gpurun_out
{
SAG_UMMA_STREAMK=0 timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c21_a.json 2>/dev/null | cut -c1-150
SAG_UMMA_STREAMK=0 timeout 300 python bench.py --no-cpu-baseline --precision bf16 --layer-table gpurun_out/r2c21_b.json 2>/dev/null | cut -c1-150
SAG_UMMA_STREAMK=0 SAG_UMMA_HALO=1 timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c21_c.json 2>/dev/null | cut -c1-150
python - <<'P'
import json
t={d:json.load(open('gpurun_out/r2c21_%s.json'%d))['layers'] for d in 'abc'}
for i,x in enumerate(t['a']):
    if x['cat'] in ('conv','deconv','pointwise') and x['us']>15: print('%-34s'%x['name'], x['tile'], ' '.join('%6.1f'%t[d][i]['us'] for d in 'abc'))
P
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
} > gpurun_out/r2c21.txt 2>&1
