#!/bin/bash
# where does the epilogue pass time go: per-layer times with parts of the TMA-store epilogue switched off (results invalid)
mkdir -p gpurun_out
{
for dbg in 0 1 2 3 7; do
  SAG_UMMA_STREAMK=0 SAG_UMMA_EPI_DEBUG=$dbg timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c20_layers_$dbg.json 2>/dev/null | cut -c1-150
done
python - <<'P'
import json
t={d:json.load(open('gpurun_out/r2c20_layers_%d.json'%d))['layers'] for d in (0,1,2,3,7)}
for i,x in enumerate(t[0]):
    if x['cat'] in ('conv','deconv') and x['us']>25: print('%-34s'%x['name'], x['tile'], ' '.join('%6.1f'%t[d][i]['us'] for d in (0,1,2,3,7)))
P
} > gpurun_out/r2c20.txt 2>&1
