#!/bin/bash
# stream-K v2 (pipelined raw store, prefetched fix-up, conv5 on 256-wide pairs): parity, bench, layer table, traces
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "stream_k or tcgen05 or reproducible" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_bench_config.py -x -q 2>&1 | tail -5
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-180
SAG_UMMA_STREAMK=0 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-180
done
timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c19_layers.json > /dev/null 2>&1
SAG_UMMA_STREAMK=0 timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c19_layers_nosk.json > /dev/null 2>&1
python - <<'P'
import json
a=json.load(open('gpurun_out/r2c19_layers.json'))['layers']; b=json.load(open('gpurun_out/r2c19_layers_nosk.json'))['layers']
for x,y in zip(a,b):
    if x['cat']=='conv' and x['us']>30: print(x['name'], 'sk', round(x['us'],1), x['tile'], 'nosk', round(y['us'],1), y['tile'])
P
for mt in 98 25; do
  echo "=== MT=$mt"
  SAG_UMMA_TRACE=$mt SAG_UMMA_TRACE_N=2 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "umma trace" | grep -A4 "KC=36\|KC=72"
done
} > gpurun_out/r2c19.txt 2>&1
