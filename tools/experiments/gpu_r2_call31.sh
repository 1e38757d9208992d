#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_parity.py -x -q -k "config2 or config3 or reference_batch or reproducible or graph or fused or stream_k" 2>&1 | tail -3
for i in 1 2; do
echo "--- deconv stream-K on"; timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
echo "--- deconv stream-K off"; SAG_UMMA_STREAMK_DECONV=0 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
done
timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c31_a.json >/dev/null 2>&1
SAG_UMMA_STREAMK_DECONV=0 timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c31_b.json >/dev/null 2>&1
python - <<'P'
import json
a=json.load(open('gpurun_out/r2c31_a.json'))['layers']; b=json.load(open('gpurun_out/r2c31_b.json'))['layers']
for x,y in zip(a,b):
    if x['cat'] in ('deconv','fc'): print('%-34s on %6.1f  off %6.1f  tile %s split %s'%(x['name'],x['us'],y['us'],x['tile'],x['split']))
P
} > gpurun_out/r2c31.txt 2>&1
