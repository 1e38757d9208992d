#!/bin/bash
mkdir -p gpurun_out
{
for i in 1 2; do
echo "--- streamk=0"; timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
echo "--- streamk=1 (eff 80)"; SAG_UMMA_STREAMK=1 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
done
echo "--- streamk=1 (eff 70: conv5 + conv4 only)"; SAG_UMMA_STREAMK=1 SAG_UMMA_STREAMK_EFF=70 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
echo "--- streamk=1 no overlap"; SAG_UMMA_STREAMK=1 SAG_OVERLAP=0 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
echo "--- streamk=0 no overlap"; SAG_OVERLAP=0 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
SAG_UMMA_STREAMK=1 timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c29_a.json >/dev/null 2>&1
timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c29_b.json >/dev/null 2>&1
python - <<'P'
import json
a=json.load(open('gpurun_out/r2c29_a.json'))['layers']; b=json.load(open('gpurun_out/r2c29_b.json'))['layers']
for x,y in zip(a,b):
    if x['cat']=='conv' and x['us']>30: print('%-34s sk %6.1f  nosk %6.1f'%(x['name'],x['us'],y['us']))
P
} > gpurun_out/r2c29.txt 2>&1
