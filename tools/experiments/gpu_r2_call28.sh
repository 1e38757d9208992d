#!/bin/bash
mkdir -p gpurun_out
{
for i in 1 2; do
echo "--- stream=1"; SAG_BN_STREAM=1 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
echo "--- stream=0"; SAG_BN_STREAM=0 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
done
SAG_BN_STREAM=1 timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c28_a.json >/dev/null 2>&1
SAG_BN_STREAM=0 timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c28_b.json >/dev/null 2>&1
python - <<'P'
import json
a=json.load(open('gpurun_out/r2c28_a.json'))['layers']; b=json.load(open('gpurun_out/r2c28_b.json'))['layers']
sa=sb=0
for x,y in zip(a,b):
    if x['cat']=='pointwise': sa+=x['us']; sb+=y['us']; print('%-34s %6.1f %6.1f'%(x['name'],x['us'],y['us']))
print('pointwise total', sa, sb)
P
} > gpurun_out/r2c28.txt 2>&1
