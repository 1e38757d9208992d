#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "halo or reproducible or graph or stream_k" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_bench_config.py -x -q 2>&1 | tail -3
run() { echo "--- $1"; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c26_l.json 2>/dev/null | cut -c1-170
python - <<'P'
import json
a=json.load(open('gpurun_out/r2c26_l.json'))['layers']
print(' '.join('%s=%.1f'%(x['name'].split('/')[-2] if '/' in x['name'] else x['name'][:12], x['us']) for x in a if 'conv1/conv' in x['name'] or 'ingest' in x['name'] or 'conv2_1/conv_1' in x['name']))
P
}
run "default" X=1
run "default again" X=1
run "im2col conv1 (previous)" SAG_UMMA_HALO_CONV1=0
SAG_UMMA_TRACE=1 true
} > gpurun_out/r2c26.txt 2>&1
