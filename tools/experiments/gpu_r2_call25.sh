#!/bin/bash
# conv1 on the halo kernel: resident weights, one exact activation plane for uint8 frames; A/B on one box
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_bench_config.py -x -q -s -k "uint8 or config2_audio_video" 2>&1 | grep -v "^$" | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "halo or reproducible or graph" 2>&1 | tail -3
run() { echo "--- $1"; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c25_l.json 2>/dev/null | cut -c1-170
python - <<'P'
import json
a=json.load(open('gpurun_out/r2c25_l.json'))['layers']
print(' '.join('%s=%.1f'%(x['name'].split('/')[-2] if '/' in x['name'] else x['name'][:12], x['us']) for x in a if 'conv1/conv' in x['name'] or 'ingest' in x['name'] or 'conv2_1/conv_1' in x['name']))
P
}
run "default (halo conv1, resident weights, integer frames, pair)" X=1
run "non-pair conv1" SAG_UMMA_HALO_CONV1_PAIR=0
run "float-split frames" SAG_UMMA_INT_FRAMES=0
run "float-split frames, weights streamed" SAG_UMMA_INT_FRAMES=0 SAG_UMMA_HALO_BRES=0
run "im2col conv1 (previous)" SAG_UMMA_HALO_CONV1=0
} > gpurun_out/r2c25.txt 2>&1
