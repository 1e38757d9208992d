#!/bin/bash
# stream-K: parity (standalone + B=32 forward), then bench with / without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "stream_k or tcgen05" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_bench_config.py -x -q 2>&1 | tail -8
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err; echo "bench exit $?"; cut -c1-200 gpurun_out/r2c17_bench.json
SAG_UMMA_STREAMK=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c17_bench_nosk.json 2>/dev/null; cut -c1-200 gpurun_out/r2c17_bench_nosk.json
SAG_UMMA_STREAMK_EFF=80 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c17_bench_sk80.json 2>/dev/null; cut -c1-200 gpurun_out/r2c17_bench_sk80.json
timeout 300 python bench.py --no-cpu-baseline --layer-table gpurun_out/r2c17_layers.json > /dev/null 2>&1
python - <<'P'
import json
d=json.load(open('gpurun_out/r2c17_layers.json'))
for r in d['layers']:
    if r['cat'] in ('conv',) and r['us']>30: print(r['name'], round(r['us'],1), r['tile'], r['split'])
P
