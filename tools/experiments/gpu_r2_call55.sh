#!/bin/bash
# halo kernel, resident weights (conv1): all vertical taps issued by one elected block -- A/B with the general loop
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_parity.py -m gpu -q -x -k "uint8 or config2 or halo or reproducible" > gpurun_out/r2c55_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2c55_pytest.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c55_layers_$tag.json > gpurun_out/r2c55_$tag.json 2> gpurun_out/r2c55_$tag.err
  echo "$tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c55_$tag.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['roofline']['breakdown_ms_per_step']['conv'], round(d['roofline']['frac'],4))
t=json.load(open('gpurun_out/r2c55_layers_$tag.json'))['layers']; print([ (r['name'], round(r['us'],1)) for r in t if 'conv1/conv' in r['name'] or 'ingest' in r['name']])"; tail -1 gpurun_out/r2c55_$tag.err; }
run fast A=1
run slow SAG_HALO_FAST_TAPS=0
run fast2 A=1
run slow2 SAG_HALO_FAST_TAPS=0
