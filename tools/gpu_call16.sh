#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or forward" 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/c16.json 2> gpurun_out/c16.err
python -c "import json,sys; d=json.load(open('gpurun_out/c16.json')); print(round(d['value'],1), d['ms_per_step'], round(d['e2e']['value'],1), d['roofline']['breakdown_ms_per_step'])"
bash tools/gpu_call15.sh | head -7
