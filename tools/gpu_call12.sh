#!/bin/bash
mkdir -p gpurun_out
for kb in 20 40 72; do
  SAG_ISTFT_SMEM_KB=$kb timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c12_$kb.json 2> gpurun_out/c12_$kb.err
  echo "smem cap $kb KB: $(python -c "import json,sys; d=json.load(open('gpurun_out/c12_$kb.json')); print(round(d['value'],1), d['roofline']['breakdown_ms_per_step']['istft'])" 2>&1 | tail -1)"
done
