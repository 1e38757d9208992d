#!/bin/bash
# Sweep (tile width, K split) of the wave-quantised layers; per-launch times from SAG_PROF_DUMP.
mkdir -p gpurun_out
i=0
while IFS= read -r cfg; do
  i=$((i+1))
  SAG_PROF_DUMP=1 SAG_UMMA_FORCE="$cfg" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/sw_$i.json 2> gpurun_out/sw_$i.err
  echo "run $i [$cfg]: $(python -c "import json; d=json.load(open('gpurun_out/sw_$i.json')); print(round(d['value'],1), d['roofline']['breakdown_ms_per_step'])" 2>&1 | tail -1)"
done < tools/sweep_cfgs.txt
