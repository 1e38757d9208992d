#!/bin/bash
mkdir -p gpurun_out
for sub in 1 2 4; do
  SAG_ISTFT_MIX_SUB=$sub timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c14_$sub.json 2> gpurun_out/c14_$sub.err
  echo "sub $sub: $(python -c "import json,sys; d=json.load(open('gpurun_out/c14_$sub.json')); print(round(d['value'],1), d['roofline']['breakdown_ms_per_step']['istft'])" 2>&1 | tail -1)"
done
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or smoke or istft" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mask_gains|istft_mix|metrics_kernel|stft_kernel|bn_relu" -c 12 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "^\s+(mask_gains|istft_mix|void|metrics|stft|bn_relu)|gpu__time" | paste - - | awk '{print $1, $(NF)}' | head -12
