#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mask_gains|istft_mix|metrics_kernel|stft_kernel|bn_relu|space_to_depth|mel_lsd" -c 14 --csv --log-file gpurun_out/c15_k.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/c15_k.csv') if not l.startswith('==')]
for r in csv.DictReader(lines):
    print(r['Kernel Name'][:50], r['Grid Size'], r['Metric Value'], r['Metric Unit'])
PY
