#!/bin/bash
# GPU call: parity tests, A/B bench of the kernel switches, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
for v in default "SAG_UMMA_CONCAT=0" "SAG_UMMA_FIXUP=0" "SAG_UMMA_NARROW_FC=0" "SAG_UMMA_CONCAT=0 SAG_UMMA_FIXUP=0 SAG_UMMA_NARROW_FC=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  if [ "$v" = default ]; then envs=""; else envs="$v"; fi
  env $envs timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_$tag.json 2> gpurun_out/c1_bench_$tag.err
  echo "$tag: $(python -c "import json,sys; d=json.load(open('gpurun_out/c1_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['breakdown_ms_per_step'])" 2>&1 | tail -1)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c1_ncu_bench.log 2>&1
echo "ncu exit $?"
