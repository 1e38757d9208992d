#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/c7_pytest.log
for v in "SAG_PDL=1" "SAG_PDL=0"; do
  env $v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench_$v.json 2> gpurun_out/c7_bench_$v.err
  echo "$v: $(python -c "import json,sys; d=json.load(open('gpurun_out/c7_bench_$v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['breakdown_ms_per_step'])" 2>&1 | tail -1)"
done
