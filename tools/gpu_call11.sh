#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/c11_pytest.log | cut -c1-250
for v in "SAG_UMMA_L2PF=1" "SAG_UMMA_L2PF=0"; do
  env $v SAG_PROF_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench_$v.json 2> gpurun_out/c11_bench_$v.err
  echo "$v: $(python -c "import json,sys; d=json.load(open('gpurun_out/c11_bench_$v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['breakdown_ms_per_step'])" 2>&1 | tail -1)"
done
