#!/bin/bash
mkdir -p gpurun_out
SAG_UMMA_PAIR=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma_gather" > gpurun_out/c10_pair_test.log 2>&1
rc=$?; echo "pair tma test exit $rc"; tail -12 gpurun_out/c10_pair_test.log | cut -c1-250
if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=name,utilization.gpu --format=csv; exit 0; fi
SAG_UMMA_PAIR=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/c10_pytest.log | cut -c1-250
for v in "SAG_UMMA_PAIR=1" "SAG_UMMA_PAIR=0"; do
  env $v SAG_PROF_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_$v.json 2> gpurun_out/c10_bench_$v.err
  echo "$v: $(python -c "import json,sys; d=json.load(open('gpurun_out/c10_bench_$v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['breakdown_ms_per_step'])" 2>&1 | tail -1)"
done
