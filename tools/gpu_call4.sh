#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma_gather" > gpurun_out/c4_tma_test.log 2>&1
echo "tma test exit $?"; tail -25 gpurun_out/c4_tma_test.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/c4_pytest.log
for v in "SAG_UMMA_TMA=1" "SAG_UMMA_TMA=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v SAG_PROF_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench_$tag.json 2> gpurun_out/c4_bench_$tag.err
  echo "$tag: $(python -c "import json,sys; d=json.load(open('gpurun_out/c4_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['breakdown_ms_per_step'])" 2>&1 | tail -1)"
done
