#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/c6_pytest.log
SAG_PROF_DUMP=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
python -c "import json,sys; d=json.load(open('gpurun_out/c6_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['breakdown_ms_per_step'])"
for mt in 256 1568 98; do
  SAG_UMMA_TRACE=$mt timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep "umma trace"
done
