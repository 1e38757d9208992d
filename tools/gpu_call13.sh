#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/c13_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
SAG_PROF_DUMP=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
python -c "import json,sys; d=json.load(open('gpurun_out/c13_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['breakdown_ms_per_step'])"
