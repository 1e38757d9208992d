#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c17_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/c17_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/c17.json 2> gpurun_out/c17.err
python -c "import json,sys; d=json.load(open('gpurun_out/c17.json')); print(round(d['value'],1), d['ms_per_step'], round(d['e2e']['value'],1), d['roofline']['breakdown_ms_per_step'])"
