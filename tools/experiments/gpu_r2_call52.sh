#!/bin/bash
# conv kernels capped at 128 registers (SAG_CONV_REG128 build): do the batch-norm passes of other lanes co-reside?
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c52_layers_$tag.json > gpurun_out/r2c52_$tag.json 2> gpurun_out/r2c52_$tag.err
  echo "$tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c52_$tag.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['roofline']['breakdown_ms_per_step'])"; tail -1 gpurun_out/r2c52_$tag.err; }
run l3 A=1
run l1 SAG_LANES=1
run l3b A=1
run l4 SAG_LANES=4
timeout 600 python -m pytest tests/test_gpu_bench_config.py -m gpu -q -x > gpurun_out/r2c52_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2c52_pytest.log | cut -c1-300
