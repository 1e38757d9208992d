#!/bin/bash
# 128-register conv kernels + shared memory reserved for co-resident CTAs
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r2c53_$tag.json 2> gpurun_out/r2c53_$tag.err
  echo "$tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c53_$tag.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['roofline']['breakdown_ms_per_step'])"; tail -1 gpurun_out/r2c53_$tag.err; }
run r0 A=1
run r4k SAG_UMMA_SMEM_RESERVE=4096
run r8k SAG_UMMA_SMEM_RESERVE=8192
run r8k_l4 SAG_UMMA_SMEM_RESERVE=8192 SAG_LANES=4
run r8k_l1 SAG_UMMA_SMEM_RESERVE=8192 SAG_LANES=1
