#!/bin/bash
mkdir -p gpurun_out
{
for taps in 4 9; do
  SAG_HALO_TRACE=$taps timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "halo trace"
done
} > gpurun_out/r2c27.txt 2>&1
