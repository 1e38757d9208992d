#!/bin/bash
mkdir -p gpurun_out
{
SAG_HALO_TRACE=4 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "halo trace" | head -4
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-170
} > gpurun_out/r2c33.txt 2>&1
