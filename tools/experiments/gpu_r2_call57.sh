#!/bin/bash
mkdir -p gpurun_out
SAG_UMMA_TRACE=256 SAG_LANES=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c57_trace.err
grep "umma trace" gpurun_out/r2c57_trace.err | head -16
