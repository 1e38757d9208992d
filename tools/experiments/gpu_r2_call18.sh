#!/bin/bash
# role traces of conv4_x with and without stream-K
mkdir -p gpurun_out
for sk in 1 0; do
  echo "=== MT=98 streamk=$sk"
  SAG_UMMA_STREAMK=$sk SAG_UMMA_TRACE=98 SAG_UMMA_TRACE_N=3 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "umma trace" | grep -A4 "KC=36\|KC=18"
done > gpurun_out/r2c18_trace.txt 2>&1
