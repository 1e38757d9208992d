#!/bin/bash
mkdir -p gpurun_out
{
for dbg in 0 1 2 3; do
  echo "=== halo debug $dbg (1 = no TMEM loads in the epilogue, 2 = no staging writes / stores)"
  SAG_HALO_DEBUG=$dbg SAG_HALO_TRACE=4 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "halo trace" | head -4
  SAG_HALO_DEBUG=$dbg SAG_HALO_TRACE=9 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "halo trace" | head -4
done
} > gpurun_out/r2c32.txt 2>&1
