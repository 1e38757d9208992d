#!/bin/bash
mkdir -p gpurun_out
for t in 4 9; do
SAG_HALO_TRACE=$t SAG_LANES=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c56_trace$t.err
grep "halo trace" gpurun_out/r2c56_trace$t.err | head -8
done
SAG_HALO_FAST_TAPS=0 SAG_HALO_TRACE=4 SAG_LANES=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c56_trace4s.err
grep "halo trace" gpurun_out/r2c56_trace4s.err | head -4
