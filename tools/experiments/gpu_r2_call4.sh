#!/bin/bash
# round 2, call 4: how many 2-CTA clusters of the contraction kernel are co-resident; pair mode with the grid clamped to that
mkdir -p gpurun_out
SAG_UMMA_PAIR=1 SAG_UMMA_DEBUG=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c4_layers_pair1.json > gpurun_out/r2c4_bench_pair1.json 2> gpurun_out/r2c4_bench_pair1.err
echo "bench pair=1 exit $?"; cut -c1-200 gpurun_out/r2c4_bench_pair1.json; grep umma gpurun_out/r2c4_bench_pair1.err | head
for mt in 1568 392 49; do
SAG_UMMA_PAIR=1 SAG_UMMA_TRACE=$mt timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c4_trace_pair_$mt.err
grep "umma trace" gpurun_out/r2c4_trace_pair_$mt.err
SAG_UMMA_PAIR=0 SAG_UMMA_TRACE=$mt timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/r2c4_trace_single_$mt.err
grep "umma trace" gpurun_out/r2c4_trace_single_$mt.err
done
nvidia-smi --query-gpu=name,mig.mode.current,compute_mode --format=csv
