#!/bin/bash
# wall-clock view of conv4_x / conv1 launches: CTA entry stagger, role end, CTA exit; plus the new stream-K tests
mkdir -p gpurun_out
{
for mt in 98 6272 392; do
  echo "=== MT=$mt"
  SAG_UMMA_TRACE=$mt SAG_UMMA_TRACE_N=2 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "umma trace" | grep -A6 "KC=36\|KC=4 \|KC=18"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "stream_k" 2>&1 | tail -5
} > gpurun_out/r2c23.txt 2>&1
