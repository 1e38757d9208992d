#!/bin/bash
mkdir -p gpurun_out
for mt in 6272 1568 392 98 256; do
  SAG_UMMA_TRACE=$mt timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep "umma trace" > gpurun_out/c3_trace_$mt.txt
done
cat gpurun_out/c3_trace_*.txt
timeout 900 ncu --set full --clock-control none -k regex:gather_gemm_umma_kernel -s 37 -c 37 -o gpurun_out/c3_gemm_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c3_ncu1.log 2>&1
echo "ncu1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_umma_kernel -s 44 -c 1 -o gpurun_out/c3_conv2x_src -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c3_ncu2.log 2>&1
echo "ncu2 exit $?"
ls -la gpurun_out/*.ncu-rep
