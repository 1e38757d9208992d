#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_umma_kernel -s 73 -c 1 -o gpurun_out/c5_deconv1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c5_ncu1.log 2>&1
echo "ncu1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_gemm_umma_kernel -s 43 -c 2 -o gpurun_out/c5_conv1_conv2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c5_ncu2.log 2>&1
echo "ncu2 exit $?"
ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
