#!/bin/bash
# ncu --set full of every non-GEMM kernel of one forward + metrics (4th step; 28 launches per step)
tag=${1:-sk}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"bn_apply_stats|bn_relu_maxpool_stats|mask_gains|istft_mix|metrics_kernel|stft_kernel|splitk_reduce|space_to_depth" -s 84 -c 28 -o gpurun_out/${tag}_small_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${tag}_ncu.log; du -sh gpurun_out
