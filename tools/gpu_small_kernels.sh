#!/bin/bash
# ncu of every non-GEMM kernel of one forward + metrics (4th step; 28 launches per step).
# COST: with --set full this took 200 s of box time in round 1 (ncu saves / restores the >1 GB workspace around each of the ~40 replay
# passes of each kernel).  The sections below need ~10 passes; pass `full` as the second argument only when the source page is needed.
tag=${1:-sk}
mode=${2:-sections}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
if [ "$mode" = full ]; then sel="--set full"; else sel="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats"; fi
timeout 600 ncu $sel --clock-control none -k regex:"bn_apply_stats|bn_relu_maxpool_stats|mask_gains|istft_mix|metrics_kernel|stft_kernel|splitk_reduce|space_to_depth" -s 84 -c 28 -o gpurun_out/${tag}_small -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${tag}_ncu.log | cut -c1-300; du -sh gpurun_out
