#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or tma_gather or cta_pair or stft_matches or audio_only_fp32 or deconv2d_tcgen05 or mel or evaluate_rows" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|misaligned" gpurun_out/san_memcheck.log | head -10
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or stft_matches or istft_matches" > gpurun_out/san_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/san_racecheck.log | head -10
