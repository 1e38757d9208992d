#!/bin/bash
# compute-sanitizer over the kernels added in the second half of round 2: JPEG decode (memcheck + racecheck), TMA-staged STFT,
# the lanes path of inference_stream
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_jpeg.py tests/test_gpu_parity.py -m gpu -x -q -k "jpeg or decode or stft_matches or stft_other or inference_stream" > gpurun_out/r2san_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|misaligned" gpurun_out/r2san_memcheck.log | head -10
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_jpeg.py tests/test_gpu_parity.py -m gpu -x -q -k "gpu_decode_is_bit or stft_matches" > gpurun_out/r2san_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2san_racecheck.log | head -10
