#!/bin/bash
# JPEG frame decode on the GPU: parity tests vs PIL, the on-disk eval test, timing of a batch of 32 frames
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_jpeg.py tests/test_gpu_parity.py -m gpu -q -x -k "jpeg or decode or folder" > gpurun_out/r2c38_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2c38_pytest.log | cut -c1-300
timeout 300 python tools/jpeg_timing.py > gpurun_out/r2c38_jpeg_timing.txt 2>&1
echo "timing exit $?"; cat gpurun_out/r2c38_jpeg_timing.txt | tail -12
