#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_jpeg.py -m gpu -q -x > gpurun_out/r2c49_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2c49_pytest.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:"jpeg_huffman" -c 2 --csv --log-file gpurun_out/r2c49_jpeg_ncu.csv python tools/jpeg_ncu.py > gpurun_out/r2c49_jpeg_ncu.log 2>&1
grep jpeg_ gpurun_out/r2c49_jpeg_ncu.csv | cut -d, -f13- | cut -c1-200
timeout 300 python tools/jpeg_timing.py 2>&1 | grep -E "sub_bytes=(128|256)"
