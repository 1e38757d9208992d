#!/bin/bash
# quick GPU check after a kernel change: the reference-golden and fused-path parity tests, then one bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_goldens.py tests/test_gpu_parity.py -m gpu -x -q -k "reference or fused or deploy or forward" > gpurun_out/q_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/q_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 3 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
echo "bench exit $?"; cut -c1-1500 gpurun_out/q_bench.json
