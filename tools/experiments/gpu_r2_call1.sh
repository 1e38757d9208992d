#!/bin/bash
# round 2, first call: the whole -m gpu suite (incl. the new B=32 / B=10 / B=16 / config-3 parity tests and the un-gated
# stage-method test), smoke(), one bench line with the per-launch CUDA-event dump
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2c1_pytest.log 2>&1
echo "pytest exit $?"; grep -E "config2|config3|passed|failed|FAILED|Error" gpurun_out/r2c1_pytest.log | cut -c1-400 | tail -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c1_smoke.log 2>&1
echo "smoke exit $?"; tail -6 gpurun_out/r2c1_smoke.log | cut -c1-200
SAG_PROF_DUMP=1 timeout 300 python bench.py --steps 30 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench exit $?"; cut -c1-1200 gpurun_out/r2c1_bench.json
