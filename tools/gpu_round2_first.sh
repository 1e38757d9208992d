#!/bin/bash
# First GPU call of the next round: the full -m gpu suite (the streaming-kernel restructure of the end of round 1 has only been
# run in host emulation), one bench line, and the launch list -- compare bn_apply / bn_relu_maxpool / space_to_depth16 with
# profiles/r1_small_kernels_ncu.txt (275 / 89 / 34 us) and the step with profiles/r1_final_bench.json (1.915 ms).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2_pytest.log | cut -c1-300
SAG_TEST_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k stage_methods > gpurun_out/r2_stage.log 2>&1
echo "stage methods exit $?"; tail -15 gpurun_out/r2_stage.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench exit $?"; cut -c1-400 gpurun_out/r2_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_l.log 2>&1
echo "ncu launches exit $?"
