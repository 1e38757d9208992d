#!/bin/bash
# round 2, call 2: suite after the plumbing batch (uint8 ingest, batch planning in sag_workspace_bytes, fixed-order BN sums,
# profile records), bench default + configs 1/3/4/5 short
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c2_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2c2_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 3 --layer-table gpurun_out/r2c2_layers.json > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench exit $?"; cut -c1-600 gpurun_out/r2c2_bench.json; tail -3 gpurun_out/r2c2_bench.err
for c in 1 3 5; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_c$c.json 2> gpurun_out/r2c2_bench_c$c.err
  echo "config $c exit $?"; cut -c1-400 gpurun_out/r2c2_bench_c$c.json; tail -2 gpurun_out/r2c2_bench_c$c.err
done
timeout 300 python bench.py --frames f32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_f32frames.json 2> gpurun_out/r2c2_bench_f32frames.err
echo "f32 frames exit $?"; cut -c1-300 gpurun_out/r2c2_bench_f32frames.json
