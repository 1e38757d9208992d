#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -k "precision_plan or uint8 or evaluate_from or deploy" > gpurun_out/r2c11_pytest.log 2>&1
echo "pytest exit $?"; grep -E "precision plan|passed|failed|Error" gpurun_out/r2c11_pytest.log | cut -c1-200
timeout 300 python bench.py --precision mixed --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c11_layers_mixed.json > gpurun_out/r2c11_bench_mixed.json 2> gpurun_out/r2c11_bench_mixed.err
echo "bench mixed exit $?"; cut -c1-200 gpurun_out/r2c11_bench_mixed.json; tail -2 gpurun_out/r2c11_bench_mixed.err
i=0
for f in "25:512:256:2" "25:512:256:3" "49:256:256:2" "49:256:256:3" "25:512:256:2,49:256:256:2"; do
  i=$((i+1))
  SAG_UMMA_FORCE="$f" timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/r2c11_layers_f$i.json > gpurun_out/r2c11_bench_f$i.json 2> gpurun_out/r2c11_bench_f$i.err
  echo "force $f exit $?"; cut -c1-160 gpurun_out/r2c11_bench_f$i.json; tail -1 gpurun_out/r2c11_bench_f$i.err | cut -c1-200
done
