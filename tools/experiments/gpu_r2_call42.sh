#!/bin/bash
# CUDA-graph replay at B=32 (host-insensitive e2e?), faster entropy-decoder kernel, lanes in evaluate_batches
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "jpeg or decode or folder or evaluate" > gpurun_out/r2c42_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2c42_pytest.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r2c42_$tag.json 2> gpurun_out/r2c42_$tag.err
  echo "$tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c42_$tag.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"; tail -1 gpurun_out/r2c42_$tag.err; }
run base A=1
run graph SAG_BENCH_GRAPH=2
run base2 A=1
run graph2 SAG_BENCH_GRAPH=2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"jpeg_" -c 3 --csv --log-file gpurun_out/r2c42_jpeg_ncu.csv python tools/jpeg_ncu.py > gpurun_out/r2c42_jpeg_ncu.log 2>&1
grep jpeg_ gpurun_out/r2c42_jpeg_ncu.csv | cut -d, -f5,15- | cut -c1-200
timeout 300 python tools/jpeg_timing.py 2>&1 | grep -E "sub_bytes=256|PIL"
