#!/bin/bash
# lanes in inference_stream (e2e), TMA-staged STFT: parity tests, then bench at 1 / 2 / 3 lanes; B=1 and B=10 graphs with lanes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "stream or stft or graph or golden or forward" > gpurun_out/r2c36_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2c36_pytest.log | cut -c1-300
for n in 1 2 3; do
  timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --lanes $n > gpurun_out/r2c36_lanes$n.json 2> gpurun_out/r2c36_lanes$n.err
  echo "lanes $n exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c36_lanes$n.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'stft', d['roofline']['breakdown_ms_per_step']['stft'])"; tail -2 gpurun_out/r2c36_lanes$n.err
done
for cfg in "--config 1" "--batch 10"; do
for n in 1 2 3; do
  timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline $cfg --lanes $n > gpurun_out/r2c36_small.json 2> gpurun_out/r2c36_small.err
  echo "$cfg lanes $n exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c36_small.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"; tail -2 gpurun_out/r2c36_small.err
done
done
