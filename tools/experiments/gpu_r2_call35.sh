#!/bin/bash
# forwards in flight: 1 / 2 / 3 lanes, same box
mkdir -p gpurun_out
for n in 1 2 3 1 2; do
  timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --lanes $n > gpurun_out/r2c35_lanes$n.json 2> gpurun_out/r2c35_lanes$n.err
  echo "lanes $n exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c35_lanes$n.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"; tail -2 gpurun_out/r2c35_lanes$n.err
done
