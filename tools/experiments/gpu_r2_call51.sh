#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r2c51_$tag.json 2> gpurun_out/r2c51_$tag.err
  echo "$tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c51_$tag.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"; tail -1 gpurun_out/r2c51_$tag.err; }
run base A=1
run prio SAG_LANE_PRIO=0,-1,-2
run prio2 SAG_LANE_PRIO=0,-1,-1
run two_prio SAG_LANES=2 SAG_LANE_PRIO=0,-1
run base2 A=1
