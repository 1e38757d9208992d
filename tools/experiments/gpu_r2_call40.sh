#!/bin/bash
# with three forwards in flight: is stream-K / the in-forward side streams still worth their overhead?
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/r2c40_$tag.json 2> gpurun_out/r2c40_$tag.err
  echo "$tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c40_$tag.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"; tail -1 gpurun_out/r2c40_$tag.err; }
run base A=1
run nostreamk SAG_UMMA_STREAMK=0
run nooverlap SAG_BENCH_OPTS=overlap=0
run both SAG_UMMA_STREAMK=0 SAG_BENCH_OPTS=overlap=0
run lanes4 SAG_LANES=4
run base2 A=1
