#!/bin/bash
# Final multi-GPU lines (gpurun --gpus 8): configs[1] weak-scaled over >= 1.5 s of timed region, the driver's default 20-step line,
# configs[4] (10k-window eval pass); three forwards in flight per rank.
N=${1:-8}
tag=${2:-r2fmg}
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
      > gpurun_out/${tag}_${name}_n$N.json 2> gpurun_out/${tag}_${name}_n$N.err
  echo "$name N=$N exit $?"; cut -c1-260 gpurun_out/${tag}_${name}_n$N.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/${tag}_${name}_n$N.json
}
run c2_long --steps 1200 --warmup 5
run c2 --steps 20 --warmup 3
run c5 --config 5 --warmup 3
