#!/bin/bash
# Multi-GPU lines (run under gpurun --gpus N): configs[1] weak-scaled for >= 2 s of timed region, configs[3] (YT-All-shaped stream)
# and configs[4] (10k-window eval pass) clip-sharded with the single all-gather.
N=${1:-8}
tag=${2:-r2mg}
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
      > gpurun_out/${tag}_${name}_n$N.json 2> gpurun_out/${tag}_${name}_n$N.err
  echo "$name N=$N exit $?"; cut -c1-260 gpurun_out/${tag}_${name}_n$N.json
}
run c2_long --steps 1200 --warmup 5
run c2 --steps 20 --warmup 3
run c4 --config 4 --warmup 3
run c5 --config 5 --warmup 3
timeout 300 python bench.py --steps 1200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_c2_long_n1.json 2> gpurun_out/${tag}_c2_long_n1.err
echo "c2_long N=1 exit $?"; cut -c1-260 gpurun_out/${tag}_c2_long_n1.json
