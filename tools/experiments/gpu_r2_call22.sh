#!/bin/bash
# role traces of conv1 (6272 M tiles): bf16x3, plain bf16, epilogue stores / statistics off
mkdir -p gpurun_out
{
for v in "0 mixed" "0 bf16" "7 mixed" "1 mixed"; do
  set -- $v
  echo "=== epi_debug=$1 precision=$2"
  SAG_UMMA_STREAMK=0 SAG_UMMA_EPI_DEBUG=$1 SAG_UMMA_TRACE=6272 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precision $2 2>&1 >/dev/null | grep "umma trace"
done
} > gpurun_out/r2c22.txt 2>&1
