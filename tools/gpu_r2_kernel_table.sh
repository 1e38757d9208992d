#!/bin/bash
# one serial forward + metrics step (one lane, side streams off) under ncu with the roofline metrics of every launch
mkdir -p gpurun_out
SAG_LANES=1 SAG_OVERLAP=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_kernel_table.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_kernel_table.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r2_kernel_table.csv
