#!/bin/bash
# Round-2 evidence: bench lines (ours + reference arm), ncu launch list, DRAM / L2 traffic of the contraction launches of one
# forward, one --set full capture of a conv2_x launch (halo-resident kernel on CTA pairs; conv1 + 4 conv2_x launches per forward, the 14th halo launch is a conv2_x of the 3rd forward).  ncu runs use SAG_OVERLAP=0 so that the launch order is the serial one.
tag=${1:-r2ev}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 30 --warmup 3 --layer-table gpurun_out/${tag}_layers.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; cut -c1-300 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
echo "ref exit $?"; cut -c1-200 gpurun_out/${tag}_bench_reference.json
SAG_OVERLAP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_l.log 2>&1
echo "ncu launches exit $?"
SAG_OVERLAP=0 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gather_gemm_umma_kernel|halo_conv_umma_kernel|splitk_reduce" -s 129 -c 43 --csv --log-file gpurun_out/${tag}_gemm_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_d.log 2>&1
echo "ncu dram exit $?"
SAG_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo_conv_umma_kernel -s 13 -c 1 -o gpurun_out/${tag}_conv2x_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_f.log 2>&1
echo "ncu full exit $?"
du -sh gpurun_out
