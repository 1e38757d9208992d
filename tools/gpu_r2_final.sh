#!/bin/bash
# Final round-2 evidence on one B200: whole GPU suite + smoke, bench lines (default = 3 forwards in flight; 1 lane; reference arm; the other
# BASELINE configurations), ncu launch list of a serial forward (one lane, side streams off), ncu pass over the JPEG decode kernels
# and the TMA-staged STFT kernel.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?"; tail -5 gpurun_out/${tag}_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 60 --warmup 3 --layer-table gpurun_out/${tag}_layers.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; cut -c1-300 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --steps 60 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/${tag}_bench_lanes1.json 2> gpurun_out/${tag}_bench_lanes1.err
echo "bench lanes1 exit $?"; cut -c1-200 gpurun_out/${tag}_bench_lanes1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
echo "ref exit $?"; cut -c1-200 gpurun_out/${tag}_bench_reference.json
for cfg in "c1:--config 1 --steps 300" "b10:--batch 10 --steps 200" "c3:--config 3 --steps 40"; do
  name=${cfg%%:*}; args=${cfg#*:}
  timeout 300 python bench.py --warmup 3 --no-cpu-baseline $args > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  echo "$name exit $?"; cut -c1-200 gpurun_out/${tag}_bench_$name.json
done
SAG_LANES=1 SAG_OVERLAP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_l.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread --clock-control none -k regex:"jpeg_" --csv --log-file gpurun_out/${tag}_jpeg_ncu.csv python tools/jpeg_ncu.py > gpurun_out/${tag}_jpeg_ncu.log 2>&1
echo "ncu jpeg exit $?"; tail -3 gpurun_out/${tag}_jpeg_ncu.log
SAG_LANES=1 SAG_OVERLAP=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:"stft_kernel|istft_mix_kernel" -s 6 -c 2 --csv --log-file gpurun_out/${tag}_stft_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_s.log 2>&1
echo "ncu stft exit $?"
timeout 300 python tools/jpeg_timing.py > gpurun_out/${tag}_jpeg_timing.txt 2>&1
echo "jpeg timing exit $?"; tail -12 gpurun_out/${tag}_jpeg_timing.txt | cut -c1-220
du -sh gpurun_out
