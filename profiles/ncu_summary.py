"""Summarise the kernels of an ncu report: python profiles/ncu_summary.py <file.ncu-rep> -> selected raw metrics + top stalls."""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        # shared-memory bandwidth is what bounds the contraction kernels: wavefronts the tensor core reads
        # (tcgen05.mma operands), the bytes the TMA unit writes into shared memory, the shared pipe's activity
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed_op_global_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    for vals in r[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
        for k in KEYS:
            for h in d:
                if h == k or h.endswith('.' + k):
                    print('%-80s %-16s %s' % (k, d[h][1], d[h][0][:90]))
                    break
        st = []
        for h, v in d.items():
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and v[0]:
                try:
                    st.append((float(v[0].replace(',', '')), h))
                except ValueError:
                    pass
        for v, h in sorted(st, reverse=True)[:8]:
            print('  stall %-60s %.2f' % (h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), v))
        print()


if __name__ == '__main__':
    main(sys.argv[1])
