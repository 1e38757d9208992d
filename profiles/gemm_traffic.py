"""DRAM / L2 traffic and duration of the contraction launches of ONE forward from an ncu metrics CSV:
python profiles/gemm_traffic.py <gemm_dram.csv> <out.json>
(csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,... --csv
 -k regex:"gather_gemm_umma_kernel|splitk_reduce"); one forward = the launches from one audio conv1 (the only 32-wide
tile) to the next.  bench.py reads the JSON to fill roofline.traffic."""
import csv
import json
import sys


def main(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    launches = {}
    order = []
    for r in csv.DictReader(lines):
        k = int(r['ID'])
        if k not in launches:
            launches[k] = {'name': r['Kernel Name'], 'grid': r['Grid Size']}
            order.append(k)
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3,
                 'msecond': 1e3, '%': 1.0}.get(u, 1.0)
        launches[k][r['Metric Name']] = v * scale
    starts = [i for i, k in enumerate(order) if 'gather_gemm_umma_kernel<32,' in launches[k]['name'].replace('(int)', '')]
    assert len(starts) >= 1, 'no forward in the capture'
    sel = [launches[k] for k in order[starts[0]:(starts[1] if len(starts) > 1 else len(order))]]
    gemm = [l for l in sel if 'gather_gemm_umma_kernel' in l['name'] or 'halo_conv_umma_kernel' in l['name']]
    red = [l for l in sel if 'splitk_reduce' in l['name']]

    def tot(ls, m):
        return sum(l.get(m, 0.0) for l in ls)

    res = {'kernel': 'gather_gemm_umma_kernel (all %d contraction launches of one B=32 audio+video forward, bf16x3)' % len(gemm),
           'launches': len(gemm),
           'traffic_bytes_per_launch': (tot(gemm, 'dram__bytes_read.sum') + tot(gemm, 'dram__bytes_write.sum')) / max(len(gemm), 1),
           'dram_read_bytes_per_step': tot(gemm, 'dram__bytes_read.sum'), 'dram_write_bytes_per_step': tot(gemm, 'dram__bytes_write.sum'),
           'l2_bytes_per_step': tot(gemm, 'lts__t_bytes.sum'), 'ncu_us_per_step': tot(gemm, 'gpu__time_duration.sum'),
           'tensor_pipe_active_pct_time_weighted': sum(l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0) *
                                                       l.get('gpu__time_duration.sum', 0.0) for l in gemm) / max(tot(gemm, 'gpu__time_duration.sum'), 1e-9),
           'splitk_reduce': {'launches': len(red), 'dram_bytes_per_step': tot(red, 'dram__bytes_read.sum') + tot(red, 'dram__bytes_write.sum'),
                             'ncu_us_per_step': tot(red, 'gpu__time_duration.sum')},
           'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,'
                     'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none (%s)' % path.split('/')[-1]}
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))
    print('per launch: grid, us, tensor%%, dram MB')
    for l in gemm:
        print('  %-60s %-12s %8.1f us %5.1f %% %8.1f MB' % (l['name'].replace('void sag::<unnamed>::', '').replace('(int)', '')[:60], l['grid'],
                                                          l.get('gpu__time_duration.sum', 0), l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0),
                                                          (l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0)) / 1e6))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
