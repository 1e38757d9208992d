"""DRAM traffic of the dominant kernel family from one `ncu --set full` capture of a bench step:
python profiles/ncu_traffic.py <file.ncu-rep> <kernel regex> <out.json>
Sums dram__bytes_read.sum + dram__bytes_write.sum over the matching launches and divides by their count (per launch,
like roofline.achieved in bench.py); bench.py reads the JSON to fill roofline.traffic."""
import csv
import json
import re
import subprocess
import sys


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main(rep, pattern, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    n, rd, wr, dur = 0, 0.0, 0.0, 0.0
    for r in rows[2:]:
        if not re.search(pattern, r[col['Kernel Name']]):
            continue
        n += 1
        rd += to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
        wr += to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
        d = float(r[col['gpu__time_duration.sum']].replace(',', ''))
        dur += d * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[units[col['gpu__time_duration.sum']]]
    res = {'kernel': pattern, 'launches': n, 'dram_read_bytes_per_launch': rd / max(n, 1), 'dram_write_bytes_per_launch': wr / max(n, 1),
           'traffic_bytes_per_launch': (rd + wr) / max(n, 1), 'ncu_us_per_launch': dur / max(n, 1), 'source': rep.split('/')[-1]}
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res))


if __name__ == '__main__':
    main(*sys.argv[1:4])
