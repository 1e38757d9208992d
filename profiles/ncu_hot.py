"""Hot SASS instructions of the kernels in an ncu report (needs --import-source on / -lineinfo):
python profiles/ncu_hot.py <file.ncu-rep> [top N] -> per kernel: total samples, top instructions by stall samples."""
import csv
import subprocess
import sys


def main(rep, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == 'Kernel Name':
            name = rows[i][1]
            hdr = rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
                if len(rows[j]) >= len(hdr) - 2:
                    body.append(rows[j])
                j += 1
            col = {h: k for k, h in enumerate(hdr)}
            cs, ci, cx = col['# Samples'], col['Instructions Executed'], col['Source']
            stall_cols = [(h, k) for h, k in col.items() if h.startswith('stall_') and 'Not Issued' not in h]
            tot = sum(int(r[cs] or 0) for r in body)
            tot_inst = sum(int(r[ci] or 0) for r in body)
            print('== %s\n   samples %d, warp instructions %d' % (name[:110], tot, tot_inst))
            order = sorted(range(len(body)), key=lambda k: -int(body[k][cs] or 0))[:top]
            for k in sorted(order):
                r = body[k]
                st = sorted(((int(r[c] or 0), h) for h, c in stall_cols), reverse=True)[:2]
                print('   %5d  %5.1f%%  inst %9s  #%-5d %-60s %s' % (int(r[cs] or 0), 100.0 * int(r[cs] or 0) / max(tot, 1), r[ci], k,
                                                               r[cx].strip()[:60], ' '.join('%s=%d' % (h[6:], v) for v, h in st if v)))
            i = j
        else:
            i += 1


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
