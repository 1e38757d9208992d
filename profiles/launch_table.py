"""Print the launches of one forward (between two stft_kernel launches) from an ncu launch-list CSV."""
import csv, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lines = [l for l in open(path) if not l.startswith('==')]
rows = [(r['Kernel Name'], r['Grid Size'], float(r['Metric Value'].replace(',', '')) / (1e3 if r['Metric Unit'] == 'ns' else 1.0))
        for r in csv.DictReader(lines) if r['Metric Name'] == 'gpu__time_duration.sum']
starts = [i for i, x in enumerate(rows) if 'stft_kernel' in x[0] and 'istft' not in x[0]]
s, e = starts[which], starts[which + 1]
tot = 0.0
for name, grid, us in rows[s:e]:
    short = name.replace('void sag::<unnamed>::', '').replace('sag::', '')
    short = short[:short.index('(')] if '(' in short else short
    tot += us
    print('%-44s %-16s %9.1f us' % (short[:44], grid, us))
print('total %.1f us' % tot)
