"""Aggregate an ncu --metrics gpu__time_duration.sum launch list (csv) per kernel.
python tools/launch_summary.py gpurun_out/x/launches.csv [first_launch last_launch]"""
import collections, csv, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
agg = collections.OrderedDict()
for r in rows[lo:hi]:
    v = float(r['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'msecond': 1e3, 'usecond': 1.0, 'nsecond': 1e-3}.get(r['Metric Unit'], 1.0)
    a = agg.setdefault(r['Kernel Name'][:72], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('launches %d..%d of %d, total %.1f us' % (lo, hi, len(rows), tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-74s n=%4d total=%10.1f us mean=%9.1f us share=%.3f' % (k, a[0], a[1], a[1] / a[0], a[1] / tot))
