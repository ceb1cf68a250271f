"""Per-kernel summary of an ncu launch list (gpu__time_duration.sum, --csv): count, mean, total, share."""
import collections
import csv
import sys

d = collections.defaultdict(list)
for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        name = r['Kernel Name'].split('(')[0].replace('void ', '')
        d[name + ' grid ' + r['Grid Size']].append(float(r['Metric Value'].replace(',', '')) / 1e6)
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print('%-60s n=%3d mean %.3f ms total %7.2f ms  %4.1f %%' % (k[:60], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
print('total %.2f ms over %d launches' % (tot, sum(len(v) for v in d.values())))
