"""Summarise an ncu launch-list CSV (gpu__time_duration.sum per launch): per-kernel totals and shares.
usage: python profiles/summarize_launches.py gpurun_out/launches_TAG.csv [--seq N]   (--seq prints the first N launches)"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
hdr = rows[0]; ik = hdr.index('Kernel Name'); im = hdr.index('Metric Name'); iv = hdr.index('Metric Value'); iid = hdr.index('ID')
d = collections.OrderedDict(); extra = collections.defaultdict(dict)
for r in rows[1:]:
    if r[im] == 'gpu__time_duration.sum':
        d[r[iid]] = (r[ik].split('(')[0].split('::')[-1], float(r[iv].replace(',', '')) / 1e3)
    else:
        extra[r[iid]][r[im]] = r[iv]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, (n, t) in d.items():
    agg[n][0] += 1; agg[n][1] += t
tot = sum(v[1] for v in agg.values())
print('total %.2f ms over %d launches (ncu serialised, cold cache: compare SHARES)' % (tot / 1e3, len(d)))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-30s n=%4d %10.1f us %5.1f%% avg %8.1f us' % (n, c, t, 100 * t / tot, t / c))
if '--seq' in sys.argv:
    N = int(sys.argv[sys.argv.index('--seq') + 1])
    for i, (k, (n, t)) in enumerate(d.items()):
        if i >= N: break
        print('%4s %-24s %8.1f us  %s' % (k, n, t, ' '.join('%s=%s' % (a.split('__')[-1], b) for a, b in extra[k].items())))
