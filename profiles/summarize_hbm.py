"""Per-kernel DRAM traffic / throughput from an ncu metrics CSV written by profiles/ncu_hbm.sh.
usage: python profiles/summarize_hbm.py profiles/hbm_TAG.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
hdr = rows[0]; ik = hdr.index('Kernel Name'); im = hdr.index('Metric Name'); iv = hdr.index('Metric Value'); iid = hdr.index('ID'); iu = hdr.index('Metric Unit')
d = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv: continue
    e = d.setdefault(r[iid], {'k': r[ik].split('(')[0].split('::')[-1]})
    try: v = float(r[iv].replace(',', ''))
    except ValueError: continue
    if r[im].startswith('dram__bytes'): v *= {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(r[iu], 1)
    if r[im] == 'gpu__time_duration.sum': v *= {'us': 1, 'usecond': 1, 'ns': 1e-3, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}.get(r[iu], 1)
    e[r[im]] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
for e in d.values():
    if 'gpu__time_duration.sum' not in e: continue
    a = agg[e['k']]; a[0] += 1; a[1] += e['gpu__time_duration.sum']; a[2] += e.get('dram__bytes_read.sum', 0); a[3] += e.get('dram__bytes_write.sum', 0)
    a[4] += e.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', 0); a[5] += e.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0)
print('%d launches (ncu serialised, cold cache)' % len(d))
print('| kernel | launches | avg time | DRAM read + write per launch | DRAM GB/s | DRAM %% of peak | SM %% of peak |')
print('|---|---|---|---|---|---|---|')
for k, (n, t, r, w, p, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('| `%s` | %d | %.1f µs | %.1f + %.1f MB | %.0f | %.0f %% | %.0f %% |' % (k, n, t / n, r / n / 1e6, w / n / 1e6, (r + w) / t / 1e3, p / n, sm / n))
