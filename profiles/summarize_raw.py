"""Per-kernel summary of an `ncu --page raw --csv` dump (one row per launch, one column per metric): launches, average
duration, DRAM read + write per launch, achieved DRAM GB/s and its fraction of the measured HBM peak, SM throughput.
usage: python profiles/summarize_raw.py raw.csv [peak_GBps=6545]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6545.0
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, k, scale=None):
    try:
        v = float(r[ix[k]].replace(",", ""))
    except (ValueError, KeyError, IndexError):
        return 0.0
    u = units[ix[k]]
    if scale == "bytes":
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    if scale == "us":
        v *= {"us": 1, "usecond": 1, "ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    return v


agg = collections.OrderedDict()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    m = re.search(r"k_\w+(<[^>]*>)?", r[ix["Kernel Name"]])
    k = m.group(0) if m else r[ix["Kernel Name"]]
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += val(r, "gpu__time_duration.sum", "us")
    a[2] += val(r, "dram__bytes_read.sum", "bytes")
    a[3] += val(r, "dram__bytes_write.sum", "bytes")
    a[4] += val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    a[5] += val(r, "launch__registers_per_thread")
print("| kernel | launches | avg time | DRAM read + write per launch | DRAM GB/s | of %.0f GB/s | SM %% of peak | regs |" % peak)
print("|---|---|---|---|---|---|---|---|")
for k, (n, t, rd, wr, sm, rg) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (rd + wr) / t / 1e3 if t else 0
    print("| `%s` | %d | %.1f us | %.1f + %.1f MB | %.0f | %.0f %% | %.0f %% | %d |" % (k, n, t / n, rd / n / 1e6, wr / n / 1e6, gbs, 100 * gbs / peak, sm / n, rg / n))
