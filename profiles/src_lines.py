"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line: instructions executed and
warp-stall samples per line (top N), plus the stall-reason totals.  usage: python profiles/src_lines.py dump.csv file.cu [N]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2]).read().splitlines() if len(sys.argv) > 2 else []
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
iex, ismp = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and "(Not" not in h]
ex, sm = defaultdict(float), defaultdict(float)
st = defaultdict(float)
for r in rows[hi + 1:]:
    if len(r) <= iex:
        continue
    try:
        ln = int(r[0])
        e, s = float(r[iex] or 0), float(r[ismp] or 0)
    except ValueError:
        continue
    ex[ln] += e
    sm[ln] += s
    for i in stall:
        try:
            st[hdr[i]] += float(r[i] or 0)
        except (ValueError, IndexError):
            pass
tot, tots = sum(ex.values()), sum(sm.values())
print("total warp instructions %.2f M, stall samples %d" % (tot / 1e6, tots))
print("stall reasons:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(sum(st.values()), 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
for ln, s in sorted(sm.items(), key=lambda kv: -kv[1])[:top]:
    print("%5d  %5.1f%% samples  %5.1f%% instr | %s" % (ln, 100 * s / max(tots, 1), 100 * ex[ln] / max(tot, 1), src[ln - 1].strip()[:120] if ln - 1 < len(src) else ""))
