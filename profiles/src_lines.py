"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line: instructions executed and
warp-stall samples per (file, line) (top N), plus the stall-reason totals.  The dump has one section per source file
("File Path" rows).  usage: python profiles/src_lines.py dump.csv [N]   (source text is read from the repo's csrc/)"""
import csv
import os
import sys
from collections import defaultdict

ROOT = os.environ.get("VF_SRC_DIR") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "visual_foresight_b200", "csrc")
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 30
ex, sm, st = defaultdict(float), defaultdict(float), defaultdict(float)
cur, hdr, iex, ismp, stall = "?", None, None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1])
        continue
    if r[0] == "Line No":
        hdr = r
        iex, ismp = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
        continue
    if hdr is None or len(r) <= iex:
        continue
    try:
        ln = int(r[0])
        e, s_ = float(r[iex] or 0), float(r[ismp] or 0)
    except ValueError:
        continue
    ex[(cur, ln)] += e
    sm[(cur, ln)] += s_
    for i in stall:
        try:
            st[hdr[i]] += float(r[i] or 0)
        except (ValueError, IndexError):
            pass
text = {}
def line(f, ln):
    if f not in text:
        p = os.path.join(ROOT, f)
        text[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return text[f][ln - 1].strip()[:110] if 0 < ln <= len(text[f]) else ""
tot, tots = sum(ex.values()), sum(sm.values())
print("total warp instructions %.2f M, stall samples %d  (line numbers: the source as of the capture)" % (tot / 1e6, tots))
print("stall reasons:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(sum(st.values()), 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
for (f, ln), s_ in sorted(sm.items(), key=lambda kv: -kv[1])[:top]:
    print("%-16s %5d  %5.1f%% samples  %5.1f%% instr | %s" % (f, ln, 100 * s_ / max(tots, 1), 100 * ex[(f, ln)] / max(tot, 1), line(f, ln)))
