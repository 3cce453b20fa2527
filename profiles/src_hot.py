"""Hot SASS lines + per-region stall samples of one kernel from an ncu source-page CSV.
usage: ncu -i X.ncu-rep --page source --csv --kernel-id ::regex:NAME:N > src.csv ; python profiles/src_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
isrc = hdr.index('Source'); iall = hdr.index('Warp Stall Sampling (All Samples)'); iex = hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
lines = []
for r in rows[2:]:
    try: lines.append((int(r[iall]), r[isrc], int(r[iex] or 0), r))
    except Exception: pass
tot = sum(l[0] for l in lines)
print('SASS lines', len(lines), 'total samples', tot)
KEYS = ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'STG', 'LDG', 'SYNCS', 'UTCBAR', 'BAR.SYNC', 'SHFL', 'STS', 'LDS')
for st in range(0, len(lines), 250):
    ch = lines[st:st + 250]; s = sum(c[0] for c in ch)
    if s > tot * 0.005:
        ops = sorted({k for c in ch for k in KEYS if k in c[1]})
        print('region %5d-%5d %6d %5.1f%% %s' % (st, st + 250, s, 100 * s / tot, ops))
print('--- hottest')
for n, (s, src, ex, r) in sorted(enumerate(lines), key=lambda t: -t[1][0])[:top]:
    st = {hdr[i][6:]: int(r[i]) for i in stall if r[i] not in ('', '0')}
    print('%5d %6d %5.1f%% ex=%8d %-64s %s' % (n, s, 100 * s / tot, ex, src[:64], sorted(st.items(), key=lambda kv: -kv[1])[:3]))
