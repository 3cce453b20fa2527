"""Attribute ncu warp-stall samples of k_conv_mma to its warp roles (producer / issuer / stager / epilogue) by SASS region."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, iall, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
lines = []
for r in rows[2:]:
    try:
        lines.append((int(r[iall]), r[isrc], r))
    except Exception:
        pass
tot = sum(l[0] for l in lines)
# region boundaries from marker instructions
def first(pat):
    for i, l in enumerate(lines):
        if pat in l[1]:
            return i
    return None
def last(pat):
    idx = None
    for i, l in enumerate(lines):
        if pat in l[1]:
            idx = i
    return idx
marks = {'UBLKCP': (first('UBLKCP'), last('UBLKCP')), 'UTCHMMA': (first('UTCHMMA'), last('UTCHMMA')),
         'F2FP': (first('F2FP'), last('F2FP')), 'LDTM': (first('LDTM'), last('LDTM')), 'STG': (first('STG'), last('STG'))}
print('total samples', tot, 'markers', marks)
# print cumulative samples in windows of 100 SASS lines with the dominant stall
for start in range(0, len(lines), 150):
    chunk = lines[start:start + 150]
    s = sum(c[0] for c in chunk)
    if s < tot * 0.01:
        continue
    st = {}
    for c in chunk:
        for i in stall_cols:
            v = c[2][i]
            if v not in ('', '0'):
                st[hdr[i][6:]] = st.get(hdr[i][6:], 0) + int(v)
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    tags = [k for k, (a, b) in marks.items() if a is not None and not (b < start or a >= start + 150)]
    hot = max(chunk, key=lambda c: c[0])
    print('%5d-%5d  %6d (%4.1f%%)  %-28s top=%s  hottest: %s (%d)' % (start, start + 150, s, 100.0 * s / tot, ','.join(tags), top, hot[1][:50], hot[0]))

# exact sums by role region
def rng(a, b):
    return sum(l[0] for l in lines[a:b])
pa, pb = marks['UBLKCP']
fa, fb = marks['F2FP']
la = marks['LDTM'][0]
sa, sb = marks['STG']
ua, ub = marks['UTCHMMA']
print('producer  region [0,%d): %d' % (fa - 250, rng(0, fa - 250)))
print('stager    region [%d,%d): %d' % (fa - 250, la - 250, rng(fa - 250, la - 250)))
print('epilogue  region [%d,%d): %d' % (la - 250, sb + 50, rng(la - 250, sb + 50)))
print('issuer    region [%d,%d): %d' % (sb + 50, ub + 10, rng(sb + 50, ub + 10)))
print('tail      region [%d,end): %d' % (ub + 10, rng(ub + 10, len(lines))))
# epilogue work (excluding the wait loop = lines with PHASECHK / ISETP 0x4000001 neighbours)
work = sum(l[0] for l in lines[la:sb + 50])
print('epilogue non-wait samples (LDTM..last STG):', work)
work_s = sum(l[0] for l in lines[fa - 40:fb + 120])
print('stager non-wait samples (around F2FP..):', work_s)
