"""Aggregate an ncu source-page CSV export by (file, function section, line) and by stall reason.
usage: python tools/ncu_phases.py x.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sections = []
cur = None
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == 'File Path':
        cur = dict(file=r[1].split('/')[-1], func='', hdr=None, lines=[])
        sections.append(cur)
    elif r and r[0] == 'Function Name' and cur is not None:
        cur['func'] = r[1]
    elif r and r[0] == 'Line No' and cur is not None:
        cur['hdr'] = r
    elif r and cur is not None and cur['hdr'] is not None:
        cur['lines'].append(r)
    i += 1
def num(v):
    try: return int(v)
    except Exception: return 0
tot = 0
agg = []
stalls = collections.Counter()
for s in sections:
    h = s['hdr']
    if h is None: continue
    if h[0] != 'Line No': continue
    c_samp = h.index('# Samples'); c_inst = h.index('Instructions Executed')
    stall_cols = [(x, k) for k, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    for r in s['lines']:
        try: ln = int(r[0])
        except Exception: continue
        sm = num(r[c_samp])
        tot += sm
        st = sorted(((num(r[k]), x) for x, k in stall_cols), reverse=True)
        for v, x in st: stalls[x] += v
        agg.append((sm, s['file'], ln, r[1].strip()[:90], num(r[c_inst]), st[:2]))
print('sections', [(s['file'], len(s['lines'])) for s in sections])
print('total samples', tot)
print('stalls', [(k[6:], round(100 * v / max(1, sum(stalls.values())), 1)) for k, v in stalls.most_common(12)])
# by file + coarse line ranges
byfile = collections.Counter()
for sm, f, ln, src, inst, st in agg: byfile[f] += sm
print('by file', {k: round(100 * v / tot, 1) for k, v in byfile.items()})
for sm, f, ln, src, inst, st in sorted(agg, reverse=True)[:top]:
    ss = ' '.join(f"{x[6:]}:{100*v/max(sm,1):.0f}%" for v, x in st if v)
    print(f"{f[:14]:14s}:{ln:4d} {100*sm/tot:5.1f}% [{ss}] | {src}")
