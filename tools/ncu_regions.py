"""Aggregate an ncu source-page export (--print-source sass,cuda) by named source regions.
usage: python tools/ncu_regions.py src.csv kernel.cuh   (regions = '// ---' style markers found by regex below)"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2]).read().split("\n")
marks = [(1, "helpers")]
pat = re.compile(r"// -+ (K0: vectors|assemble|unpivoted blocked|explicit inverses|K6: epilogue)|// --- (P1|P2|P3)|// ---- (K3 right|forward|D\^|backward|iterative ref|K4 \+ K5)|// (sigma = trace|H \+= sigma|ADMM initial)|^template <int kThreads")
for i, l in enumerate(src, 1):
    m = pat.search(l)
    if m: marks.append((i, l.strip()[:50]))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]
cs, ci, cw, ce = (hdr.index(n) for n in ('# Samples', 'Instructions Executed', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Excessive'))
stall = [(h, i) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def num(v):
    try: return int(v)
    except Exception: return 0
agg = {}
for r in rows[hi + 1:]:
    try: ln = int(r[0])
    except Exception: continue
    reg = [m for m in marks if m[0] <= ln][-1]
    a = agg.setdefault(reg, [0, 0, 0, 0, {}])
    a[0] += num(r[cs]); a[1] += num(r[ci]); a[2] += num(r[cw]); a[3] += num(r[ce])
    for h, i in stall: a[4][h] = a[4].get(h, 0) + num(r[i])
ts, ti, tw = (sum(a[k] for a in agg.values()) for k in (0, 1, 2))
print(f"total samples {ts} inst {ti} wavefronts {tw}")
for reg in sorted(agg):
    a = agg[reg]
    st = sorted(a[4].items(), key=lambda x: -x[1])[:3]
    print(f"{reg[0]:4d} {reg[1]:52s} samp {100*a[0]/ts:5.1f}% inst {100*a[1]/ti:5.1f}% wf {100*a[2]/max(tw,1):5.1f}% exc {100*a[3]/max(tw,1):5.1f}% " +
          " ".join(f"{h[6:]}:{100*v/max(a[0],1):.0f}%" for h, v in st))
