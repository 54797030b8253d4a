"""Summarise an ncu source-page export per CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > x.csv; python tools/ncu_lines.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]
def col(name, nth=0):
    return [i for i, h in enumerate(hdr) if h == name][nth]
c_samp, c_inst = col('# Samples'), col('Instructions Executed')
c_wf, c_exc = col('L1 Wavefronts Shared'), col('L1 Wavefronts Shared Excessive')
stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def num(v):
    try: return int(v)
    except Exception: return 0
lines = []
for r in rows[hi + 1:]:
    if r and r[0] not in ('', None):
        try: ln = int(r[0])
        except Exception: continue
        st = sorted(((num(r[i]), h) for h, i in stall_cols), reverse=True)[:2]
        lines.append((ln, r[1], num(r[c_samp]), num(r[c_inst]), num(r[c_wf]), num(r[c_exc]), st))
ts, ti, tw = (sum(l[k] for l in lines) for k in (2, 3, 4))
print('total samples', ts, 'warp-inst', ti, 'smem wavefronts', tw, 'excessive', sum(l[5] for l in lines))
for l in sorted(lines, key=lambda x: -x[2])[:top]:
    st = ' '.join(f"{h[6:]}:{100*v/max(l[2],1):.0f}%" for v, h in l[6] if v)
    print(f"{l[0]:4d} samp {100*l[2]/ts:5.1f}% inst {100*l[3]/ti:5.1f}% wf {100*l[4]/max(tw,1):5.1f}% exc {100*l[5]/max(tw,1):5.1f}% [{st}] | {l[1].strip()[:100]}")
