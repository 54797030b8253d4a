"""Summarise an ncu source-page export per CUDA source line (all source files of the kernel).
usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > x.csv; python tools/ncu_lines.py x.csv [N]"""
import csv, os, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
def num(v):
    try: return int(v)
    except Exception: return 0
lines, stall_tot = [], {}
fname, hdr = "?", None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r
        col = lambda name: [i for i, h in enumerate(hdr) if h == name][0]
        c_samp, c_inst = col('# Samples'), col('Instructions Executed')
        c_wf, c_exc = col('L1 Wavefronts Shared'), col('L1 Wavefronts Shared Excessive')
        stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None: continue
    try: ln = int(r[0])
    except Exception: continue
    st_all = [(num(r[i]), h) for h, i in stall_cols]
    for v, h in st_all: stall_tot[h] = stall_tot.get(h, 0) + v
    lines.append((ln, r[1], num(r[c_samp]), num(r[c_inst]), num(r[c_wf]), num(r[c_exc]), sorted(st_all, reverse=True)[:2], fname))
ts, ti, tw = (sum(l[k] for l in lines) for k in (2, 3, 4))
print('total samples', ts, 'warp-inst', ti, 'smem wavefronts', tw, 'excessive', sum(l[5] for l in lines))
sa = sum(stall_tot.values()) or 1
print('stall reasons (all samples): ' + ', '.join(f"{h[6:]} {100*v/sa:.1f}%" for h, v in sorted(stall_tot.items(), key=lambda x: -x[1])[:8]))
byfile = {}
for l in lines: byfile[l[7]] = byfile.get(l[7], 0) + l[2]
print('samples by file: ' + ', '.join(f"{f} {100*v/max(ts,1):.1f}%" for f, v in sorted(byfile.items(), key=lambda x: -x[1])[:5]))
for l in sorted(lines, key=lambda x: -x[2])[:top]:
    st = ' '.join(f"{h[6:]}:{100*v/max(l[2],1):.0f}%" for v, h in l[6] if v)
    print(f"{l[7][6:-4]:>7s}:{l[0]:4d} samp {100*l[2]/max(ts,1):5.1f}% inst {100*l[3]/max(ti,1):5.1f}% wf {100*l[4]/max(tw,1):5.1f}% [{st}] | {l[1].strip()[:96]}")
