"""Hottest SASS instructions of an ncu source-page export (with the CUDA line they belong to).
usage: python tools/ncu_sass_top.py x_src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None; cur = None; out = []
def num(v):
    try: return int(v)
    except Exception: return 0
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "": cur = (r[0], r[1].strip()[:70])
    else: out.append((num(r[6]), r[3].strip(), cur))
tot = sum(o[0] for o in out) or 1
for o in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{100*o[0]/tot:5.2f}% {o[1][:64]:64s} | {o[2][0]}: {o[2][1]}")
