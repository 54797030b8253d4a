"""Top SASS instructions by stall samples from an ncu source-page export (--print-source sass,cuda).
usage: python tools/ncu_sass_top.py src.csv [N] [lo_line hi_line]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 10**9)
hdr = rows[2]
cs, ci = hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [(h, i) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def num(v):
    try: return int(v)
    except Exception: return 0
# second section: per-SASS rows (line number column empty); remember the last seen CUDA line
out = []
tot = 0
sass_rows = [r for r in rows[3:] if len(r) > ci and r[2] not in ('-', 'Address') and r[2].startswith('0x')]
for r in sass_rows:
    tot += num(r[cs])
for k, r in enumerate(sass_rows):
    st = sorted(((num(r[i]), h) for h, i in stall), reverse=True)[:3]
    out.append((num(r[cs]), k, r[3][:64], num(r[ci]), st))
print("sass rows", len(sass_rows), "total samples", tot)
for s, k, ins, ex, st in sorted(out, reverse=True)[:N]:
    print(f"#{k:5d} {100*s/max(tot,1):5.2f}% exec {ex:10d} {ins:64s} " + " ".join(f"{h[6:]}:{v}" for v, h in st if v))
