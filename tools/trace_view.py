"""Pretty-print a FCCQP_TRACE file: per-warp event timeline (cycles relative to QP start).
usage: python tools/trace_view.py trace.txt [first_event] [n_events]"""
import sys, collections
ev = collections.defaultdict(list)
for l in open(sys.argv[1]):
    w, c, t = l.split(); ev[int(w)].append((int(c), int(t)))
t0 = min(e[0][0] for e in ev.values())
names = {1:"qp", 2:"staged", 3:"asm-issued", 4:"asm-done", 5:"sigma", 6:"sig-red", 7:"rhs0", 10:"col", 11:"acc-done", 12:"P1-done", 13:"bar1", 14:"B-done", 15:"bar2", 16:"hB", 17:"hK",
         20:"xinv", 30:"solve", 31:"f-arr", 32:"f-bar", 33:"b-arr", 34:"b-bar", 35:"solved", 41:"refined", 51:"projected", 60:"epi", 61:"end"}
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
for w in sorted(ev):
    print(f"--- warp {w}: {len(ev[w])} events, total {ev[w][-1][0]-t0} cycles")
    prev = None
    out = []
    for c, t in ev[w][lo:lo+n]:
        d = c - prev if prev is not None else 0
        out.append(f"{names.get(t,t)}@{c-t0}(+{d})")
        prev = c
    print("  ".join(out))
