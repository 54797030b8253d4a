"""Static SASS instruction count per source region (nvdisasm -g -c output) for the first kernel in the file.
usage: cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > dis.txt; python tools/sass_size.py dis.txt kernel.cuh"""
import re, sys
dis = open(sys.argv[1]).read().split("\n")
src = open(sys.argv[2]).read().split("\n")
pat = re.compile(r"// -+ (K0: vectors|assemble|unpivoted blocked|explicit inverses|K6: epilogue)|// --- (\(A\)|\(B\)|diagonal tile)|// ---- (K3 right|forward|D\^|backward|iterative ref|K4 \+ K5)|// (sigma = trace|ADMM initial)|^template <int kThreads|^__device__ __forceinline__ void (factor_diag_tile|accumulate_column)|^template <int kOwn")
marks = [(1, "helpers")]
for i, l in enumerate(src, 1):
    m = pat.search(l)
    if m: marks.append((i, l.strip()[:60]))
cur = None; counts = {}; nk = 0; total = 0
for l in dis:
    if l.startswith(".text."):
        nk += 1
        if nk > 1: break
    m = re.search(r'//## File ".*fccqp_kernel.cuh", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l) and cur is not None:
        reg = [mk for mk in marks if mk[0] <= cur][-1]
        counts[reg] = counts.get(reg, 0) + 1; total += 1
print("total SASS instructions:", total, "=", total * 16 // 1024, "KiB")
for reg in sorted(counts): print(f"{reg[0]:4d} {reg[1]:62s} {counts[reg]:6d} {100*counts[reg]/total:5.1f}%")
