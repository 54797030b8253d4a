"""profiles/dram_traffic.json (read by bench.py for roofline.traffic) from an ncu metrics-only capture AT the bench batch:
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:fccqp_struct_kernel -s 1 -c 1 --csv
    --log-file gpurun_out/r2_dram65536.csv python tools/prof_run.py 65536 2
usage: python tools/make_dram_traffic.py gpurun_out/r2_dram65536.csv <commit> [batch]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, commit = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
rows = [r for r in csv.reader(open(path)) if len(r) > 14]
m, kern = {}, None
for r in rows:
    if r[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        kern = r[4]
        m[r[12]] = float(r[14].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(r[13], 1)
rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
out = {"kernel": kern, "commit": commit, "batch_profiled": B, "dram_read_bytes": rd, "dram_write_bytes": wr,
       "bytes_per_qp": (rd + wr) / B, "bytes_per_launch_at_batch_65536": (rd + wr) * 65536 / B,
       "note": "ncu metrics-only capture of one launch at the bench batch itself (walking log tiled to 2^16, cold)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
print(out)
