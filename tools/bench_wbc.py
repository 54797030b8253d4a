"""End-to-end throughput when the batch is assembled on the GPU: robot
quantities in page-locked HOST memory -> H2D -> fccqp_wbc_assemble -> batched cold solve -> z, n_iter
back to the host, against the same QPs shipped as dense Q/A through the reference-shaped host path.
usage: python tools/bench_wbc.py [shape] [batch] [steps]   -> one JSON line"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200 import synthetic as syn, wbc
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
name = sys.argv[1] if len(sys.argv) > 1 else "cassie_like"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
shp = syn.SHAPES[name]
dev = torch.device("cuda:0")
gen = min(B, 4096)
t0 = syn.make_terms(shp, gen)
idx = np.arange(B) % gen
terms = syn.WBCTerms(shp, *(np.ascontiguousarray(getattr(t0, k)[idx]) for k in
                            ("M", "Jh", "Jc", "Jy", "W", "ydd_cmd", "bias", "gamma_h", "gamma_c", "friction_coeffs")))
s = FCCQPBatch(shp.n, shp.m, shp.nc, shp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
s.time_kernel = False
keys = ("M", "Jh", "Jc", "Jy", "W", "ydd_cmd", "bias", "gamma_h", "gamma_c", "friction_coeffs")
pinned = {k: torch.from_numpy(getattr(terms, k)).pin_memory() for k in keys}   # caller-owned page-locked terms
def step():
    t = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}           # H2D of the robot quantities
    Q, b, A, beq, mu, lb, ub = wbc.assemble(t, shape=shp, device=dev)
    s.Solve(Q, b, A, beq, mu, lb, ub)
    sol = s.GetSolution()
    z = sol.z.cpu(); it = sol.details.n_iter.cpu()              # D2H of the results (synchronises)
    return float(z[:, 0].sum()) + float(it.sum())
for _ in range(2): step()
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(steps): step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t1) / steps
# assembly kernel alone
t = wbc.to_device(terms, dev); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = wbc.assemble(t, shape=shp, device=dev); e1.record(); torch.cuda.synchronize()
asm_ms = e0.elapsed_time(e1)
out_bytes = sum(x.numel() * 8 for x in out[:4])
# dense host path on the same QPs
qp = syn.assemble_numpy(terms)
host = [torch.from_numpy(a).pin_memory().numpy() for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
hs = FCCQPBatch(shp.n, shp.m, shp.nc, shp.lambda_c_start); hs.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6)); hs.zero_copy_outputs = True
for _ in range(2): hs.Solve(*host)
t1 = time.perf_counter()
for _ in range(steps): hs.Solve(*host); hs.GetSolution()
dt_dense = (time.perf_counter() - t1) / steps
print(json.dumps({"shape": name, "n": shp.n, "m": shp.m, "batch": B,
                  "assembled_e2e": {"qps_per_s": B / dt, "ms_per_step": 1e3 * dt, "h2d_bytes_per_step": terms.nbytes()},
                  "dense_e2e": {"qps_per_s": B / dt_dense, "ms_per_step": 1e3 * dt_dense,
                                "h2d_bytes_per_step": int(sum(a.nbytes for a in host))},
                  "assemble_kernel": {"ms": asm_ms, "GB_per_s_written": out_bytes / asm_ms / 1e6}}))
