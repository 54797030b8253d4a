"""Developer GPU check: parity vs goldens + quick timing.  Run under gpurun:
    python tools/gpu_check.py [--quick]
Writes a report to gpurun_out/gpu_check.txt as well as stdout.
"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "gpu_check.txt"), "a")
def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); out.write(s + "\n"); out.flush()

from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
from fcc_qp_b200 import _native as nat

quick = "--quick" in sys.argv
qp = load_walking_log()
G = os.path.join(ROOT, "tests", "golden")
opts = FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6)

def compare(tag, sol, gold, qp):
    z = np.asarray(sol.z.cpu() if hasattr(sol.z, "cpu") else sol.z)
    gi = lambda a: np.asarray(a.cpu() if hasattr(a, "cpu") else a)
    err = np.abs(z - gold["z"]).max(1) / np.maximum(1, np.abs(gold["z"]).max(1))
    nm = (gi(sol.details.n_iter) != gold["n_iter"]).sum()
    sm = (gi(sol.details.solve_status) != gold["status"]).sum()
    o1, o2 = qp.objective(z), qp.objective(gold["z"])
    oe = (np.abs(o1 - o2) / np.maximum(1, np.abs(o2))).max()
    P(f"[{tag}] z rel err max {err.max():.3e} (argmax {err.argmax()}), n_iter mismatches {nm}, status mismatches {sm}, obj rel {oe:.3e}")
    P(f"   nonfinite z rows: {(~np.isfinite(z)).any(1).sum()}, res_fcone maxdiff {np.abs(gi(sol.details.eps_friction_cone)-gold['res_fcone']).max():.3e}, "
      f"fcone_viol maxdiff {np.abs(gi(sol.details.friction_cone_viol)-gold['fcone_viol']).max():.3e}, bounds_viol maxdiff {np.abs(gi(sol.details.bounds_viol)-gold['bounds_viol']).max():.3e}")
    if nm:
        idx = np.nonzero(gi(sol.details.n_iter) != gold["n_iter"])[0][:10]
        P("   n_iter mism idx", idx, gi(sol.details.n_iter)[idx], gold["n_iter"][idx])
    bad = np.nonzero(err > 1e-6)[0]
    if len(bad):
        P("   bad idx (first 10):", bad[:10], err[bad[:10]])
    return err.max(), nm

P("=== device count", nat.lib().fccqp_device_count())
# 1. host path, cold, whole log as one batch
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start)
s.set_options(opts)
gold = np.load(os.path.join(G, "walking_cold.npz"))
for B in ([4, 2019] if not quick else [4]):
    sub = qp.take(np.arange(B))
    t0 = time.time()
    s.Solve(sub.Q, sub.b, sub.A_eq, sub.b_eq, sub.friction_coeffs, sub.lb, sub.ub)
    sol = s.GetSolution()
    P(f"host cold B={B}: wall {time.time()-t0:.4f}s, launch info {nat.last_launch_info()}")
    compare(f"host cold B={B}", sol, {k: gold[k][:B] for k in gold.files if k != 'opts'}, sub)

# 2. device path via torch
import torch
dev = torch.device("cuda:0")
tq = lambda a: torch.as_tensor(a, device=dev)
s2 = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start)
s2.set_options(opts)
args = [tq(a) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
s2.Solve(*args); torch.cuda.synchronize()
sol = s2.GetSolution()
P(f"device cold B=2019: kernel {sol.details.device_time*1e3:.3f} ms -> {2019/sol.details.device_time:.0f} QP/s")
compare("device cold B=2019", sol, gold, qp)

# 3. warm sequential replay through batch-of-1 (fcc_qp_test.py loop) on first K QPs
K = 300 if not quick else 20
goldw = np.load(os.path.join(G, "walking_warm.npz"))
s3 = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s3.set_options(opts)
zs, its = [], []
t0 = time.time()
for i in range(K):
    s3.set_warm_start(i > 0)
    s3.Solve(qp.Q[i:i+1], qp.b[i:i+1], qp.A_eq[i:i+1], qp.b_eq[i:i+1], qp.friction_coeffs[i], qp.lb[i], qp.ub[i])
    r = s3.GetSolution(); zs.append(r.z[0]); its.append(int(r.details.n_iter[0]))
P(f"warm sequential K={K}: {(time.time()-t0)/K*1e6:.1f} us/QP through python batch-of-1")
zs = np.array(zs); its = np.array(its)
err = np.abs(zs - goldw["z"][:K]).max(1) / np.maximum(1, np.abs(goldw["z"][:K]).max(1))
P(f"[warm seq] z rel err max {err.max():.3e}, n_iter mismatches {(its != goldw['n_iter'][:K]).sum()}")

# 4. throughput at 2^16 (tiled log), device resident
if not quick:
    big = qp.tile(65536)
    args = [tq(a) for a in (big.Q, big.b, big.A_eq, big.b_eq, big.friction_coeffs, big.lb, big.ub)]
    s4 = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s4.set_options(opts)
    for rep in range(4):
        s4.Solve(*args); torch.cuda.synchronize()
        dt = s4.GetSolution().details.device_time
        P(f"device cold B=65536 rep {rep}: {dt*1e3:.2f} ms -> {65536/dt/1e6:.3f} M QP/s  {nat.last_launch_info()}")
    sol = s4.GetSolution()
    z = sol.z.cpu().numpy(); idx = np.arange(65536) % 2019
    err = np.abs(z - gold["z"][idx]).max(1) / np.maximum(1, np.abs(gold["z"][idx]).max(1))
    P(f"[tiled 65536] z rel err max {err.max():.3e}, n_iter mism {(sol.details.n_iter.cpu().numpy() != gold['n_iter'][idx]).sum()}")
    # warm at 2^16: second solve reusing state
    s4.set_warm_start(True)
    for rep in range(3):
        s4.Solve(*args); torch.cuda.synchronize()
        dt = s4.GetSolution().details.device_time
        P(f"device warm B=65536 rep {rep}: {dt*1e3:.2f} ms -> {65536/dt/1e6:.3f} M QP/s")
P("done")
