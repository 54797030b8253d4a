"""Phase profile with ONE walking-log QP repeated B times through batch stride 0 (every load is an L1/L2 hit):
separates memory latency from instruction cost in the per-QP phases.  FCCQP_NO_SHARED=1 keeps the general /
structure-exploiting path.  usage: python tools/prof_same.py [B] [qp index]"""
import os, sys
os.environ["FCCQP_NO_SHARED"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = sys.argv[3] if len(sys.argv) > 3 else "same"
log = load_walking_log()
dev = torch.device("cuda:0")
if mode == "same":
    qp = log.take(np.array([idx]))
    args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    args = [a.expand(B, *a.shape[1:]) for a in args]
else:   # "copies": the same QP materialised B times (HBM traffic like the real thing, no iterating QPs)
    qp = log.take(np.full(B, idx))
    args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
for r in range(2):
    s.Solve(*args); torch.cuda.synchronize()
    dt = s.GetSolution().details.device_time
    print(f"{mode} qp {idx} B={B} rep {r}: {dt*1e3:.2f} ms -> {B/dt/1e6:.3f} M QP/s iters {int(s.GetSolution().details.n_iter.max())}", flush=True)
