"""Throughput on long-running QPs only: the walking-log QPs that run to max_iter, tiled.
usage: python tools/heavy_run.py [B]   (honours FCCQP_FULL_INVERSE_AT, FCCQP_CTAS_PER_SM)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
gold = np.load(os.path.join(ROOT, "tests", "golden", "walking_cold.npz"))
idx = np.nonzero(gold["n_iter"] == 100)[0]
qp = load_walking_log().take(idx[np.arange(B) % len(idx)])
dev = torch.device("cuda:0")
args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
for r in range(3):
    s.Solve(*args); torch.cuda.synchronize()
    dt = s.GetSolution().details.device_time
print(f"heavy B={B}: {dt*1e3:.2f} ms -> {B/dt/1e3:.1f} k QP/s, {dt*1.965e9*min(592, B)/B/100:.0f} cycles per iteration per CTA-slot (incl. factorizations)")
