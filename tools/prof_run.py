"""Device-resident solve of the tiled walking log; used under FCCQP_PROFILE=1 and ncu.
usage: python tools/prof_run.py [B] [reps] [cold|warm]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "cold"
qp = load_walking_log().tile(B)
dev = torch.device("cuda:0")
args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
s.schedule_from_previous = bool(os.environ.get("LPT"))   # FCCQP_SCHEDULE_LPT from the second Solve on
if mode == "warm":
    s.Solve(*args); s.set_warm_start(True)
for r in range(reps):
    s.Solve(*args); torch.cuda.synchronize()
    dt = s.GetSolution().details.device_time
    print(f"{mode} B={B} rep {r}: {dt*1e3:.2f} ms -> {B/dt/1e6:.3f} M QP/s", flush=True)
