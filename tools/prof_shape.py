"""Device-resident solve of a synthetic shape; used under FCCQP_PROFILE=1 and ncu.
usage: python tools/prof_shape.py [shape] [B] [reps] [cold|warm]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200 import synthetic
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
name = sys.argv[1] if len(sys.argv) > 1 else "humanoid"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "cold"
qp = synthetic.make_batch(synthetic.SHAPES[name], B)
dev = torch.device("cuda:0")
args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
if mode == "warm":
    s.Solve(*args); s.set_warm_start(True)
for r in range(reps):
    s.Solve(*args); torch.cuda.synchronize()
    sol = s.GetSolution()
    dt = sol.details.device_time
    it = sol.details.n_iter.cpu().numpy()
    print(f"{name} {mode} B={B} rep {r}: {dt*1e3:.2f} ms -> {B/dt/1e6:.3f} M QP/s; iterating {np.mean(it > 0):.3f}, mean iters {it.mean():.2f}, at max {np.mean(it >= 100):.3f}", flush=True)
