"""Small cold + warm batched solves for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).
usage: compute-sanitizer --tool racecheck python tools/sanitize_run.py [shape]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200 import synthetic as syn
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
shape = sys.argv[1] if len(sys.argv) > 1 else "walking"
qp = load_walking_log().take(np.arange(0, 2019, 48)) if shape == "walking" else syn.make_batch(syn.SHAPES[shape], 24)
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
it = s.GetSolution().details.n_iter
s.set_warm_start(True)
s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
print(shape, "QPs", qp.batch, "cold iterations", np.unique(it, return_counts=True), "warm", np.unique(s.GetSolution().details.n_iter, return_counts=True))
