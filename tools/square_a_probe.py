import sys, os, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import oracle
from test_gpu_random_shapes import random_qps, rel
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
OPTS = dict(max_iter=200, rho=1e-3, eps_fcone=1e-7, eps_bound=1e-7)
n, m, nc, lcs = 16, 16, 6, 10
qp = random_qps(np.random.default_rng(77 * n + m), 1024, n, m, nc, lcs)
ref = oracle.Oracle("port").solve_batch(qp, warm_mode=0, nthreads=8, **OPTS)
s = FCCQPBatch(n, m, nc, lcs); s.set_options(FCCQPOptionsB(**OPTS))
s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
z = s.GetSolution().z
e = np.abs(z - ref["z"]).max(1) / np.maximum(1, np.abs(ref["z"]).max(1))
i = int(e.argmax())
print(os.environ.get("FCCQP_NO_WARP"), "max err", e.max(), "at", i, "cond(A)", np.linalg.cond(qp.A_eq[i]), "n>1e-7:", (e > 1e-7).sum())
