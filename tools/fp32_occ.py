import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
from fcc_qp_b200.synthetic import random_qps
from fcc_qp_b200 import _native as nat
dev = torch.device("cuda:0"); B = 1 << 18
for (n, m, nc, lcs) in ((6, 3, 3, 3), (12, 6, 6, 3), (24, 8, 6, 0)):
    base = random_qps(np.random.default_rng(n), 4096, n, m, nc, lcs)
    args = [torch.as_tensor(a, device=dev, dtype=torch.float32).repeat((B // 4096,) + (1,) * (a.ndim - 1)) for a in (base.Q, base.b, base.A_eq, base.b_eq, base.friction_coeffs, base.lb, base.ub)]
    for tag, opts in (("paper", dict(max_iter=15, rho=1e-3, eps_fcone=1e-4, eps_bound=1e-4)), ("long", dict(max_iter=200, rho=1e-3, eps_fcone=1e-4, eps_bound=1e-4))):
        s = FCCQPBatch(n, m, nc, lcs, precision="fp32"); s.set_options(FCCQPOptionsB(**opts))
        best = 1e9
        for _ in range(4):
            s.Solve(*args); torch.cuda.synchronize(); best = min(best, s.GetSolution().details.device_time)
        print(os.environ.get("FCCQP_WARP_F32_CTAS", "6"), n, m, tag, f"{B / best / 1e6:.2f} M QP/s", nat.last_launch_info(), flush=True)
