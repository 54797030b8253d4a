"""Throughput of small QPs: warp-per-QP kernel against the CTA-per-QP kernels (FCCQP_NO_WARP=1 in the environment).
usage: python tools/bench_small.py [B]   -> one JSON line per shape"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.synthetic import random_qps
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
from fcc_qp_b200 import _native as nat
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = torch.device("cuda:0")
# "tight": nearly every random QP runs to max_iter (an x-update benchmark); "paper": the settings of fccqp.pdf Table 1
SETS = dict(tight=dict(max_iter=200, rho=1e-3, eps_fcone=1e-7, eps_bound=1e-7), paper=dict(max_iter=15, rho=1e-3, eps_fcone=1e-4, eps_bound=1e-4))
for setting, (n, m, nc, lcs) in [(s_, shp) for s_ in SETS for shp in [(6, 3, 3, 0), (12, 6, 6, 3), (18, 6, 6, 6), (24, 8, 6, 0)]]:
    base = random_qps(np.random.default_rng(n), 4096, n, m, nc, lcs)
    idx = np.arange(B) % 4096
    args = [torch.as_tensor(np.ascontiguousarray(a[idx]), device=dev) for a in
            (base.Q, base.b, base.A_eq, base.b_eq, base.friction_coeffs, base.lb, base.ub)]
    s = FCCQPBatch(n, m, nc, lcs); s.set_options(FCCQPOptionsB(**SETS[setting]))
    best = 1e9
    for _ in range(4):
        s.Solve(*args); torch.cuda.synchronize()
        best = min(best, s.GetSolution().details.device_time)
    it = s.GetSolution().details.n_iter.cpu().numpy()
    print(json.dumps(dict(setting=setting, n=n, m=m, nc=nc, batch=B, ms=best * 1e3, mqps=B / best / 1e6, mean_iterations=float(it.mean()),
                          at_max_iter=float((it >= SETS[setting]['max_iter']).mean()), launch=nat.last_launch_info(),
                          kernel="cta" if os.environ.get("FCCQP_NO_WARP") else "warp")), flush=True)
