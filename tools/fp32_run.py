"""FCCQP_PRECISION_FP32 (float32 data + FP32 arithmetic, warp kernel) against FP64 on random small QPs: error on z and on
the objective, status / iteration-count agreement, and the rate of both modes.  usage: python tools/fp32_run.py [B]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
from fcc_qp_b200.synthetic import random_qps
from fcc_qp_b200 import _native as nat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = torch.device("cuda:0")


def kkt_cond(qp, i, rho=1e-3):
    n, m = qp.n, qp.m
    K = np.zeros((n + m, n + m)); K[:n, :n] = qp.Q[i] + rho * np.eye(n); K[n:, :n] = qp.A_eq[i]; K[:n, n:] = qp.A_eq[i].T
    return np.linalg.cond(K)


# error against conditioning: the rows of A_eq scaled by 10^u, u uniform in [-s, s] (the solution does not change, the KKT
# matrix gets worse), FP32 against FP64 at 200 iterations
for (n, m, nc, lcs) in ((12, 6, 6, 3), (24, 8, 6, 0)):
    from fcc_qp_b200.synthetic import scale_constraint_rows
    for spread in (0.5, 1.0, 1.5, 2.0):
        qp = scale_constraint_rows(random_qps(np.random.default_rng(100 + n), 2048, n, m, nc, lcs), np.random.default_rng(7), spread)
        cond = np.array([kkt_cond(qp, i) for i in range(qp.batch)])
        out = {}
        for prec in ("fp64", "fp32"):
            dt = torch.float64 if prec == "fp64" else torch.float32
            args = [torch.as_tensor(a, device=dev, dtype=dt) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
            s = FCCQPBatch(n, m, nc, lcs, precision=prec); s.set_options(FCCQPOptionsB(max_iter=200, rho=1e-3, eps_fcone=1e-4, eps_bound=1e-4))
            s.Solve(*args); torch.cuda.synchronize()
            sol = s.GetSolution()
            out[prec] = (sol.z.cpu().numpy(), sol.details.solve_status.cpu().numpy())
        err = np.abs(out["fp64"][0] - out["fp32"][0]).max(axis=1) / np.maximum(1.0, np.abs(out["fp64"][0]).max(axis=1))
        row = {"n": n, "m": m, "row_scale_spread": spread, "status_equal": float((out["fp64"][1] == out["fp32"][1]).mean()),
               "status2_fp32": int((out["fp32"][1] == 2).sum())}
        for lo, hi in ((0, 1e2), (1e2, 1e3), (1e3, 1e4), (1e4, 1e5), (1e5, 1e9)):
            sel = (cond >= lo) & (cond < hi)
            if sel.any():
                row[f"cond<{hi:g}"] = {"count": int(sel.sum()), "z_err_p50": float(np.median(err[sel])), "z_err_max": float(err[sel].max())}
        print(json.dumps(row), flush=True)

for (n, m, nc, lcs) in ((6, 3, 3, 3), (12, 6, 6, 3), (18, 6, 6, 0), (24, 8, 6, 0)):
    base = random_qps(np.random.default_rng(n), 4096, n, m, nc, lcs)
    # condition number of the rho-KKT matrix of a sample
    conds = []
    for i in range(0, 4096, 64):
        K = np.zeros((n + m, n + m)); K[:n, :n] = base.Q[i] + 1e-3 * np.eye(n); K[n:, :n] = base.A_eq[i]; K[:n, n:] = base.A_eq[i].T
        conds.append(np.linalg.cond(K))
    for tag, opts in (("paper", dict(max_iter=15, rho=1e-3, eps_fcone=1e-4, eps_bound=1e-4)),
                      ("converged", dict(max_iter=500, rho=1e-3, eps_fcone=1e-4, eps_bound=1e-4))):
        res = {}
        for prec in ("fp64", "fp32"):
            dt = torch.float64 if prec == "fp64" else torch.float32
            reps = B // base.batch
            args = [torch.as_tensor(a, device=dev, dtype=dt) for a in (base.Q, base.b, base.A_eq, base.b_eq, base.friction_coeffs, base.lb, base.ub)]
            args = [a.repeat((reps,) + (1,) * (a.dim() - 1)) for a in args]
            s = FCCQPBatch(n, m, nc, lcs, precision=prec); s.set_options(FCCQPOptionsB(**opts))
            best = 1e9
            for _ in range(3):
                s.Solve(*args); torch.cuda.synchronize()
                best = min(best, s.GetSolution().details.device_time)
            sol = s.GetSolution()
            res[prec] = dict(z=sol.z[:4096].cpu().numpy(), it=sol.details.n_iter[:4096].cpu().numpy(),
                             st=sol.details.solve_status[:4096].cpu().numpy(), qps=B / best, launch=nat.last_launch_info())
        z64, z32 = res["fp64"]["z"], res["fp32"]["z"]
        err = np.abs(z64 - z32).max(axis=1) / np.maximum(1.0, np.abs(z64).max(axis=1))
        obj = lambda z: 0.5 * np.einsum("bi,bij,bj->b", z, base.Q, z) + np.einsum("bi,bi->b", base.b, z)
        oerr = np.abs(obj(z64) - obj(z32)) / np.maximum(1.0, np.abs(obj(z64)))
        eqv = np.abs(np.einsum("bij,bj->bi", base.A_eq, z32) - base.b_eq).max(axis=1) if m else np.zeros(1)
        print(json.dumps({"n": n, "m": m, "nc": nc, "settings": tag, "batch": B, "cond_p50": float(np.median(conds)), "cond_max": float(np.max(conds)),
                          "z_err_p50": float(np.median(err)), "z_err_p99": float(np.quantile(err, 0.99)), "z_err_max": float(err.max()),
                          "obj_err_max": float(oerr.max()), "eq_residual_max_fp32": float(eqv.max()),
                          "status_equal": float((res["fp64"]["st"] == res["fp32"]["st"]).mean()),
                          "iters_equal": float((res["fp64"]["it"] == res["fp32"]["it"]).mean()),
                          "mean_iters_fp64": float(res["fp64"]["it"].mean()), "mean_iters_fp32": float(res["fp32"]["it"].mean()),
                          "qps_fp64": res["fp64"]["qps"], "qps_fp32": res["fp32"]["qps"], "speedup": res["fp32"]["qps"] / res["fp64"]["qps"],
                          "smem_fp64": res["fp64"]["launch"]["smem_bytes"], "smem_fp32": res["fp32"]["launch"]["smem_bytes"]}), flush=True)
