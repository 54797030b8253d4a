"""Developer aid: why the polish accepts / rejects (GPU against the numpy restatement, per criterion) on the walking log."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, oracle
from oracle import polish as pol
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
from fcc_qp_b200.logdata import QPBatch, load_walking_log
qp = load_walking_log()
B, n, m = qp.batch, qp.n, qp.m
dev = torch.device("cuda:0")
args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
s = FCCQPBatch(n, m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(max_iter=100, rho=5e-5, eps_fcone=1e-6, eps_bound=1e-6))
s.Solve(*args); a = s.GetSolution()
z0, bv0, fv0 = a.z.cpu().numpy(), a.details.bounds_viol.cpu().numpy(), a.details.friction_cone_viol.cpu().numpy()
st0 = a.details.solve_status.cpu().numpy()
x, mux, muc = (t.clone().cpu().numpy() for t in s.GetState())
p = s.Polish(); torch.cuda.synchronize()
flag = p.details.polished.cpu().numpy()
inner = s._polish_buf["inner"].GetSolution()
yg, sg = inner.z.cpu().numpy(), inner.details.solve_status.cpu().numpy()
print("gpu accepted", flag.mean(), "gpu inner status", np.unique(sg, return_counts=True))
Qp, bp, Ap, beqp, rot = pol.prepare(qp, x, mux, muc)
print("kinds", np.unique(rot[:, :, 0], return_counts=True), "bound-fixed per QP max", 0)
o = oracle.Oracle("port")
fr = np.broadcast_to(qp.friction_coeffs, (B, 4)); lb = np.broadcast_to(qp.lb, (B, n)); ub = np.broadcast_to(qp.ub, (B, n))
for name, y, ys in (("gpu-y", yg, sg),):
    reasons = dict(status=0, eq=0, bound=0, cone=0, obj=0)
    worst = dict(eq=0.0, cone=0.0, obj=-1e9)
    for q in range(B):
        xp = y[q].copy()
        for c in range(4):
            o_ = qp.lambda_c_start + 3 * c
            if rot[q, c, 0] == 1: xp[o_:o_ + 3] = 0
            elif rot[q, c, 0] == 2:
                d = rot[q, c, 1:]; h = np.hypot(d[0], d[1]); t1 = np.array((-d[1] / h, d[0] / h, 0.0)) if h > 0 else np.zeros(3)
                xp[o_:o_ + 3] = y[q, o_] * d + y[q, o_ + 1] * t1
        if ys[q] != 0: reasons["status"] += 1
        terms = qp.A_eq[q] * xp[None, :]; mag = np.abs(qp.b_eq[q]) + np.abs(terms).sum(1)
        e = (np.abs(terms.sum(1) - qp.b_eq[q]) / mag).max(); worst["eq"] = max(worst["eq"], e)
        if e > 1e-7: reasons["eq"] += 1
        if (xp < lb[q] - 1e-6).any() or (xp > ub[q] + 1e-6).any(): reasons["bound"] += 1
        cv = max(np.hypot(xp[38 + 3 * c], xp[39 + 3 * c]) - fr[q, c] * xp[40 + 3 * c] for c in range(4)); worst["cone"] = max(worst["cone"], cv)
        if cv > 1e-6: reasons["cone"] += 1
        obj = lambda v: 0.5 * v @ qp.Q[q] @ v + qp.b[q] @ v
        fa = obj(z0[q]); g = (obj(xp) - fa) / max(1.0, abs(fa)); worst["obj"] = max(worst["obj"], g)
        if g > 1e-3: reasons["obj"] += 1
    print(name, "rejections by criterion", reasons, "worst", worst)
print("ADMM: status1", (st0 == 1).sum(), "fv max", fv0.max())
