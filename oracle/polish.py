"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the opt-in solution polish (include/fccqp.h, fccqp_polish_prepare /
fccqp_polish_finish; kernels in fcc_qp_b200/csrc/fccqp_polish.cuh).

The reference has NO polish step (src/fcc_qp.cpp returns the ADMM iterate), so this oracle is builder-authored and no parity
with the reference is claimed for the feature; what IS the reference's is the inner solve: the polished problem is an
equality-constrained QP handed to the compiled reference / its C restatement (Oracle.solve_batch with nc = 0 and infinite
bounds, i.e. the KKT pre-solve of src/fcc_qp.cpp:159-178).
"""
from __future__ import annotations

import numpy as np


def project_cone3(f0, f1, f2, mu):
    """constraint_utils.cpp:5-25, scalar, same branches and order of operations as the device function."""
    r = np.sqrt(f0 * f0 + f1 * f1)
    if mu * f2 >= r:
        return 0, (f0, f1, f2)
    if f2 < -mu * r:
        return 1, (0.0, 0.0, 0.0)
    ratio = mu * f2 / r
    r0, r1, r2 = ratio * f0, ratio * f1, f2
    sq = r0 * r0 + r1 * r1 + r2 * r2
    if sq > 0.0:
        nr = np.sqrt(sq); r0, r1, r2 = r0 / nr, r1 / nr, r2 / nr
    d = r0 * f0 + r1 * f1 + r2 * f2
    return 2, (d * r0, d * r1, d * r2)


def prepare(qp, x, mux, muc):
    """-> (Qp, bp, Ap, beqp, rot): the polished equality-constrained QPs (same n, m) and the per-contact classification
    rot[B, nc/3, 4] = (kind, o^): 0 inside the cone, 1 apex, 2 boundary (tangent plane at o^)."""
    B, n, m, nc, lcs = qp.batch, qp.n, qp.m, qp.nc, qp.lambda_c_start
    ncon = nc // 3
    Qp = np.zeros((B, n, n)); bp = np.zeros((B, n)); Ap = np.zeros((B, m, n)); beqp = np.zeros((B, m))
    rot = np.zeros((B, max(ncon, 1), 4))
    lb = np.broadcast_to(qp.lb, (B, n)); ub = np.broadcast_to(qp.ub, (B, n))
    fr = np.broadcast_to(qp.friction_coeffs, (B, qp.friction_coeffs.shape[-1]))
    for q in range(B):
        T = np.eye(n)                     # x = T y  (identity outside rotated contacts)
        fixed = np.zeros(n, bool); val = np.zeros(n)
        for i in list(range(lcs)) + list(range(lcs + nc, n)):
            a = x[q, i] + mux[q, i]
            if a <= lb[q, i]:
                fixed[i], val[i] = True, lb[q, i]
            elif a >= ub[q, i]:
                fixed[i], val[i] = True, ub[q, i]
        for c in range(ncon):
            o = lcs + 3 * c
            f = x[q, o:o + 3] + muc[q, 3 * c:3 * c + 3]
            kind, proj = project_cone3(f[0], f[1], f[2], fr[q, c])
            no = np.sqrt(proj[0] ** 2 + proj[1] ** 2 + proj[2] ** 2)
            if kind == 0:
                continue
            if kind == 1 or not no > 0.0:
                rot[q, c, 0] = 1.0
                fixed[o:o + 3] = True
            else:
                d = np.array(proj) / no
                rot[q, c] = (2.0, d[0], d[1], d[2])
                T[o:o + 3, o] = d          # alpha (along the ray) takes the slot of variable o
                h = np.sqrt(d[0] * d[0] + d[1] * d[1])
                if h > 0.0:                # beta (horizontal tangent t1) the slot of o + 1; the normal direction is fixed at 0
                    T[o:o + 3, o + 1] = (-d[1] / h, d[0] / h, 0.0)
                else:
                    fixed[o + 1] = True
                fixed[o + 2] = True
        free = ~fixed
        Qy = T.T @ qp.Q[q] @ T; by = T.T @ qp.b[q]
        Ay = qp.A_eq[q] @ T if m else np.zeros((0, n))
        g = by + Qy @ val                 # coupling of the fixed values into the free rows
        Qo = np.eye(n); Qo[np.ix_(free, free)] = Qy[np.ix_(free, free)]
        bo = -val.copy(); bo[free] = g[free]
        Ao = np.zeros((m, n)); Ao[:, free] = Ay[:, free]
        Qp[q], bp[q], Ap[q] = Qo, bo, Ao
        if m:
            beqp[q] = qp.b_eq[q] - Ay @ val
    return Qp, bp, Ap, beqp, rot


def finish(qp, rot, y, y_status, z, bviol, fviol, eps_fcone, eps_bound, eps_objective=1e-3):
    """-> (z, bviol, fviol, polished): accepted QPs take the rotated-back point."""
    B, n, nc, lcs = qp.batch, qp.n, qp.nc, qp.lambda_c_start
    lb = np.broadcast_to(qp.lb, (B, n)); ub = np.broadcast_to(qp.ub, (B, n))
    fr = np.broadcast_to(qp.friction_coeffs, (B, qp.friction_coeffs.shape[-1]))
    z, bviol, fviol = z.copy(), bviol.copy(), fviol.copy()
    polished = np.zeros(B, np.int32)
    for q in range(B):
        xp = y[q].copy()
        for c in range(nc // 3):
            o = lcs + 3 * c
            if rot[q, c, 0] == 1.0:
                xp[o:o + 3] = 0.0
            elif rot[q, c, 0] == 2.0:
                d = rot[q, c, 1:4]
                h = np.sqrt(d[0] * d[0] + d[1] * d[1])
                t1 = np.array((-d[1] / h, d[0] / h, 0.0)) if h > 0.0 else np.zeros(3)
                xp[o:o + 3] = y[q, o] * d + y[q, o + 1] * t1
        ok = y_status[q] == 0 and np.all(np.isfinite(xp)) and np.all(xp >= lb[q] - eps_bound) and np.all(xp <= ub[q] + eps_bound)
        fv = 0.0
        for c in range(nc // 3):
            o = lcs + 3 * c
            v = np.sqrt(xp[o] ** 2 + xp[o + 1] ** 2) - fr[q, c] * xp[o + 2]
            ok = ok and bool(v <= eps_fcone)
            fv += max(v, 0.0)
        if qp.m:
            terms = qp.A_eq[q] * xp[None, :]
            mag = np.abs(qp.b_eq[q]) + np.abs(terms).sum(axis=1)
            floor = 1e-10 * np.abs(qp.A_eq[q]).sum(axis=1) * max(1.0, np.abs(xp).max())
            ok = ok and bool(np.all(np.abs(terms.sum(axis=1) - qp.b_eq[q]) <= 1e-7 * mag + floor))
        obj = lambda v: 0.5 * v @ qp.Q[q] @ v + qp.b[q] @ v
        fa = obj(z[q])
        ok = ok and bool(obj(xp) <= fa + eps_objective * max(1.0, abs(fa)))
        if ok:
            z[q] = xp
            bviol[q] = np.linalg.norm(xp - np.clip(xp, lb[q], ub[q]))
            fviol[q] = fv
            polished[q] = 1
    return z, bviol, fviol, polished


def polish(orc, qp, x, mux, muc, z, bviol, fviol, eps_fcone, eps_bound, eps_objective=1e-3, nthreads=8):
    """Whole step with `orc` (an oracle.Oracle) as the equality-constrained solver."""
    from fcc_qp_b200.logdata import QPBatch
    Qp, bp, Ap, beqp, rot = prepare(qp, x, mux, muc)
    n = qp.n
    inner = QPBatch(n, qp.m, 0, 0, Qp, bp, Ap, beqp, np.zeros((qp.batch, 0)), np.full((qp.batch, n), -np.inf),
                    np.full((qp.batch, n), np.inf))
    r = orc.solve_batch(inner, warm_mode=0, nthreads=nthreads, max_iter=10, rho=1e-3, eps_fcone=eps_fcone, eps_bound=eps_bound)
    return finish(qp, rot, r["z"], r["status"], z, bviol, fviol, eps_fcone, eps_bound, eps_objective) + (rot,)
