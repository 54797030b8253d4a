"""Opt-in solution polish: the numpy restatement (oracle/polish.py, builder-authored -- the reference has no polish step) on a
problem whose answer is known in closed form, with the reference's C restatement as the equality-constrained solver."""
import numpy as np

import oracle
from oracle import polish as pol


def cone_projection_batch(B, seed=0):
    """min 1/2 |lam - f|^2 over one friction cone: the minimiser is project_to_friction_cone(f)."""
    from fcc_qp_b200.logdata import QPBatch
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((B, 3)) * 2
    mu = rng.uniform(0.3, 1.0, (B, 1))
    qp = QPBatch(3, 0, 3, 0, np.tile(np.eye(3), (B, 1, 1)), -f, np.zeros((B, 0, 3)), np.zeros((B, 0)), mu,
                 np.full((B, 3), -np.inf), np.full((B, 3), np.inf))
    exact = np.array([pol.project_cone3(f[i, 0], f[i, 1], f[i, 2], mu[i, 0])[1] for i in range(B)])
    return qp, exact


def admm_state(qp, **opts):
    """ADMM through the C restatement, one solver object per QP: (z, x, mu_x, mu_c, bounds_viol, fcone_viol)."""
    o = oracle.Oracle("port")
    out = [[] for _ in range(6)]
    for i in range(qp.batch):
        s = o.solver(qp.n, qp.m, qp.nc, qp.lambda_c_start)
        s.set_options(**opts)
        q = qp.qp(i)
        s.Solve(q["Q"], q["b"], q["A_eq"], q["b_eq"], q["friction_coeffs"], q["lb"], q["ub"])
        sol, st = s.GetSolution(), s.get_state()
        for k, v in enumerate((sol["z"], st[0], st[1], st[2], sol["bounds_viol"], sol["fcone_viol"])):
            out[k].append(v)
    return [np.array(v) for v in out]


def test_polish_restatement_lands_on_the_exact_cone_projection():
    qp, exact = cone_projection_batch(96)
    opts = dict(max_iter=8, rho=0.3, eps_fcone=1e-9, eps_bound=1e-9)       # eight iterations: nowhere near converged
    z, x, mux, muc, bv, fv = admm_state(qp, **opts)
    admm_err = np.abs(z - exact).max(axis=1)
    # (objective test switched off: these ADMM points are far outside the cone, their objective is far below the optimum --
    #  the mechanics are under test here: classification, tangent plane, rotation back)
    zp, bvp, fvp, flag, rot = pol.polish(oracle.Oracle("port"), qp, x, mux, muc, z, bv, fv, eps_fcone=1e-9, eps_bound=1e-9,
                                         eps_objective=1e9)
    assert flag.mean() >= 0.95
    acc = flag == 1
    assert np.abs(zp[acc] - exact[acc]).max() <= 1e-9          # exact to rounding where ADMM was at 1e-3 ... 1e-6
    assert np.array_equal(zp[~acc], z[~acc])                   # rejected QPs keep the ADMM point
    assert admm_err[acc].max() > 1e-2                          # ... so the polish did something
    # with the default objective test the same far-from-converged points are (conservatively) left alone
    flag_default = pol.polish(oracle.Oracle("port"), qp, x, mux, muc, z, bv, fv, eps_fcone=1e-9, eps_bound=1e-9)[3]
    assert flag_default.mean() <= 0.3
    assert set(np.unique(rot[:, 0, 0])) == {0.0, 1.0, 2.0}     # all three branches of the projection present
    assert fvp[acc].max() <= 1e-9


def small_qps(B, seed=4):
    """Random small QPs with equality constraints, bounds outside the contact block and two friction cones."""
    from fcc_qp_b200.synthetic import random_qps
    qp = random_qps(np.random.default_rng(seed), B, 12, 6, 6, 3)
    qp.lb[:, 3:9] = -np.inf; qp.ub[:, 3:9] = np.inf         # (bounds on contact variables are not part of the polish guess)
    return qp


def test_polish_with_active_bounds_and_equalities():
    """ADMM stopped at eps 1e-6, then polished: accepted points satisfy A_eq z = b_eq to rounding and every bound and cone to
    the tolerance, and are closer to the solution (two ADMM runs to 1e-11 with different rho, where they agree) than the
    ADMM point was; a converged run is a fixed point of the polish."""
    qp = small_qps(96)
    o = oracle.Oracle("port")
    L1 = admm_state(qp, max_iter=100000, rho=3.0, eps_fcone=1e-11, eps_bound=1e-11)
    L2 = admm_state(qp, max_iter=100000, rho=10.0, eps_fcone=1e-11, eps_bound=1e-11)
    known = np.abs(L1[0] - L2[0]).max(axis=1) <= 1e-8
    assert known.mean() >= 0.8
    z, x, mux, muc, bv, fv = admm_state(qp, max_iter=300, rho=3.0, eps_fcone=1e-6, eps_bound=1e-6)
    zp, bvp, fvp, flag, rot = pol.polish(o, qp, x, mux, muc, z, bv, fv, eps_fcone=1e-6, eps_bound=1e-6)
    acc = flag == 1
    assert acc.mean() >= 0.8
    res = np.abs(np.einsum("bij,bj->bi", qp.A_eq, zp) - qp.b_eq).max(axis=1)
    assert res[acc].max() <= 1e-9
    assert (zp[acc] >= qp.lb[acc] - 1e-6).all() and (zp[acc] <= qp.ub[acc] + 1e-6).all() and fvp[acc].max() <= 2e-6
    S = acc & known
    ea, ep = np.abs(z - L1[0]).max(axis=1)[S], np.abs(zp - L1[0]).max(axis=1)[S]
    assert np.median(ep) <= 0.5 * np.median(ea) and ep.max() <= ea.max(), (np.median(ep), np.median(ea), ep.max(), ea.max())
    # fixed point: polishing the converged run moves it by rounding only
    zc, bvc, fvc, flagc, _ = pol.polish(o, qp, L1[1], L1[2], L1[3], L1[0], L1[4], L1[5], eps_fcone=1e-9, eps_bound=1e-9)
    okc = (flagc == 1) & known
    assert okc.mean() >= 0.8 and np.median(np.abs(zc - L1[0]).max(axis=1)[okc]) <= 1e-9
