"""Edge cases of the reference's behaviour, on the GPU through the C ABI."""
import os

import numpy as np
import pytest

import oracle
from conftest import LOG_OPTS

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def batch_solver(n, m, nc, lcs, **opts):
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    s = FCCQPBatch(n, m, nc, lcs)
    s.set_options(FCCQPOptionsB(**opts))
    return s


INF3 = (np.full(3, -np.inf), np.full(3, np.inf))


def test_cone_known_answers():
    k = np.load(os.path.join(G, "kats.npz"))
    f = k["cone_f"]
    B = f.shape[0]
    s = batch_solver(3, 0, 3, 0, max_iter=2000, rho=1.0, eps_fcone=1e-10, eps_bound=1e-10)
    s.Solve(np.tile(np.eye(3), (B, 1, 1)), -f, np.zeros((B, 0, 3)), np.zeros((B, 0)), [0.5], *INF3)
    z = s.GetSolution().z
    assert np.abs(z - k["cone_z"]).max() < 1e-7
    assert np.abs(z[4]).max() == 0.0     # f_z == 0 quirk: projects to the origin (constraint_utils.cpp:20-23)


def test_equality_only_closed_form():
    k = np.load(os.path.join(G, "kats.npz"))
    n, m = k["eq_Q"].shape[0], k["eq_A"].shape[0]
    s = batch_solver(n, m, 0, 0, max_iter=100, rho=1e-3, eps_fcone=1e-6, eps_bound=1e-6)
    s.Solve(k["eq_Q"][None], k["eq_b"][None], k["eq_A"][None], k["eq_beq"][None], np.zeros(0), np.full(n, -np.inf),
            np.full(n, np.inf))
    r = s.GetSolution()
    assert r.details.n_iter[0] == 0 and r.details.solve_status[0] == 0
    assert np.abs(r.z[0] - k["eq_z"]).max() < 1e-9
    assert r.details.eps_bounds[0] == 0.0 and r.details.eps_friction_cone[0] == 0.0


def test_empty_and_single_batches(walking_log):
    qp = walking_log
    s = batch_solver(qp.n, qp.m, qp.nc, qp.lambda_c_start, **LOG_OPTS)
    e = qp.take(np.arange(0))
    s.Solve(e.Q, e.b, e.A_eq, e.b_eq, qp.friction_coeffs[0], qp.lb[0], qp.ub[0])
    assert s.GetSolution().z.shape == (0, qp.n)
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    one = qp.take(np.array([17]))
    s.Solve(one.Q, one.b, one.A_eq, one.b_eq, one.friction_coeffs, one.lb, one.ub)
    assert np.abs(s.GetSolution().z[0] - gold["z"][17]).max() / np.abs(gold["z"][17]).max() < 1e-6


def test_no_contacts_with_finite_bounds():
    """nc = 0 with finite bounds segfaults the reference in Release (fcc_qp.cpp:98); here the
    infinity norm of an empty residual is 0 and the box-constrained QP is solved."""
    rng = np.random.default_rng(3)
    B, n, m = 5, 8, 3
    G_ = rng.standard_normal((B, n, n)); Q = G_ @ G_.transpose(0, 2, 1) + np.eye(n)
    A = rng.standard_normal((B, m, n)); b = rng.standard_normal((B, n)) * 3; beq = rng.standard_normal((B, m))
    lb, ub = np.full(n, -0.3), np.full(n, 0.3)
    opts = dict(max_iter=4000, rho=1.0, eps_fcone=1e-9, eps_bound=1e-9)
    s = batch_solver(n, m, 0, 0, **opts)
    s.Solve(Q, b, A, beq, np.zeros(0), lb, ub)
    r = s.GetSolution()
    from fcc_qp_b200.logdata import QPBatch
    ref = oracle.Oracle("port").solve_batch(QPBatch(n, m, 0, 0, Q, b, A, beq, np.zeros((B, 0)), np.tile(lb, (B, 1)),
                                                    np.tile(ub, (B, 1))), warm_mode=0, **opts)
    assert np.abs(r.z - ref["z"]).max() < 1e-7
    assert np.array_equal(r.details.n_iter, ref["n_iter"])
    assert (r.details.eps_friction_cone == 0).all()


def test_no_equality_rows():
    rng = np.random.default_rng(5)
    B, n = 4, 9
    G_ = rng.standard_normal((B, n, n)); Q = G_ @ G_.transpose(0, 2, 1) + np.eye(n)
    b = rng.standard_normal((B, n)) * 2
    opts = dict(max_iter=3000, rho=0.5, eps_fcone=1e-9, eps_bound=1e-9)
    s = batch_solver(n, 0, 9, 0, **opts)
    mu = rng.uniform(0.3, 0.9, (B, 3))
    s.Solve(Q, b, np.zeros((B, 0, n)), np.zeros((B, 0)), mu, np.full(n, -np.inf), np.full(n, np.inf))
    r = s.GetSolution()
    from fcc_qp_b200.logdata import QPBatch
    ref = oracle.Oracle("port").solve_batch(QPBatch(n, 0, 9, 0, Q, b, np.zeros((B, 0, n)), np.zeros((B, 0)), mu,
                                                    np.full((B, n), -np.inf), np.full((B, n), np.inf)), warm_mode=0, **opts)
    assert np.abs(r.z - ref["z"]).max() < 1e-7 and np.array_equal(r.details.n_iter, ref["n_iter"])


def test_max_iter_boundaries(walking_log):
    """n_iter is 0-based and status derives from n_iter == max_iter (fcc_qp.cpp:107,203): a QP that
    converges on the last permitted iteration is a success."""
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    i6 = int(np.nonzero(gold["n_iter"] == 6)[0][0])
    one = walking_log.take(np.array([i6]))
    for mi, want_it, want_st in ((7, 6, 0), (6, 6, 1), (1, 1, 1)):
        s = batch_solver(60, 38, 12, 38, **dict(LOG_OPTS, max_iter=mi))
        s.Solve(one.Q, one.b, one.A_eq, one.b_eq, one.friction_coeffs, one.lb, one.ub)
        d = s.GetSolution().details
        assert (d.n_iter[0], d.solve_status[0]) == (want_it, want_st), mi


def test_strided_device_inputs(walking_log):
    """Column-major A_eq and non-contiguous Q stacks are consumed in place on the device."""
    import torch
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    qp = walking_log.take(np.arange(64))
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(a, device=dev)
    A_cm = t(np.ascontiguousarray(qp.A_eq.transpose(0, 2, 1))).transpose(1, 2)      # [B,m,n] view, column-major
    Qpad = torch.zeros((64, 60, 64), dtype=torch.float64, device=dev); Qpad[:, :, :60] = t(qp.Q)
    s = batch_solver(60, 38, 12, 38, **LOG_OPTS)
    s.Solve(Qpad[:, :, :60], t(qp.b), A_cm, t(qp.b_eq), t(qp.friction_coeffs[0]), t(qp.lb[0]), t(qp.ub[0]))
    z = s.GetSolution().z.cpu().numpy()
    err = np.abs(z - gold["z"][:64]).max(1) / np.maximum(1, np.abs(gold["z"][:64]).max(1))
    assert err.max() <= 1e-6
    assert np.array_equal(s.GetSolution().details.n_iter.cpu().numpy(), gold["n_iter"][:64])


@pytest.mark.parametrize("n,m", [(6, 3), (40, 12)])      # warp kernel / CTA kernels
@pytest.mark.parametrize("perturb,consistent", [(0.0, True), (1e-14, True), (0.0, False)])
def test_rank_deficient_constraints(n, m, perturb, consistent):
    """(Nearly) duplicate equality rows make the KKT matrix singular.  The reference's LDLT fails and its COD fall-back
    returns the minimum-norm answer (src/fcc_qp.cpp:164-177), whose x part is the unique minimiser when the dependent
    rows are CONSISTENT.  The device factorization is unpivoted: it retries once with a regularised constraint block
    (kRegDelta), which reproduces that x, and checks A_eq x = b_eq on the result.  Bar: either the answer equals the
    compiled reference's within 1e-6 under the same status, or the status is FCCQP_STATUS_NUMERICAL_ISSUE -- never a
    different answer under a success status.  Consistent cases must be solved, the inconsistent one (where COD returns a
    least-squares compromise) must be flagged."""
    import oracle
    rng = np.random.default_rng(11 + n)
    G = rng.standard_normal((n, n))
    Q = (G @ G.T / n + np.eye(n))[None]
    A = rng.standard_normal((1, m, n)); A[0, 2] = A[0, 1] * (1.0 + perturb)
    beq = rng.standard_normal((1, m)); beq[0, 2] = beq[0, 1] if consistent else beq[0, 1] + 0.7
    b = rng.standard_normal((1, n))
    opts = dict(max_iter=10, rho=1e-3, eps_fcone=1e-6, eps_bound=1e-6)
    s = batch_solver(n, m, 0, 0, **opts)
    lb, ub = np.full(n, -np.inf), np.full(n, np.inf)
    s.Solve(Q, b, A, beq, np.zeros(0), lb, ub)
    r = s.GetSolution()
    o = oracle.Oracle("ref" if oracle.have("ref") else "port").solver(n, m, 0, 0)
    o.set_options(opts["max_iter"], opts["rho"], opts["eps_fcone"], opts["eps_bound"])
    o.Solve(Q[0], b[0], A[0], beq[0], [], lb, ub)
    ref = o.GetSolution()
    same = np.isfinite(r.z).all() and np.abs(r.z[0] - ref["z"]).max() <= 1e-6 * max(1.0, np.abs(ref["z"]).max())
    st = int(r.details.solve_status[0])
    assert st == 2 or (same and st == ref["status"]), (st, r.z[0], ref["z"])
    if consistent:
        assert st == ref["status"] and same
        assert np.abs(A[0] @ r.z[0] - beq[0]).max() <= 1e-8 * max(1.0, np.abs(beq).max())
    else:
        assert st == 2


def test_rank_deficient_constraints_in_a_batch():
    """The same through the reduced kernel's hand-over: a structured batch with one degenerate QP in it (two identical,
    consistent dynamics rows).  On this input the reference's LDLT does NOT report failure (its pivot test misses the
    dependent row, `presolve_path` 1), no COD runs, and FCCQP::Solve returns a point that violates its own equality
    constraints by ~10 under status 0 -- there is nothing to be in parity with.  The device path returns the actual
    minimiser: checked against the null-space solution of the equality-constrained QP (numpy), and A_eq z = b_eq."""
    import oracle
    from fcc_qp_b200 import synthetic as syn
    qp = syn.make_batch(syn.QUADRUPED, 192)              # (the golden set)
    A = qp.A_eq.copy(); beq = qp.b_eq.copy()
    A[5, 7] = A[5, 3]; beq[5, 7] = beq[5, 3]            # QP 5: two identical, consistent rows
    s = batch_solver(qp.n, qp.m, qp.nc, qp.lambda_c_start, **LOG_OPTS)
    s.Solve(qp.Q, qp.b, A, beq, qp.friction_coeffs, qp.lb, qp.ub)
    r = s.GetSolution()
    o = oracle.Oracle("ref" if oracle.have("ref") else "port").solver(qp.n, qp.m, qp.nc, qp.lambda_c_start)
    o.set_options(LOG_OPTS["max_iter"], LOG_OPTS["rho"], LOG_OPTS["eps_fcone"], LOG_OPTS["eps_bound"])
    o.Solve(qp.Q[5], qp.b[5], A[5], beq[5], qp.friction_coeffs[5], qp.lb[5], qp.ub[5])
    ref = o.GetSolution()
    assert np.abs(A[5] @ ref["z"] - beq[5]).max() > 1.0          # the reference's answer is not a solution
    assert int(r.details.solve_status[5]) == 0 and int(r.details.n_iter[5]) == 0
    _, sv, Vt = np.linalg.svd(A[5])
    rk = int((sv > 1e-10 * sv[0]).sum())
    assert rk == qp.m - 1
    Z = Vt[rk:].T
    xp = np.linalg.lstsq(A[5], beq[5], rcond=None)[0]
    xt = xp + Z @ np.linalg.solve(Z.T @ qp.Q[5] @ Z, -Z.T @ (qp.Q[5] @ xp + qp.b[5]))
    assert np.abs(r.z[5] - xt).max() <= 1e-6 * max(1.0, np.abs(xt).max())
    assert np.abs(A[5] @ r.z[5] - beq[5]).max() <= 1e-8 * max(1.0, np.abs(beq[5]).max())
    others = np.arange(192) != 5
    gold = np.load(os.path.join(G, "synthetic_quadruped_cold.npz"))
    assert (np.abs(r.z[others] - gold["z"][others]).max(1) <= 1e-6 * np.maximum(1.0, np.abs(gold["z"][others]).max(1))).all()


def test_well_conditioned_neighbours_are_not_flagged():
    """The pivot-size test must not fire on a merely badly scaled problem: costs from 1e-6 to 1e4 and constraint rows
    scaled from 1e-3 to 1e3, full row rank."""
    import oracle
    rng = np.random.default_rng(12)
    n, m, B = 12, 5, 16
    Q = np.zeros((B, n, n)); A = np.zeros((B, m, n))
    for k in range(B):
        G = rng.standard_normal((n, n)); sc = 10.0 ** rng.uniform(-3, 2, n)
        Q[k] = (G @ G.T + np.eye(n)) * sc[:, None] * sc[None, :]
        A[k] = rng.standard_normal((m, n)) * (10.0 ** rng.uniform(-3, 3, m))[:, None]
    b = rng.standard_normal((B, n)); beq = rng.standard_normal((B, m))
    s = batch_solver(n, m, 0, 0, max_iter=10, rho=1e-3, eps_fcone=1e-6, eps_bound=1e-6)
    lb, ub = np.full(n, -np.inf), np.full(n, np.inf)
    s.Solve(Q, b, A, beq, np.zeros(0), lb, ub)
    r = s.GetSolution()
    assert (r.details.solve_status == 0).all(), r.details.solve_status
    res = np.abs(np.einsum("bij,bj->bi", A, r.z) - beq).max(1) / np.maximum(1.0, np.abs(A).max((1, 2)) * np.abs(r.z).max(1))
    assert res.max() <= 1e-8


def test_too_large_problem_is_refused():
    from fcc_qp_b200 import FCCQPError
    n, m = 300, 100
    s = batch_solver(n, m, 0, 0, **LOG_OPTS)
    with pytest.raises(FCCQPError) as e:
        s.Solve(np.tile(np.eye(n), (1, 1, 1)), np.zeros((1, n)), np.zeros((1, m, n)), np.zeros((1, m)), np.zeros(0),
                np.full(n, -np.inf), np.full(n, np.inf))
    assert e.value.code == -3
