"""Opt-in solution polish on the device (fccqp_polish_prepare / _finish + a second batched solve; FCCQPBatch.Polish()) against
its numpy restatement (oracle/polish.py, builder-authored: the reference has no polish step) started from the SAME ADMM state,
with the reference's C restatement as the restatement's equality-constrained solver."""
import numpy as np
import pytest

import oracle
from oracle import polish as pol
from test_polish import small_qps, cone_projection_batch

pytestmark = pytest.mark.gpu


def gpu_solve_and_polish(qp, slack=None, **opts):
    import torch
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    dev = torch.device("cuda:0")
    args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(**opts))
    if slack is not None:
        s.polish_objective_slack = slack
    s.Solve(*args)
    a = s.GetSolution()
    admm = dict(z=a.z.cpu().numpy(), bv=a.details.bounds_viol.cpu().numpy(), fv=a.details.friction_cone_viol.cpu().numpy(),
                n_iter=a.details.n_iter.cpu().numpy(), status=a.details.solve_status.cpu().numpy())
    x, mux, muc = (t.clone().cpu().numpy() for t in s.GetState())
    p = s.Polish()
    torch.cuda.synchronize()
    out = dict(z=p.z.cpu().numpy(), bv=p.details.bounds_viol.cpu().numpy(), fv=p.details.friction_cone_viol.cpu().numpy(),
               flag=p.details.polished.cpu().numpy(), n_iter=p.details.n_iter.cpu().numpy(), status=p.details.solve_status.cpu().numpy())
    return admm, (x, mux, muc), out


def check_against_restatement(qp, admm, state, out, eps_f, eps_b, slack=1e-3):
    zr, bvr, fvr, flagr, rot = pol.polish(oracle.Oracle("port"), qp, state[0], state[1], state[2], admm["z"], admm["bv"], admm["fv"],
                                          eps_fcone=eps_f, eps_bound=eps_b, eps_objective=slack)
    same = out["flag"] == flagr
    assert same.mean() >= 0.99, (~same).sum()          # (a QP may sit on an acceptance threshold)
    rel = np.abs(out["z"] - zr).max(axis=1) / np.maximum(1.0, np.abs(zr).max(axis=1))
    assert rel[same].max() <= 1e-6, rel[same].max()
    acc = same & (flagr == 1)
    assert np.abs(out["bv"] - bvr)[acc].max() <= 1e-6 and np.abs(out["fv"] - fvr)[acc].max() <= 1e-6
    # rejected QPs keep the ADMM result bit for bit; counts and status are never touched
    rej = out["flag"] == 0
    assert np.array_equal(out["z"][rej], admm["z"][rej])
    assert np.array_equal(out["n_iter"], admm["n_iter"]) and np.array_equal(out["status"], admm["status"])
    return flagr


def test_polish_cone_projection_is_exact():
    qp, exact = cone_projection_batch(256, seed=3)
    admm, state, out = gpu_solve_and_polish(qp, slack=1e9, max_iter=8, rho=0.3, eps_fcone=1e-9, eps_bound=1e-9)
    check_against_restatement(qp, admm, state, out, 1e-9, 1e-9, slack=1e9)
    acc = out["flag"] == 1
    assert acc.mean() >= 0.95 and np.abs(admm["z"] - exact)[acc].max() > 1e-2
    assert np.abs(out["z"] - exact)[acc].max() <= 1e-9 and out["fv"][acc].max() <= 1e-9


def test_polish_small_qps_matches_restatement_and_improves():
    qp = small_qps(512, seed=11)
    admm, state, out = gpu_solve_and_polish(qp, max_iter=300, rho=3.0, eps_fcone=1e-6, eps_bound=1e-6)
    flag = check_against_restatement(qp, admm, state, out, 1e-6, 1e-6)
    acc = out["flag"] == 1
    assert acc.mean() >= 0.6         # (the rest: guesses that fix too much or leave a bound / cone violated -- left as ADMM wrote them)
    res = np.abs(np.einsum("bij,bj->bi", qp.A_eq, out["z"]) - qp.b_eq).max(axis=1)
    assert res[acc].max() <= 1e-9
    assert (out["z"][acc] >= qp.lb[acc] - 1e-6).all() and (out["z"][acc] <= qp.ub[acc] + 1e-6).all() and out["fv"][acc].max() <= 2e-6


def test_polish_walking_log(walking_log):
    """The logged Cassie QPs (structure kernel for both solves): against the restatement from the same state; the QPs that ended
    at max_iter with a friction-cone violation of up to 5.6e-4 come back inside the cones where accepted."""
    admm, state, out = gpu_solve_and_polish(walking_log, max_iter=100, rho=5e-5, eps_fcone=1e-6, eps_bound=1e-6)
    check_against_restatement(walking_log, admm, state, out, 1e-6, 1e-6)
    acc = out["flag"] == 1
    assert acc.mean() >= 0.9
    res = np.abs(np.einsum("bij,bj->bi", walking_log.A_eq, out["z"]) - walking_log.b_eq).max(axis=1)
    assert res[acc].max() <= 1e-7
    assert out["fv"][acc].max() <= 4e-6 and out["bv"][acc].max() <= 1e-6      # (fcone_viol sums over the four contacts)
    late = acc & (admm["status"] == 1)
    print("accepted", acc.mean(), "of the QPs at max_iter:", late.sum(), "/", (admm["status"] == 1).sum(),
          "cone violation before", admm["fv"][late].max() if late.any() else None, "after", out["fv"][late].max() if late.any() else None,
          "moved by (max over accepted)", np.abs(out["z"] - admm["z"])[acc].max())
