"""Opt-in ADMM over-relaxation (fccqp_options::relaxation) -- an extension that is NOT in the
reference.  Its oracle is the C restatement with the same three changed lines (oracle/fccqp_oracle.c, do_admm);
relaxation = 1 is the reference's iteration and is covered by every other parity test."""
import numpy as np
import pytest

import oracle
from conftest import LOG_OPTS
from fcc_qp_b200 import synthetic as syn


@pytest.fixture()
def port():
    o = oracle.Oracle("port")
    yield o
    o.set_relaxation(1.0)


def test_oracle_relaxation_reduces_iterations(port, walking_log):
    qp = walking_log.take(np.arange(0, 2019, 2))
    base = port.solve_batch(qp, warm_mode=0, nthreads=8, **LOG_OPTS)
    port.set_relaxation(1.5)
    rel = port.solve_batch(qp, warm_mode=0, nthreads=8, **LOG_OPTS)
    it0, it1 = base["n_iter"], rel["n_iter"]
    assert np.array_equal(it0 == 0, it1 == 0)                 # QPs that stop at the pre-solve point are untouched
    assert (it1 == 100).sum() < (it0 == 100).sum()            # fewer QPs run out of iterations
    assert rel["fcone_viol"].max() < 0.1 * base["fcone_viol"].max()
    quad = syn.make_batch(syn.QUADRUPED, 192)
    port.set_relaxation(1.0); a = port.solve_batch(quad, warm_mode=0, nthreads=8, **LOG_OPTS)["n_iter"]
    port.set_relaxation(1.5); b = port.solve_batch(quad, warm_mode=0, nthreads=8, **LOG_OPTS)["n_iter"]
    assert b[a > 0].mean() < 0.5 * a[a > 0].mean()


def test_relaxation_option_is_validated():
    from fcc_qp_b200 import _native as nat
    import ctypes as C
    o = nat.Options()
    nat.lib().fccqp_default_options(C.byref(o))
    assert o.relaxation == 1.0 and o.max_iter == 1000


@pytest.mark.gpu
@pytest.mark.parametrize("alpha", [1.5, 0.8])
def test_gpu_matches_relaxed_oracle(port, walking_log, alpha):
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    port.set_relaxation(alpha)
    long_running = 0
    for qp in (walking_log.take(np.arange(0, 2019, 2)), syn.make_batch(syn.HUMANOID, 96), syn.make_batch(syn.MULTICONTACT, 48)):
        ref = port.solve_batch(qp, warm_mode=0, nthreads=8, **LOG_OPTS)
        s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start)
        s.set_options(FCCQPOptionsB(relaxation=alpha, **LOG_OPTS))
        s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
        sol = s.GetSolution()
        err = np.abs(sol.z - ref["z"]).max(1) / np.maximum(1.0, np.abs(ref["z"]).max(1))
        same = sol.details.n_iter == ref["n_iter"]
        assert (~same).mean() <= 0.02
        assert err[same].max() <= 1e-6
        long_running += int((ref["n_iter"] > 8).sum())
    assert long_running > 0                     # the operator path for long-running QPs is exercised too


@pytest.mark.gpu
def test_gpu_relaxation_through_the_dropin_object_and_bad_values(walking_log):
    from fcc_qp import FCCQP, FCCQPOptions
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    o = FCCQPOptions()
    assert o.relaxation == 1.0
    o.rho, o.eps_fcone, o.eps_bound, o.max_iter, o.relaxation = 5e-5, 1e-6, 1e-6, 100, 1.5
    s = FCCQP(60, 38, 12, 38)
    s.set_options(o)
    idx = int(np.nonzero(np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "walking_cold.npz"))["n_iter"] == 6)[0][0])
    q = walking_log.qp(idx)
    s.Solve(q["Q"], q["b"], q["A_eq"], q["b_eq"], q["friction_coeffs"], q["lb"], q["ub"])
    assert 0 < s.GetSolution().details.n_iter < 6
    b = FCCQPBatch(60, 38, 12, 38)
    b.set_options(FCCQPOptionsB(relaxation=2.5, **LOG_OPTS))
    with pytest.raises(Exception):
        b.Solve(walking_log.Q[:2], walking_log.b[:2], walking_log.A_eq[:2], walking_log.b_eq[:2],
                walking_log.friction_coeffs[:2], walking_log.lb[:2], walking_log.ub[:2])
