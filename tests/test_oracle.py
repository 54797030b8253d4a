"""The CPU oracle (C restatement, oracle/fccqp_oracle.c) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_goldens.py), against the live compiled reference when
oracle/_ref is present, and against the known answers listed in SURVEY.md section 4."""
import os

import numpy as np
import pytest

import oracle
from conftest import LOG_OPTS

G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-8  # restatement vs compiled reference: rounding only (measured 7e-11)


def rel_err(z, zref):
    return (np.abs(z - zref).max(1) / np.maximum(1.0, np.abs(zref).max(1))).max()


@pytest.fixture(scope="module")
def port():
    oracle.build(ref=False)
    return oracle.Oracle("port")


def check(r, gold, scalars=True):
    assert rel_err(r["z"], gold["z"]) <= TOL
    assert np.array_equal(r["n_iter"], gold["n_iter"])
    assert np.array_equal(r["status"], gold["status"])
    if scalars:
        for k in ("res_bounds", "res_fcone", "bounds_viol", "fcone_viol"):
            assert np.abs(r[k] - gold[k]).max() <= 1e-6


def test_port_walking_cold(port, walking_log):
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    check(port.solve_batch(walking_log, warm_mode=0, nthreads=4, **LOG_OPTS), gold)
    # golden statistics quoted in SURVEY.md 8c
    h = dict(zip(*np.unique(gold["n_iter"], return_counts=True)))
    assert h == {0: 1978, 6: 14, 9: 1, 100: 26}


def test_port_walking_warm_sequential(port, walking_log):
    """fcc_qp_test.py:86-89: one solver object, set_warm_start(i > 0)."""
    gold = np.load(os.path.join(G, "walking_warm.npz"))
    check(port.solve_batch(walking_log, warm_mode=1, nthreads=1, **LOG_OPTS), gold)


def test_port_paper_settings(port, walking_log):
    gold = np.load(os.path.join(G, "walking_cold_paper.npz"))
    mi, rho, ef, eb = gold["opts"]
    check(port.solve_batch(walking_log, warm_mode=0, nthreads=4, max_iter=int(mi), rho=rho, eps_fcone=ef,
                           eps_bound=eb), gold)


@pytest.mark.parametrize("name,B", [("humanoid", 192), ("quadruped", 192), ("multicontact", 96)])
def test_port_synthetic(port, name, B):
    from fcc_qp_b200 import synthetic as syn
    gold = np.load(os.path.join(G, f"synthetic_{name}_cold.npz"))
    qp = syn.make_batch(syn.SHAPES[name], B)
    chk = np.array([qp.Q.sum(), qp.A_eq.sum(), qp.b.sum(), qp.b_eq.sum(), qp.friction_coeffs.sum()])
    assert np.allclose(chk, gold["in_sum"], rtol=1e-12), "synthetic generator drifted from the goldens"
    r = port.solve_batch(qp, warm_mode=0, nthreads=4, **LOG_OPTS)
    # the presolve takes the LDLT branch here (Q > 0), as in the reference
    assert rel_err(r["z"], gold["z"]) <= 1e-7
    assert (r["n_iter"] != gold["n_iter"]).mean() <= 0.02


def test_port_presolve_branch(port, walking_log):
    """On the log the reference's LDLT reports NumericalIssue and COD runs."""
    q = walking_log.qp(0)
    s = port.solver(60, 38, 12, 38)
    s.set_options(**LOG_OPTS)
    s.Solve(q["Q"], q["b"], q["A_eq"], q["b_eq"], q["friction_coeffs"], q["lb"], q["ub"])
    assert s.presolve_path() == 2
    s.set_warm_start(True)
    s.Solve(q["Q"], q["b"], q["A_eq"], q["b_eq"], q["friction_coeffs"], q["lb"], q["ub"])
    assert s.presolve_path() == 0


def test_port_kats(port):
    k = np.load(os.path.join(G, "kats.npz"))
    for f, z in zip(k["cone_f"], k["cone_z"]):
        s = port.solver(3, 0, 3, 0)
        s.set_options(2000, 1.0, 1e-10, 1e-10)
        s.Solve(np.eye(3), -f, np.zeros((0, 3)), np.zeros(0), [0.5], np.full(3, -np.inf), np.full(3, np.inf))
        assert np.abs(s.GetSolution()["z"] - z).max() < 1e-8
    # closed forms: f=(1,0,1)->(0.6,0,1.2); (3,4,1)->(0.84,1.12,2.8); polar -> 0;
    # inside -> f; f_z == 0 quirk -> 0
    assert np.allclose(k["cone_z"], [[0.6, 0, 1.2], [0.84, 1.12, 2.8], [0, 0, 0], [0, 0, 2], [0, 0, 0]], atol=1e-7)
    n, m = k["eq_Q"].shape[0], k["eq_A"].shape[0]
    s = port.solver(n, m, 0, 0)
    s.set_options(100, 1e-3, 1e-6, 1e-6)
    s.Solve(k["eq_Q"], k["eq_b"], k["eq_A"], k["eq_beq"], [], np.full(n, -np.inf), np.full(n, np.inf))
    r = s.GetSolution()
    assert r["n_iter"] == 0 and np.abs(r["z"] - k["eq_z"]).max() < 1e-10
    K = np.block([[k["eq_Q"], k["eq_A"].T], [k["eq_A"], np.zeros((m, m))]])
    assert np.abs(np.linalg.solve(K, np.concatenate([-k["eq_b"], k["eq_beq"]]))[:n] - r["z"]).max() < 1e-10


def test_port_cone_projection_direct(port):
    import ctypes as C
    out = np.zeros(3)
    dp = C.POINTER(C.c_double)
    for f, mu, want in [((1, 0, 1), 0.5, (0.6, 0, 1.2)), ((3, 4, 1), 0.5, (0.84, 1.12, 2.8)),
                        ((1, 0, -3), 0.5, (0, 0, 0)), ((0, 0, 2), 0.5, (0, 0, 2)), ((1, 1, 0), 0.5, (0, 0, 0))]:
        fa = np.array(f, dtype=np.float64)
        port.lib.fccqp_oracle_project_cone3(fa.ctypes.data_as(dp), mu, out.ctypes.data_as(dp))
        assert np.allclose(out, want, atol=1e-12)


def test_port_too_few_friction_coeffs(port):
    s = port.solver(6, 0, 6, 0)
    with pytest.raises(IndexError):
        s.Solve(np.eye(6), np.zeros(6), np.zeros((0, 6)), np.zeros(0), [0.5], np.full(6, -np.inf), np.full(6, np.inf))


@pytest.mark.skipif(not oracle.have("ref") and not os.path.isdir("/root/reference/src"),
                    reason="compiled reference (oracle/_ref) not available")
def test_port_matches_live_reference(port, walking_log):
    ref = oracle.Oracle("ref")
    sub = walking_log.take(np.arange(0, 2019, 4))
    for wm, nt in ((0, 4), (1, 1)):
        a = ref.solve_batch(sub, warm_mode=wm, nthreads=nt, **LOG_OPTS)
        b = port.solve_batch(sub, warm_mode=wm, nthreads=nt, **LOG_OPTS)
        check(b, a)
