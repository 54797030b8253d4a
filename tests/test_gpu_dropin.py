"""The drop-in surface: the reference's own replay script (fcc_qp_test.py:72-91, minus the
matplotlib plots) run unchanged against `from fcc_qp import FCCQP, FCCQPSolution, FCCQPOptions`."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def solve(qp, solver):  # fcc_qp_test.py:27-32, verbatim call shape
    solver.Solve(qp['Q'], qp['b'], qp['A_eq'], qp['b_eq'], qp['friction_coeffs'], qp['lb'], qp['ub'])
    return solver.GetSolution()


def test_reference_replay_script(walking_log):
    from fcc_qp import FCCQP, FCCQPOptions, FCCQPSolution
    gold = np.load(os.path.join(G, "walking_warm.npz"))
    qps = [walking_log.qp(i) for i in range(walking_log.batch)]
    solver = FCCQP(60, 38, 12, 38)           # fcc_qp_test.py:77
    options = FCCQPOptions()
    options.rho = 5e-5
    options.eps_fcone = 1e-6
    options.eps_bound = 1e-6
    options.max_iter = 100
    solver.set_options(options)
    results = []
    for i in range(len(qps)):
        solver.set_warm_start(i > 0)
        results.append(solve(qps[i], solver))
    assert isinstance(results[0], FCCQPSolution)
    z = np.vstack([r.z for r in results])
    n = np.array([r.details.n_iter for r in results])
    err = np.abs(z - gold["z"]).max(1) / np.maximum(1.0, np.abs(gold["z"]).max(1))
    assert err.max() <= 1e-6
    assert np.array_equal(n, gold["n_iter"])
    assert np.abs(np.array([r.details.friction_cone_viol for r in results]) - gold["fcone_viol"]).max() <= 1e-5
    assert np.abs(np.array([r.details.bounds_viol for r in results]) - gold["bounds_viol"]).max() <= 1e-5
    assert np.abs(np.array([r.details.eps_friction_cone for r in results]) - gold["res_fcone"]).max() <= 1e-5
    assert all(r.details.solve_time > 0 for r in results[:5])
    assert np.array_equal(np.array([r.details.solve_status for r in results]), gold["status"])


def test_keyword_arguments_and_layouts(walking_log):
    """kwargs as bound in src/main.cpp:43-53; F-ordered, float32 and list inputs are accepted
    (pybind forcecast)."""
    from fcc_qp import FCCQP, FCCQPOptions
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    q = walking_log.qp(3)
    s = FCCQP(num_vars=60, num_equality_constraints=38, nc=12, lambda_c_start=38)
    o = FCCQPOptions(); o.rho, o.eps_fcone, o.eps_bound, o.max_iter = 5e-5, 1e-6, 1e-6, 100
    s.set_options(o)
    s.Solve(Q=np.asfortranarray(q['Q']), b=q['b'], A_eq=np.asfortranarray(q['A_eq']), b_eq=list(q['b_eq']),
            friction_coeffs=list(q['friction_coeffs']), lb=q['lb'], ub=q['ub'])
    z = s.GetSolution().z
    assert np.abs(z - gold["z"][3]).max() / max(1.0, np.abs(gold["z"][3]).max()) <= 1e-6
    s.set_rho(5e-5); s.set_max_iter(100)
    s.Solve(q['Q'][:, ::1], q['b'], q['A_eq'], q['b_eq'], tuple(q['friction_coeffs']), q['lb'], q['ub'])
    assert np.abs(s.GetSolution().z - z).max() == 0.0
    with pytest.raises(IndexError):
        s.Solve(q['Q'], q['b'], q['A_eq'], q['b_eq'], (0.6,), q['lb'], q['ub'])   # .at(i) -> std::out_of_range
    with pytest.raises(ValueError):
        s.Solve(q['Q'][:59, :59], q['b'], q['A_eq'], q['b_eq'], q['friction_coeffs'], q['lb'], q['ub'])


def test_warm_state_roundtrip(walking_log):
    from fcc_qp import FCCQP, FCCQPOptions
    o = FCCQPOptions(); o.rho, o.eps_fcone, o.eps_bound, o.max_iter = 5e-5, 1e-6, 1e-6, 100
    a, b = FCCQP(60, 38, 12, 38), FCCQP(60, 38, 12, 38)
    a.set_options(o); b.set_options(o)
    for i in range(40, 44):
        a.set_warm_start(i > 40)
        q = walking_log.qp(i)
        a.Solve(q['Q'], q['b'], q['A_eq'], q['b_eq'], q['friction_coeffs'], q['lb'], q['ub'])
    b.SetWarmState(*a.GetWarmState())           # migrate the carried state to another object
    q = walking_log.qp(44)
    for s in (a, b):
        s.set_warm_start(True)
        s.Solve(q['Q'], q['b'], q['A_eq'], q['b_eq'], q['friction_coeffs'], q['lb'], q['ub'])
    assert np.array_equal(a.GetSolution().z, b.GetSolution().z)


def test_cpp_eigen_caller():
    """tests/cpp/dropin_main (built where Eigen headers exist, shipped with the repo snapshot): equality-only
    closed form, cone-projection KAT, warm restart, Eigen::Ref blocks with an outer stride -- through
    include/fcc_qp.hpp and the C ABI."""
    import subprocess
    exe = os.path.join(os.path.dirname(__file__), "cpp", "dropin_main")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/dropin_main not built (needs Eigen headers at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dropin_main: ok" in r.stdout
