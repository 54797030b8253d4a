"""One batched call over several devices (fccqp_batch_solve_multi; SURVEY.md 8e: the batch shards trivially, one host
thread + stream set per device, host-side scatter and gather).  The 2-device cases skip on a 1-GPU box; the entry point
itself is exercised with a device list of one everywhere."""
import os

import numpy as np
import pytest

from conftest import LOG_OPTS

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(z, zref):
    return np.abs(z - zref).max(1) / np.maximum(1.0, np.abs(zref).max(1))


def n_devices():
    from fcc_qp_b200 import _native as nat
    return nat.lib().fccqp_device_count()


def run(qp, devices, warm_from=None):
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, device=devices)
    s.set_options(FCCQPOptionsB(**LOG_OPTS))
    s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    return s, s.GetSolution()


def check_log(sol):
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    assert rel_err(sol.z, gold["z"]).max() <= 1e-6
    assert np.array_equal(sol.details.n_iter, gold["n_iter"]) and np.array_equal(sol.details.solve_status, gold["status"])


def test_multi_entry_with_one_device(walking_log):
    _, sol = run(walking_log, [0])
    check_log(sol)


def test_log_split_over_two_devices(walking_log):
    if n_devices() < 2:
        pytest.skip("needs 2 CUDA devices")
    s, sol = run(walking_log, [0, 1])
    check_log(sol)
    # warm re-solve of the same batch: the carried state travels with its shard
    s.set_warm_start(True)
    s.Solve(walking_log.Q, walking_log.b, walking_log.A_eq, walking_log.b_eq, walking_log.friction_coeffs, walking_log.lb, walking_log.ub)
    from oracle import Oracle
    orc = Oracle("port")
    lanes = orc.lanes(walking_log.batch, walking_log.n, walking_log.m, walking_log.nc, walking_log.lambda_c_start)
    lanes.set_options(**LOG_OPTS)
    lanes.solve(walking_log, warm=False)
    ref = lanes.solve(walking_log, warm=True)
    sol2 = s.GetSolution()
    assert rel_err(sol2.z, ref["z"]).max() <= 1e-6
    assert (sol2.details.n_iter != ref["n_iter"]).mean() <= 0.02


def test_uneven_split_and_shared_vectors(walking_log):
    """B not divisible by the device count, bounds and friction coefficients shared by the whole batch."""
    if n_devices() < 2:
        pytest.skip("needs 2 CUDA devices")
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    qp = walking_log.take(np.arange(0, 301))
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, device=[1, 0])
    s.set_options(FCCQPOptionsB(**LOG_OPTS))
    assert (qp.lb == qp.lb[0]).all() and (qp.ub == qp.ub[0]).all()
    s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb[0], qp.ub[0])
    sol = s.GetSolution()
    assert rel_err(sol.z, gold["z"][:301]).max() <= 1e-6 and np.array_equal(sol.details.n_iter, gold["n_iter"][:301])


def test_second_device_alone(walking_log):
    if n_devices() < 2:
        pytest.skip("needs 2 CUDA devices")
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    import torch
    qp = walking_log
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, device=1)
    s.set_options(FCCQPOptionsB(**LOG_OPTS))
    s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    check_log(s.GetSolution())
    dev = torch.device("cuda:1")
    s2 = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start)
    s2.set_options(FCCQPOptionsB(**LOG_OPTS))
    s2.Solve(*[torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)])
    torch.cuda.synchronize(dev)
    sol = s2.GetSolution()
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    assert rel_err(sol.z.cpu().numpy(), gold["z"]).max() <= 1e-6
    assert np.array_equal(sol.details.n_iter.cpu().numpy(), gold["n_iter"])


def test_bad_device_lists(walking_log):
    from fcc_qp_b200.batch import FCCQPBatch
    qp = walking_log.take(np.arange(4))
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, device=[0, 0])
    with pytest.raises(RuntimeError, match="listed twice"):
        s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, device=[0, 99])
    with pytest.raises(RuntimeError):
        s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
