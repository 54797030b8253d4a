"""fcc_qp::FCCQPBatch (include/fcc_qp.hpp) as bound in the pybind11 module: host stacks and DLPack device tensors."""
import os

import numpy as np
import pytest

from conftest import LOG_OPTS

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(z, zref):
    return np.abs(z - zref).max(1) / np.maximum(1.0, np.abs(zref).max(1))


def make(qp):
    from fcc_qp_b200.batch import FCCQPBatchCpp, FCCQPOptionsB
    s = FCCQPBatchCpp(qp.n, qp.m, qp.nc, qp.lambda_c_start)
    s.set_options(FCCQPOptionsB(**LOG_OPTS))
    return s


def test_host_stacks_match_reference(walking_log):
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    qp = walking_log
    s = make(qp)
    s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    r = s.GetSolution()
    assert rel_err(r.z, gold["z"]).max() <= 1e-6
    assert np.array_equal(r.details.n_iter, gold["n_iter"]) and np.array_equal(r.details.solve_status, gold["status"])
    assert np.abs(r.details.eps_friction_cone - gold["res_fcone"]).max() <= 1e-5 * max(1.0, np.abs(gold["z"]).max())


def test_dlpack_device_tensors_match_reference_and_reuse_outputs(walking_log):
    import torch
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    qp = walking_log
    dev = torch.device("cuda:0")
    args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    s = make(qp)
    s.Solve(*args)
    r = s.GetSolution()
    torch.cuda.synchronize()
    ptr = r.z.data_ptr()
    assert rel_err(r.z.cpu().numpy(), gold["z"]).max() <= 1e-6
    assert np.array_equal(r.details.n_iter.cpu().numpy(), gold["n_iter"])
    assert np.array_equal(r.details.solve_status.cpu().numpy(), gold["status"])
    # warm re-solve of the same QPs from the carried state: converged lanes stay converged, buffers are reused
    s.set_warm_start(True)
    s.Solve(*args)
    r2 = s.GetSolution()
    torch.cuda.synchronize()
    assert r2.z.data_ptr() == ptr
    done = gold["status"] == 0
    assert (r2.details.n_iter.cpu().numpy()[done] <= 2).all()
    # strided views (column-major A_eq, a shared bound vector) are consumed in place
    A_cm = args[2].transpose(1, 2).contiguous().transpose(1, 2)
    s.set_warm_start(False)
    s.Solve(args[0], args[1], A_cm, args[3], args[4], args[5][0], args[6][0])
    torch.cuda.synchronize()
    assert rel_err(s.GetSolution().z.cpu().numpy(), gold["z"]).max() <= 1e-6


def test_dlpack_rejects_host_and_wrong_dtype(walking_log):
    import torch
    qp = walking_log.take(np.arange(4))
    s = make(qp)
    dev = torch.device("cuda:0")
    good = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    bad = list(good); bad[0] = good[0].float()
    with pytest.raises(TypeError):
        s.Solve(*bad)
    cpu = [torch.as_tensor(a) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    with pytest.raises((TypeError, RuntimeError, ValueError)):
        s.Solve(*cpu)
