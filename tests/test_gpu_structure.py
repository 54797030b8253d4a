"""The structure-exploiting kernel (fcc_qp_b200/csrc/fccqp_struct.cuh): QPs whose
separable variables are eliminated analytically must give the reference's answers like any other --
same bar as tests/test_gpu_parity.py (1e-6 relative on z and the objective, identical iteration
counts and status against the goldens of the UNMODIFIED reference), and QPs without such structure
must be handed to the general kernel inside the same call."""
import os

import numpy as np
import pytest

from conftest import LOG_OPTS

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(z, zref):
    return np.abs(z - zref).max(1) / np.maximum(1.0, np.abs(zref).max(1))


def solve(qp, structure="probe", opts=LOG_OPTS, device_resident=True, warm_state=None, refine=False):
    import torch
    from fcc_qp_b200 import _native as nat
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start)
    s.set_options(FCCQPOptionsB(**opts))
    s.structure = structure
    s.refine = refine
    args = (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    if device_resident:
        args = [torch.as_tensor(a, device="cuda:0") for a in args]
    if warm_state is not None:
        s.SetState(*warm_state)
        s.set_warm_start(True)
    s.Solve(*args)
    if device_resident:
        torch.cuda.synchronize()
    sol = s.GetSolution()
    to = (lambda a: a.cpu().numpy()) if device_resident else np.asarray
    return to(sol.z), to(sol.details.n_iter), to(sol.details.solve_status), nat.last_struct_info(), s


def test_walking_log_reduced_kernel_matches_reference(walking_log):
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    z, it, st, info, _ = solve(walking_log)
    assert info["used"] and info["rows"] == 72 and info["rows_dense"] == 104 and info["deferred"] == 0, info
    assert rel_err(z, gold["z"]).max() <= 1e-6
    o, oref = walking_log.objective(z), walking_log.objective(gold["z"])
    assert (np.abs(o - oref) / np.maximum(1.0, np.abs(oref))).max() <= 1e-6
    assert np.array_equal(it, gold["n_iter"]) and np.array_equal(st, gold["status"])


def test_refined_reduced_kernel_is_as_close_as_the_general_one(walking_log):
    """FCCQP_STRUCTURE_REFINE: one step of iterative refinement of the reduced pre-solve against the original Q
    and A_eq brings it to the level of the general kernel (7e-11 in the numpy model; 1.2e-7 without)."""
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    z, it, st, info, _ = solve(walking_log, refine=True)
    assert info["used"]
    assert rel_err(z, gold["z"]).max() <= 5e-9
    assert np.array_equal(it, gold["n_iter"]) and np.array_equal(st, gold["status"])
    # ... also for column-major A_eq, which takes the strided residual sweep
    import torch
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    qp = walking_log.take(np.arange(0, 512))
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start)
    s.set_options(FCCQPOptionsB(**LOG_OPTS))
    s.structure, s.refine = "probe", True
    t = [torch.as_tensor(a, device="cuda:0") for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    t[2] = t[2].transpose(1, 2).contiguous().transpose(1, 2)
    s.Solve(*t)
    torch.cuda.synchronize()
    sol = s.GetSolution()
    assert rel_err(sol.z.cpu().numpy(), gold["z"][:512]).max() <= 5e-9
    assert np.array_equal(sol.details.n_iter.cpu().numpy(), gold["n_iter"][:512])


def test_reduced_and_general_kernels_agree(walking_log):
    z0, it0, st0, info0, _ = solve(walking_log, "dense")
    for refine, tol in ((False, 1e-6), (True, 1e-8)):
        z1, it1, st1, info1, _ = solve(walking_log, "probe", refine=refine)
        assert info1["used"] and not info0["used"]
        assert rel_err(z1, z0).max() <= tol
        assert np.array_equal(it1, it0) and np.array_equal(st1, st0)


@pytest.mark.parametrize("name,rows", [("humanoid", 88), ("quadruped", 56), ("multicontact", 120)])
def test_synthetic_shapes_reduced(name, rows):
    from fcc_qp_b200 import synthetic as syn
    shp = syn.SHAPES[name]
    gold = np.load(os.path.join(G, f"synthetic_{name}_cold.npz"))
    qp = syn.make_batch(shp, gold["z"].shape[0])
    z, it, st, info, _ = solve(qp)
    assert info["used"] and info["rows"] == rows and info["deferred"] == 0, info
    assert rel_err(z, gold["z"]).max() <= 1e-6
    assert np.array_equal(it, gold["n_iter"]) and np.array_equal(st, gold["status"])


def test_host_arrays_take_the_reduced_kernel(walking_log):
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    z, it, st, info, _ = solve(walking_log, "auto", device_resident=False)
    assert info["used"] and info["rows"] == 72, info
    assert rel_err(z, gold["z"]).max() <= 1e-6 and np.array_equal(it, gold["n_iter"])


def test_qps_without_structure_are_handed_to_the_general_kernel(walking_log):
    """Every third QP gets a dense positive definite Q: the probe sees both kinds, the caps fit the structured
    ones only when given explicitly -- either way every QP must come out as the oracle says."""
    from oracle import Oracle
    qp = walking_log.take(np.arange(0, 600))
    rng = np.random.default_rng(5)
    dense_idx = np.arange(0, 600, 3)
    for i in dense_idx:
        Gm = rng.standard_normal((qp.n, qp.n)) * 0.05
        qp.Q[i] = qp.Q[i] + Gm @ Gm.T
    ref = Oracle("port").solve_batch(qp, warm_mode=0, **LOG_OPTS)
    # explicit caps sized for the structured QPs: the dense ones exceed them and are deferred
    z, it, st, info, _ = solve(qp, (23, 11, 6))
    assert info["used"] and info["deferred"] == len(dense_idx), info
    assert rel_err(z, ref["z"]).max() <= 1e-6
    assert (it != ref["n_iter"]).mean() <= 0.01
    # probe: the largest structure seen is "all dense" -> nothing to gain, general kernel for everyone
    z, it, st, info, _ = solve(qp, "probe")
    assert not info["used"]
    assert rel_err(z, ref["z"]).max() <= 1e-6


def test_caps_too_small_defers_everything(walking_log):
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    qp = walking_log.take(np.arange(0, 256))
    z, it, st, info, _ = solve(qp, (8, 8, 0))
    assert info["used"] and info["deferred"] == 256, info
    assert rel_err(z, gold["z"][:256]).max() <= 1e-6 and np.array_equal(it, gold["n_iter"][:256])


def test_warm_sequence_reduced(walking_log):
    """Warm-started solves (no pre-solve, rho-KKT factorization only) through the reduced kernel: lane-wise
    carried state, compared with the C restatement driven the same way."""
    from oracle import Oracle
    qp = walking_log.take(np.arange(300, 428))
    orc = Oracle("port")
    lanes = orc.lanes(qp.batch, qp.n, qp.m, qp.nc, qp.lambda_c_start)
    lanes.set_options(**LOG_OPTS)
    r0 = lanes.solve(qp, warm=False)
    qp2 = walking_log.take(np.arange(301, 429))
    r1 = lanes.solve(qp2, warm=True)
    z0, it0, _, info, s = solve(qp)
    assert info["used"]
    assert rel_err(z0, r0["z"]).max() <= 1e-6
    s.set_warm_start(True)
    import torch
    s.Solve(*[torch.as_tensor(a, device="cuda:0") for a in (qp2.Q, qp2.b, qp2.A_eq, qp2.b_eq, qp2.friction_coeffs, qp2.lb, qp2.ub)])
    torch.cuda.synchronize()
    sol = s.GetSolution()
    assert rel_err(sol.z.cpu().numpy(), r1["z"]).max() <= 1e-6
    assert (sol.details.n_iter.cpu().numpy() != r1["n_iter"]).mean() <= 0.02


def test_drop_in_object_uses_host_classification(walking_log):
    """The single-problem FCCQP object classifies its host data on the host (no probe launch)."""
    import fcc_qp as fq
    from fcc_qp_b200 import _native as nat
    gold = np.load(os.path.join(G, "walking_cold.npz"))
    s = fq.FCCQP(60, 38, 12, 38)
    o = fq.FCCQPOptions(); o.max_iter = 100; o.rho = 5e-5; o.eps_fcone = 1e-6; o.eps_bound = 1e-6
    s.set_options(o)
    for i in (0, 300, 1234):
        q = walking_log.qp(i)
        s.set_warm_start(False)
        s.Solve(q["Q"], q["b"], q["A_eq"], q["b_eq"], q["friction_coeffs"], q["lb"], q["ub"])
        sol = s.GetSolution()
        assert nat.last_struct_info()["used"]
        assert np.abs(sol.z - gold["z"][i]).max() / max(1.0, np.abs(gold["z"][i]).max()) <= 1e-6
        assert sol.details.n_iter == gold["n_iter"][i]


def test_schedule_from_previous_changes_nothing_but_the_order(walking_log):
    """FCCQP_SCHEDULE_LPT (lanes that ran long in the previous Solve are pulled from the work queue first): bit-identical
    results with and without it, on the reduced, the general and the warp kernels; and after a change of the batch the stale
    hint is still only a hint."""
    import torch
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    from fcc_qp_b200.synthetic import random_qps
    dev = torch.device("cuda:0")
    small = random_qps(np.random.default_rng(9), 4096, 12, 6, 6, 3)
    for qp, structure, opts in ((walking_log.tile(8192), "auto", LOG_OPTS), (walking_log.tile(8192), "dense", LOG_OPTS),
                                (small, "auto", dict(max_iter=50, rho=1e-3, eps_fcone=1e-6, eps_bound=1e-6))):
        args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
        out = []
        for hint in (False, True):
            s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(**opts))
            s.structure = structure
            s.schedule_from_previous = hint
            s.Solve(*args)                       # (no hint yet: the n_iter tensor is fresh)
            s.Solve(*args)                       # hinted by the first solve
            rolled = [torch.roll(a, 17, 0) if a.dim() > 1 or a.shape[0] == qp.batch else a for a in args]
            s.Solve(*rolled)                     # the hint now points at the wrong lanes
            sol = s.GetSolution()
            torch.cuda.synchronize()
            out.append((sol.z.clone(), sol.details.n_iter.clone(), sol.details.solve_status.clone()))
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
