"""Shared-structure batches (the sampling-based-MPC case): one Q and one A_eq for the
whole batch (batch stride 0), only b, b_eq, friction coefficients vary.  The library then runs two
launches with the KKT factorizations cached per CTA; results must match the general path (same
inputs materialised per QP) and the CPU oracle."""
import numpy as np
import pytest

from fcc_qp_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
OPTS = dict(max_iter=100, rho=5e-5, eps_fcone=1e-6, eps_bound=1e-6)


def shared_batch(shape, B, seed=7):
    """B QPs that share the structure terms of draw 0 (M, Jacobians, weights) and differ in commands / bias."""
    t = syn.make_terms(shape, B, seed=shape.seed + seed)
    rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[:1], a.shape))
    t.M, t.Jh, t.Jc, t.Jy, t.W = rep(t.M), rep(t.Jh), rep(t.Jc), rep(t.Jy), rep(t.W)
    return syn.assemble_numpy(t)


@pytest.mark.parametrize("name,B", [("quadruped", 3072), ("cassie_like", 3072), ("humanoid", 1536)])
def test_shared_structure_matches_general_path_and_oracle(name, B):
    import torch
    import oracle
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    from fcc_qp_b200 import _native as nat
    shp = syn.SHAPES[name]
    qp = shared_batch(shp, B)
    assert np.array_equal(qp.Q[0], qp.Q[-1]) and np.array_equal(qp.A_eq[0], qp.A_eq[-1])
    dev = torch.device("cuda:0")
    vec = [torch.as_tensor(a, device=dev) for a in (qp.b, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    Q1, A1 = torch.as_tensor(qp.Q[:1], device=dev), torch.as_tensor(qp.A_eq[:1], device=dev)

    def run(Q, A):
        s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start)
        s.set_options(FCCQPOptionsB(**OPTS))
        n0 = nat.lib().fccqp_kernel_launch_count()
        s.Solve(Q, vec[0], A, vec[1], vec[2], vec[3], vec[4])
        sol = s.GetSolution()
        torch.cuda.synchronize()
        return (sol.z.cpu().numpy(), sol.details.n_iter.cpu().numpy(), sol.details.solve_status.cpu().numpy(),
                nat.lib().fccqp_kernel_launch_count() - n0)

    zs, its, sts, launches_shared = run(Q1.expand(B, qp.n, qp.n), A1.expand(B, qp.m, qp.n))       # batch stride 0
    zg, itg, stg, launches_general = run(torch.as_tensor(qp.Q, device=dev), torch.as_tensor(qp.A_eq, device=dev))
    # the shared path really ran (2 launches); the materialised batch takes the structure-exploiting kernel
    # (probe + reduced kernel + general kernel over the handed-over list = 3) or the general kernel alone (1)
    assert launches_shared == 2 and launches_general in (1, 3)
    rel = lambda z, ref: (np.abs(z - ref).max(1) / np.maximum(1.0, np.abs(ref).max(1))).max()
    assert rel(zs, zg) <= 1e-7
    assert (its != itg).mean() <= 0.01
    sub = np.arange(0, B, 8)
    ref = oracle.Oracle("port").solve_batch(qp.take(sub), warm_mode=0, nthreads=8, **OPTS)
    assert rel(zs[sub], ref["z"]) <= 1e-6
    assert (its[sub] != ref["n_iter"]).mean() <= 0.02
    assert np.array_equal(sts[sub][its[sub] == ref["n_iter"]], ref["status"][its[sub] == ref["n_iter"]])


def test_shared_structure_host_path_and_2d_arguments():
    """Q [n,n] + A_eq [m,n] (numpy, host memory): staged once, stride 0 through the C ABI, chunked launches."""
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    shp = syn.QUADRUPED
    B = 8192
    qp = shared_batch(shp, B)
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(**OPTS))
    s.Solve(qp.Q[0], qp.b, qp.A_eq[0], qp.b_eq, qp.friction_coeffs, qp.lb[0], qp.ub[0])
    a = s.GetSolution()
    g = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); g.set_options(FCCQPOptionsB(**OPTS))
    g.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    r = g.GetSolution()
    err = (np.abs(a.z - r.z).max(1) / np.maximum(1.0, np.abs(r.z).max(1))).max()
    assert err <= 1e-7
    assert (a.details.n_iter != r.details.n_iter).mean() <= 0.01
    assert np.array_equal(a.details.solve_status[a.details.n_iter == r.details.n_iter],
                          r.details.solve_status[a.details.n_iter == r.details.n_iter])


def test_shared_structure_warm_sequence():
    """Warm-started re-solves of a shared-structure batch (carried x, mu_x, mu_lambda_c; b and b_eq take a
    random-walk step per tick): the cached-operator single launch against the general warm path."""
    import torch
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    shp = syn.QUADRUPED
    B, T = 3072, 3
    qp = shared_batch(shp, B)
    rng = np.random.default_rng(5)
    dev = torch.device("cuda:0")
    mk = lambda: FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start)
    a, g = mk(), mk()
    for s in (a, g):
        s.set_options(FCCQPOptionsB(**OPTS))
    Q1, A1 = torch.as_tensor(qp.Q[0], device=dev), torch.as_tensor(qp.A_eq[0], device=dev)
    Qf, Af = torch.as_tensor(qp.Q, device=dev), torch.as_tensor(qp.A_eq, device=dev)
    for t in range(T):
        vec = [torch.as_tensor(x, device=dev) for x in (qp.b, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
        for s, Q, A in ((a, Q1, A1), (g, Qf, Af)):
            s.set_warm_start(t > 0)
            s.Solve(Q, vec[0], A, vec[1], vec[2], vec[3], vec[4])
        za, zg = a.GetSolution().z.cpu().numpy(), g.GetSolution().z.cpu().numpy()
        ia, ig = a.GetSolution().details.n_iter.cpu().numpy(), g.GetSolution().details.n_iter.cpu().numpy()
        err = (np.abs(za - zg).max(1) / np.maximum(1.0, np.abs(zg).max(1)))
        same = ia == ig
        assert (~same).mean() <= 0.02, t
        assert err[same].max() <= 1e-6, t      # lanes whose iteration counts agree took the same path
        qp = syn.random_walk(qp, rng)
