"""On-device WBC assembly (include/fccqp.h: fccqp_wbc_assemble) against its CPU
restatement synthetic.assemble_numpy, and the assembled QPs through the solver against the goldens
of the compiled reference.  The reference has no assembly step of its own (its callers hand Solve dense matrices), so
the assembly oracle is builder-authored: a pass here says the kernel matches OUR numpy statement of fccqp.pdf section 4,
not the reference; only the solver half (assembled QPs -> goldens) is parity with the reference."""
import os

import numpy as np
import pytest

from fcc_qp_b200 import synthetic as syn

G = os.path.join(os.path.dirname(__file__), "golden")
OPTS = dict(max_iter=100, rho=5e-5, eps_fcone=1e-6, eps_bound=1e-6)


def test_numpy_assembly_structure():
    """The CPU restatement itself: block structure of fccqp.pdf eq. 10 on the multi-contact shape."""
    t = syn.make_terms(syn.MULTICONTACT, 5)
    qp = syn.assemble_numpy(t)
    s = syn.MULTICONTACT
    nv, nu, nh, nc = s.nv, s.nu, s.nh, s.nc
    o_u, o_h, o_c, o_e = nv, nv + nu, nv + nu + nh, nv + nu + nh + nc
    assert qp.n == s.n and qp.m == s.m and qp.lambda_c_start == o_c
    assert np.array_equal(qp.Q, qp.Q.transpose(0, 2, 1))
    assert np.array_equal(qp.A_eq[:, :nv, :nv], t.M)
    assert np.array_equal(qp.A_eq[:, :nv, o_c:o_e], -t.Jc.transpose(0, 2, 1))
    assert np.array_equal(qp.A_eq[:, nv + nh:, :nv], t.Jc)
    assert np.array_equal(qp.A_eq[:, nv:nv + nh, :nv], t.Jh)
    assert np.array_equal(qp.A_eq[0, nv + nh:, o_e:], np.eye(nc))
    assert np.array_equal(qp.A_eq[0, :nv, o_u:o_h], -np.vstack([np.zeros((nv - nu, nu)), np.eye(nu)]))
    assert np.all(np.diagonal(qp.Q[:, o_h:o_c, o_h:o_c], axis1=1, axis2=2) == 0.0)   # zero-cost lambda_h
    assert np.array_equal(qp.b_eq, np.concatenate([-t.bias, -t.gamma_h, -t.gamma_c], 1))
    assert t.nbytes() < 0.3 * sum(a.nbytes for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq))   # what PCIe saves


def test_make_batch_is_terms_plus_assembly():
    for shp in (syn.HUMANOID, syn.QUADRUPED):
        a = syn.make_batch(shp, 7)
        b = syn.assemble_numpy(syn.make_terms(shp, 7))
        for k in ("Q", "b", "A_eq", "b_eq", "friction_coeffs", "lb", "ub"):
            assert np.array_equal(getattr(a, k), getattr(b, k))


@pytest.mark.gpu
@pytest.mark.parametrize("name,B", [("humanoid", 192), ("quadruped", 192), ("multicontact", 96), ("cassie_like", 64)])
def test_device_assembly_matches_numpy(name, B):
    import torch
    from fcc_qp_b200 import wbc
    t = syn.make_terms(syn.SHAPES[name], B)
    ref = syn.assemble_numpy(t)
    Q, b, A, beq, mu, lb, ub = wbc.assemble(t)
    torch.cuda.synchronize()
    Q, b, A, beq = (x.cpu().numpy() for x in (Q, b, A, beq))
    assert np.array_equal(A, ref.A_eq) and np.array_equal(beq, ref.b_eq)        # copies and signs: bit-exact
    assert np.array_equal(Q, Q.transpose(0, 2, 1))
    scale = np.abs(ref.Q).max()
    assert np.abs(Q - ref.Q).max() <= 1e-13 * scale                             # Jy' W Jy: summation order only
    assert np.abs(b - ref.b).max() <= 1e-13 * max(1.0, np.abs(ref.b).max())
    assert np.array_equal((Q != 0), (ref.Q != 0))                               # same sparsity pattern
    assert np.array_equal(lb.cpu().numpy(), ref.lb[0]) and np.array_equal(ub.cpu().numpy(), ref.ub[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name,B", [("humanoid", 192), ("multicontact", 96)])
def test_assembled_qps_solve_to_the_goldens(name, B):
    """terms -> device assembly -> batched solve, against the compiled reference run on the numpy-assembled QPs."""
    import torch
    from fcc_qp_b200 import wbc
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    gold = np.load(os.path.join(G, f"synthetic_{name}_cold.npz"))
    shp = syn.SHAPES[name]
    Q, b, A, beq, mu, lb, ub = wbc.assemble(syn.make_terms(shp, B))
    s = FCCQPBatch(shp.n, shp.m, shp.nc, shp.lambda_c_start)
    s.set_options(FCCQPOptionsB(**OPTS))
    s.Solve(Q, b, A, beq, mu, lb, ub)
    sol = s.GetSolution()
    z = sol.z.cpu().numpy()
    err = np.abs(z - gold["z"]).max(1) / np.maximum(1.0, np.abs(gold["z"]).max(1))
    assert err.max() <= 1e-6
    assert (sol.details.n_iter.cpu().numpy() != gold["n_iter"]).mean() <= 0.02


def test_assembly_needs_the_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from fcc_qp_b200 import wbc
    with pytest.raises(Exception):
        wbc.assemble(syn.make_terms(syn.QUADRUPED, 2))
