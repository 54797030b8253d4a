"""Randomised small shapes -- odd n (no 16-byte row alignment: the 8-byte / 4-byte staging paths), n and m
that are not multiples of 8 (tile padding), m = 0, nc = 0, contacts anywhere in the variable vector --
against the CPU oracle (C restatement of the reference) on the same inputs, in FP64 and in the
float32-data mode, on host and device paths, plus the strided / column-major device layouts."""
import dataclasses

import numpy as np
import pytest

import oracle
from fcc_qp_b200.logdata import QPBatch

pytestmark = pytest.mark.gpu
OPTS = dict(max_iter=200, rho=1e-3, eps_fcone=1e-7, eps_bound=1e-7)


from fcc_qp_b200.synthetic import random_qps  # noqa: E402,F401  (the generator lives with the other synthetic sets)


SHAPES = [(5, 2, 3, 1), (7, 0, 3, 4), (9, 4, 0, 0), (13, 6, 6, 5), (17, 9, 3, 14), (24, 8, 6, 0), (31, 15, 9, 20),
          (40, 39, 12, 8), (57, 30, 6, 51), (64, 64, 12, 0)]


def rel(z, ref):
    return (np.abs(z - ref).max(1) / np.maximum(1.0, np.abs(ref).max(1))).max()


@pytest.mark.parametrize("n,m,nc,lcs", SHAPES)
def test_random_shape_fp64_and_fp32_data(n, m, nc, lcs):
    import torch
    from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
    rng = np.random.default_rng(1000 * n + 10 * m + nc)
    B = 24
    qp = random_qps(rng, B, n, m, nc, lcs)
    ref = oracle.Oracle("port").solve_batch(qp, warm_mode=0, nthreads=4, **OPTS)
    args = (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    # FP64, host path
    s = FCCQPBatch(n, m, nc, lcs); s.set_options(FCCQPOptionsB(**OPTS))
    s.Solve(*args)
    sol = s.GetSolution()
    assert rel(sol.z, ref["z"]) <= 1e-6
    assert (sol.details.n_iter != ref["n_iter"]).mean() <= 0.05
    same = sol.details.n_iter == ref["n_iter"]
    assert np.array_equal(sol.details.solve_status[same], ref["status"][same])
    # FP64, device path, column-major A and a Q slab cut out of a wider allocation (row stride != n)
    dev = torch.device("cuda:0")
    Qw = torch.zeros((B, n, n + 3), dtype=torch.float64, device=dev)
    Qw[:, :, :n] = torch.as_tensor(qp.Q, device=dev)
    At = torch.as_tensor(np.ascontiguousarray(qp.A_eq.transpose(0, 2, 1)), device=dev).transpose(1, 2)   # [B,m,n], column-major
    d = FCCQPBatch(n, m, nc, lcs); d.set_options(FCCQPOptionsB(**OPTS))
    d.Solve(Qw[:, :, :n], *[torch.as_tensor(a, device=dev) for a in (qp.b,)], At,
            *[torch.as_tensor(a, device=dev) for a in (qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)])
    zd = d.GetSolution().z.cpu().numpy()
    assert rel(zd, sol.z) <= 1e-9
    # float32 data: equals the oracle on float32-rounded inputs
    r32 = lambda a: np.ascontiguousarray(a.astype(np.float32).astype(np.float64))
    qp32 = dataclasses.replace(qp, Q=r32(qp.Q), b=r32(qp.b), A_eq=r32(qp.A_eq), b_eq=r32(qp.b_eq),
                               friction_coeffs=r32(qp.friction_coeffs), lb=r32(qp.lb), ub=r32(qp.ub))
    ref32 = oracle.Oracle("port").solve_batch(qp32, warm_mode=0, nthreads=4, **OPTS)
    for path in ("host", "device"):
        f = FCCQPBatch(n, m, nc, lcs, precision="fp32_data"); f.set_options(FCCQPOptionsB(**OPTS))
        a32 = args if path == "host" else [torch.as_tensor(a, dtype=torch.float32, device=dev) for a in args]
        f.Solve(*a32)
        z32 = f.GetSolution().z
        z32 = z32.cpu().numpy() if path == "device" else z32
        assert rel(z32, ref32["z"]) <= 1e-6, path
