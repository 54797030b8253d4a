"""Generate golden outputs by running the UNMODIFIED reference (oracle/_ref).

Run in the build container only (needs /root/reference to build oracle/_ref):
    python tests/golden/make_goldens.py
Writes tests/golden/walking_{cold,warm}.npz (+ synthetic sets, see make_synthetic below).
Each file holds z, n_iter, status, eps_bounds, eps_friction_cone, bounds_viol,
friction_cone_viol for the solver settings stored alongside.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import Oracle  # noqa: E402
from fcc_qp_b200.logdata import load_walking_log  # noqa: E402

LOG_OPTS = dict(max_iter=100, rho=5e-5, eps_fcone=1e-6, eps_bound=1e-6)  # fcc_qp_test.py:78-83
KEYS = ("z", "n_iter", "status", "res_bounds", "res_fcone", "bounds_viol", "fcone_viol")


def save(name, r, opts, extra=None):
    out = {k: r[k] for k in KEYS}
    out["opts"] = np.array([opts["max_iter"], opts["rho"], opts["eps_fcone"], opts["eps_bound"]])
    out.update(extra or {})
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "n_iter hist:", dict(zip(*np.unique(r["n_iter"], return_counts=True))))


def main():
    ref = Oracle("ref")
    qp = load_walking_log()
    save("walking_cold.npz", ref.solve_batch(qp, warm_mode=0, **LOG_OPTS), LOG_OPTS)
    # fcc_qp_test.py:86-89: one solver object, set_warm_start(i > 0)
    save("walking_warm.npz", ref.solve_batch(qp, warm_mode=1, nthreads=1, **LOG_OPTS), LOG_OPTS)
    # paper settings (fccqp.pdf Table 1): eps 1e-4, max_iter 15
    paper = dict(max_iter=15, rho=5e-5, eps_fcone=1e-4, eps_bound=1e-4)
    save("walking_cold_paper.npz", ref.solve_batch(qp, warm_mode=0, **paper), paper)


def make_synthetic():
    """Goldens for the synthetic shapes.  Inputs are regenerated from
    the seeded generator at test time; `in_sum` guards against generator drift."""
    from fcc_qp_b200 import synthetic as syn
    ref = Oracle("ref")
    # (the B = 4096 humanoid set is BASELINE config 3 at a size where every code path of the kernel -- lazy
    # factorization, operator switch, deferral -- is hit hundreds of times; float32 storage of z would lose the bar,
    # so it stays float64: 2.8 MB)
    for shp, B, suffix in ((syn.HUMANOID, 192, ""), (syn.QUADRUPED, 192, ""), (syn.MULTICONTACT, 96, ""),
                           (syn.HUMANOID, 4096, "_4096")):
        qp = syn.make_batch(shp, B)
        chk = np.array([qp.Q.sum(), qp.A_eq.sum(), qp.b.sum(), qp.b_eq.sum(), qp.friction_coeffs.sum()])
        save(f"synthetic_{shp.name}_cold{suffix}.npz", ref.solve_batch(qp, warm_mode=0, nthreads=8, **LOG_OPTS),
             LOG_OPTS, dict(in_sum=chk))
    # config 5: multi-contact humanoid, T sequential warm-started batches with lane-wise state
    shp, B, T = syn.MULTICONTACT, 48, 32   # BASELINE config 5: T = 32 sequential warm-started batches
    qp = syn.make_batch(shp, B, seed=shp.seed + 1)
    rng = np.random.default_rng(shp.seed + 2)
    lanes = ref.lanes(B, qp.n, qp.m, qp.nc, qp.lambda_c_start)
    lanes.set_options(**LOG_OPTS)
    zs, its, sts = [], [], []
    for t in range(T):
        r = lanes.solve(qp, warm=t > 0)
        zs.append(r["z"]); its.append(r["n_iter"]); sts.append(r["status"])
        qp = syn.random_walk(qp, rng)
    np.savez_compressed(os.path.join(HERE, "synthetic_multicontact_warmseq.npz"), z=np.stack(zs),
                        n_iter=np.stack(its), status=np.stack(sts),
                        opts=np.array([LOG_OPTS["max_iter"], LOG_OPTS["rho"], LOG_OPTS["eps_fcone"], LOG_OPTS["eps_bound"]]))
    print("warmseq n_iter per step:", [int((a > 0).sum()) for a in its])


def make_kats():
    """Known-answer vectors of SURVEY section 4, taken from the compiled reference."""
    ref = Oracle("ref")
    out = {}
    # min 1/2 |x - f|^2  s.t. x in F(mu = 0.5): one contact, n = 3, no equality rows
    fs = np.array([[1, 0, 1], [3, 4, 1], [1, 0, -3], [0, 0, 2], [1, 1, 0]], dtype=np.float64)
    zs = []
    for f in fs:
        s = ref.solver(3, 0, 3, 0)
        s.set_options(2000, 1.0, 1e-10, 1e-10)
        s.Solve(np.eye(3), -f, np.zeros((0, 3)), np.zeros(0), [0.5], np.full(3, -np.inf), np.full(3, np.inf))
        zs.append(s.GetSolution()["z"])
    out["cone_f"], out["cone_z"] = fs, np.array(zs)
    # equality-only problem: closed-form KKT solution with n_iter = 0
    rng = np.random.default_rng(7)
    n, m = 12, 5
    G = rng.standard_normal((n, n)); Q = G @ G.T + np.eye(n); A = rng.standard_normal((m, n))
    b = rng.standard_normal(n); beq = rng.standard_normal(m)
    s = ref.solver(n, m, 0, 0)
    s.set_options(100, 1e-3, 1e-6, 1e-6)
    s.Solve(Q, b, A, beq, [], np.full(n, -np.inf), np.full(n, np.inf))
    r = s.GetSolution()
    out.update(eq_Q=Q, eq_A=A, eq_b=b, eq_beq=beq, eq_z=r["z"], eq_n_iter=np.array(r["n_iter"]))
    np.savez_compressed(os.path.join(HERE, "kats.npz"), **out)
    print("kats cone_z:\n", out["cone_z"], "\neq n_iter", r["n_iter"])


if __name__ == "__main__":
    main()
    make_synthetic()
    make_kats()
