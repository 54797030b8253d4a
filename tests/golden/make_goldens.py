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


if __name__ == "__main__":
    main()
