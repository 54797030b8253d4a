"""Where does the reduced kernel's error sit? (by iteration count)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200 import _native as nat
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
G = os.path.join(ROOT, "tests", "golden")
qp = load_walking_log(); gold = np.load(os.path.join(G, "walking_cold.npz"))
for structure in ("probe", "dense"):
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6)); s.structure = structure
    s.Solve(*[torch.as_tensor(a, device="cuda:0") for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)])
    torch.cuda.synchronize()
    sol = s.GetSolution(); z = sol.z.cpu().numpy(); it = sol.details.n_iter.cpu().numpy()
    e = np.abs(z - gold["z"]).max(1) / np.maximum(1.0, np.abs(gold["z"]).max(1))
    for k in np.unique(it):
        sel = it == k
        print(structure, "n_iter", k, "count", sel.sum(), "max err %.3e" % e[sel].max(), "median %.3e" % np.median(e[sel]))
    w = np.argmax(e); print(" worst QP", w, "n_iter", it[w], "err", e[w], "var", np.argmax(np.abs(z[w] - gold["z"][w])))
