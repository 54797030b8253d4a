"""Effect of the opt-in adaptive rho (fccqp_options::adapt_rho_interval) on the bench workload and the other shapes.
usage: python tools/adapt_run.py [B]   (structure="dense" for the interval-0 rows as well: same kernel on both sides)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200 import synthetic as syn
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
sets = (("walking_log", load_walking_log(), B), ("humanoid", syn.make_batch(syn.HUMANOID, 2048), B),
        ("quadruped", syn.make_batch(syn.QUADRUPED, 2048), B), ("multicontact", syn.make_batch(syn.MULTICONTACT, 2048), B // 4))
for name, base, Bn in sets:
    qp = base.tile(Bn)
    args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    for K, structure in ((0, "auto"), (0, "dense"), (5, "auto"), (10, "auto")):
        s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.structure = structure
        s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6, 1.0, K))
        best = 1e9
        for _ in range(3):
            s.Solve(*args); torch.cuda.synchronize(); best = min(best, s.GetSolution().details.device_time)
        d = s.GetSolution().details
        it = d.n_iter.cpu().numpy()
        print(json.dumps({"workload": name, "batch": Bn, "adapt_rho_interval": K, "structure": structure, "ms": 1e3 * best,
                          "M_qps_per_s": Bn / best / 1e6, "max_iter_fraction": float((it == 100).mean()),
                          "mean_iterations_of_iterating": float(it[it > 0].mean()),
                          "friction_cone_viol_max": float(d.friction_cone_viol.max())}), flush=True)
    del args
