"""Device-resident throughput of the other BASELINE.json shapes (configs 3-5; parity cases, not bench lines).
usage: python tools/bench_shapes.py [batch]    -> one JSON line per shape (cold and warm second solve)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200 import synthetic as syn
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
from fcc_qp_b200 import _native as nat
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
dev = torch.device("cuda:0")
def run(name, qp, B):
    qp = qp.tile(B)
    args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
    out = {"shape": name, "n": qp.n, "m": qp.m, "nc": qp.nc, "batch": B}
    for mode in ("cold", "warm"):
        s.set_warm_start(mode == "warm")
        best = 1e9
        for _ in range(3):
            if mode == "warm":      # warm solve of the same problems from the converged state of a cold solve
                s.set_warm_start(False); s.Solve(*args); s.set_warm_start(True)
            s.Solve(*args); torch.cuda.synchronize()
            best = min(best, s.GetSolution().details.device_time)
        it = s.GetSolution().details.n_iter.cpu().numpy()
        out[mode] = {"ms": 1e3 * best, "M_qps_per_s": B / best / 1e6, "iterating_fraction": float((it > 0).mean()),
                     "max_iter_fraction": float((it == 100).mean())}
    out["launch"] = nat.last_launch_info()
    print(json.dumps(out), flush=True)
run("cassie_walking_log", load_walking_log(), B)
for nm, gen in (("humanoid", 2048), ("quadruped", 2048), ("multicontact", 1024)):
    run(nm, syn.make_batch(syn.SHAPES[nm], gen), B if nm != "multicontact" else B // 4)
# shared-structure batches (one Q / A_eq for the whole batch): the two-launch cached-factor path
for nm in ("quadruped", "cassie_like"):
    shp = syn.SHAPES[nm]
    t = syn.make_terms(shp, 4096, seed=shp.seed + 7)
    rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[:1], a.shape))
    t.M, t.Jh, t.Jc, t.Jy, t.W = rep(t.M), rep(t.Jh), rep(t.Jc), rep(t.Jy), rep(t.W)
    q = syn.assemble_numpy(t).tile(B)
    vec = [torch.as_tensor(a, device=dev) for a in (q.b, q.b_eq, q.friction_coeffs, q.lb, q.ub)]
    Q1, A1 = torch.as_tensor(q.Q[:1], device=dev), torch.as_tensor(q.A_eq[:1], device=dev)
    s = FCCQPBatch(q.n, q.m, q.nc, q.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
    res = {"shape": nm + "_shared_structure", "n": q.n, "m": q.m, "batch": B}
    for label, Qx, Ax in (("shared", Q1.expand(B, q.n, q.n), A1.expand(B, q.m, q.n)),
                          ("general", torch.as_tensor(q.Q, device=dev), torch.as_tensor(q.A_eq, device=dev))):
        best = 1e9
        for _ in range(3):
            s.Solve(Qx, vec[0], Ax, vec[1], vec[2], vec[3], vec[4]); torch.cuda.synchronize()
            best = min(best, s.GetSolution().details.device_time)
        it = s.GetSolution().details.n_iter.cpu().numpy()
        res[label] = {"ms": 1e3 * best, "M_qps_per_s": B / best / 1e6, "iterating_fraction": float((it > 0).mean())}
    print(json.dumps(res), flush=True)
