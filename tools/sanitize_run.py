"""Small cold + warm batched solves for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).
usage: compute-sanitizer --tool racecheck python tools/sanitize_run.py [shape]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200 import synthetic as syn
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
shape = sys.argv[1] if len(sys.argv) > 1 else "walking"
if shape == "odd":   # odd n, no 16-byte row alignment: the 8-byte / 4-byte staging paths
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_random_shapes import random_qps
    qp = random_qps(np.random.default_rng(3), 24, 13, 6, 6, 5)
else:
    qp = load_walking_log().take(np.arange(0, 2019, 48)) if shape == "walking" else syn.make_batch(syn.SHAPES[shape], 24)
ADAPT = int(os.environ.get("SAN_ADAPT", "0"))   # adaptive rho: general / warp kernels with refactorizations
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6, 1.0, ADAPT))
s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
it = s.GetSolution().details.n_iter
s.set_warm_start(True)
s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
print(shape, "QPs", qp.batch, "cold iterations", np.unique(it, return_counts=True), "warm", np.unique(s.GetSolution().details.n_iter, return_counts=True))
# shared-structure batch (two-launch path with cached factorizations)
if shape not in ("walking", "odd"):
    shp = syn.SHAPES[shape]
    t = syn.make_terms(shp, 1300 if shp.n + shp.m > 128 else 2500, seed=shp.seed + 7)
    rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[:1], a.shape))
    t.M, t.Jh, t.Jc, t.Jy, t.W = rep(t.M), rep(t.Jh), rep(t.Jc), rep(t.Jy), rep(t.W)
    q = syn.assemble_numpy(t)
    s2 = FCCQPBatch(q.n, q.m, q.nc, q.lambda_c_start); s2.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
    s2.Solve(q.Q[0], q.b, q.A_eq[0], q.b_eq, q.friction_coeffs, q.lb[0], q.ub[0])
    print("shared-structure", shape, "QPs", q.batch, np.unique(s2.GetSolution().details.n_iter, return_counts=True)[1][:3])
# float32 problem data (widening stage-in)
s3 = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, precision="fp32_data"); s3.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
s3.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
print("fp32_data", shape, "QPs", qp.batch, np.unique(s3.GetSolution().details.n_iter, return_counts=True)[1][:3])
# FP32 arithmetic (warp kernel, float instance) where it exists
if qp.n + qp.m <= 32:
    s4 = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, precision="fp32"); s4.set_options(FCCQPOptionsB(100, 5e-5, 1e-4, 1e-4))
    s4.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    s4.set_warm_start(True)
    s4.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    print("fp32", shape, "QPs", qp.batch, np.unique(s4.GetSolution().details.n_iter, return_counts=True)[1][:3])
# opt-in solution polish (device path: prepare kernel, inner eq-only solve, finish kernel)
if os.environ.get("SAN_POLISH"):
    import torch
    dev = torch.device("cuda:0")
    targs = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    s5 = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s5.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
    s5.Solve(*targs)
    pz = s5.Polish(); torch.cuda.synchronize()
    print("polish", shape, "QPs", qp.batch, "accepted", int(pz.details.polished.sum()))
