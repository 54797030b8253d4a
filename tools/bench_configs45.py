"""BASELINE.json configs 4 and 5 at their stated sizes on ONE GPU (per-GPU share of config 4).
config 4: synthetic quadruped, 2^20 QPs over 8 GPUs -> 2^17 per GPU, cold, device-resident.
config 5: multi-contact humanoid, T = 32 sequential warm-started batches of 2^14 (b, b_eq random walk 2 %)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200 import synthetic as syn
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
dev = torch.device("cuda:0")
opts = FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6)
# ---- config 4
shp = syn.QUADRUPED
B = 1 << 17
qp = syn.make_batch(shp, 8192).tile(B)
args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(opts)
best = 1e9
for _ in range(4):
    s.Solve(*args); torch.cuda.synchronize(); best = min(best, s.GetSolution().details.device_time)
it = s.GetSolution().details.n_iter.cpu().numpy()
print(json.dumps({"config": 4, "shape": "quadruped", "batch_per_gpu": B, "ms": 1e3 * best, "M_qps_per_s_per_gpu": B / best / 1e6,
                  "projected_8gpu_2^20_ms": 1e3 * best, "iterating_fraction": float((it > 0).mean())}), flush=True)
del args, s
# ---- config 5
shp = syn.MULTICONTACT
B, T = 1 << 14, 32
qp = syn.make_batch(shp, 2048, seed=shp.seed + 1).tile(B)
rng = np.random.default_rng(shp.seed + 2)
s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(opts)
Q, A = torch.as_tensor(qp.Q, device=dev), torch.as_tensor(qp.A_eq, device=dev)
fixed = [torch.as_tensor(a, device=dev) for a in (qp.friction_coeffs, qp.lb, qp.ub)]
b, beq = torch.as_tensor(qp.b, device=dev), torch.as_tensor(qp.b_eq, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(3)
tot, per = 0.0, []
for t in range(T):
    s.set_warm_start(t > 0)
    s.Solve(Q, b, A, beq, *fixed); torch.cuda.synchronize()
    dt = s.GetSolution().details.device_time; tot += dt; per.append(dt)
    b = b * (1.0 + 0.02 * torch.randn(b.shape, device=dev, dtype=torch.float64, generator=gen))
    beq = beq * (1.0 + 0.02 * torch.randn(beq.shape, device=dev, dtype=torch.float64, generator=gen))
it = s.GetSolution().details.n_iter.cpu().numpy()
print(json.dumps({"config": 5, "shape": "multicontact", "batch": B, "steps": T, "total_ms": 1e3 * tot,
                  "M_qps_per_s": B * T / tot / 1e6, "first_cold_ms": 1e3 * per[0], "warm_ms_median": 1e3 * float(np.median(per[1:])),
                  "iterating_fraction_last": float((it > 0).mean())}), flush=True)
