"""One process, ONE call, all GPUs: fccqp_batch_solve_multi on a pinned host batch of the tiled walking log.
usage: python tools/bench_multi_call.py [n_devices] [qps_per_device]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
W = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
per = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
log = load_walking_log()
B = W * per
qs = log.take(np.arange(B) % log.batch)
hp = [torch.from_numpy(a).pin_memory().numpy() for a in (qs.Q, qs.b, qs.A_eq, qs.b_eq, qs.friction_coeffs, qs.lb, qs.ub)]
del qs
s = FCCQPBatch(log.n, log.m, log.nc, log.lambda_c_start, device=list(range(W)))
s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6)); s.zero_copy_outputs = True
s.Solve(*hp)
ts = []
for _ in range(4):
    t0 = time.perf_counter(); s.Solve(*hp); z = float(s.GetSolution().z[:, 0].sum()); ts.append(time.perf_counter() - t0)
gold = np.load(os.path.join(ROOT, "tests", "golden", "walking_cold.npz"))
idx = np.arange(B) % log.batch
ok = bool(np.array_equal(s.GetSolution().details.n_iter, gold["n_iter"][idx]))
print(json.dumps({"devices": W, "batch": B, "ms_per_call_best": 1e3 * min(ts), "ms_per_call_median": 1e3 * float(np.median(ts)),
                  "qps": B / min(ts), "h2d_bytes_per_call": int(sum(a.nbytes for a in hp)), "iteration_counts_match_goldens": ok}))
