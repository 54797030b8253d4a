"""Latency of small batched solves: eager launch against a replayed CUDA graph (host wall time per step, device synchronised).
usage: python tools/latency_run.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fcc_qp_b200.logdata import load_walking_log
from fcc_qp_b200.batch import FCCQPBatch, FCCQPOptionsB
log = load_walking_log()
dev = torch.device("cuda:0")
for B in (1, 16, 148, 740, 4096):
    qp = log.take(np.arange(B) % log.batch)
    args = [torch.as_tensor(a, device=dev) for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)]
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start); s.set_options(FCCQPOptionsB(100, 5e-5, 1e-6, 1e-6))
    s.time_kernel = False; s.zero_copy_outputs = True
    s.Solve(*args); s.set_warm_start(True)
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        for _ in range(3): s.Solve(*args)
    torch.cuda.synchronize()
    def timeit(fn, reps=200):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        return 1e6 * float(np.median(ts[20:]))
    with torch.cuda.stream(side):
        eager = timeit(lambda: s.Solve(*args))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        s.Solve(*args)
    graph = timeit(g.replay)
    print(json.dumps({"batch": B, "warm": True, "eager_us_p50": eager, "graph_replay_us_p50": graph}), flush=True)
