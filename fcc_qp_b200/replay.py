#!/usr/bin/env python
"""Headless replay of a logged WBC QP sequence -- the reference's ``fcc_qp_test.py`` without matplotlib.

The reference's only caller (``fcc_qp_test.py:72-91``) loads the pickled walking log, solves it QP
by QP through one ``FCCQP`` object with warm start after the first solve, and *plots* the torques,
solve times, iteration counts and violations.  This module runs the same loop (``--mode
sequential``, through the drop-in ``FCCQP`` object) or the whole log as one cold batch (``--mode
batch``, through ``FCCQPBatch``) and emits the numbers behind those plots as
one JSON document instead (``fcc_qp_test.py:44-69``: u, solve time, iterations, violations, the
v-dot / lambda_h / lambda_c slices)::

    python -m fcc_qp_b200.replay                          # shipped compact walking log, sequential warm
    python -m fcc_qp_b200.replay --mode batch --json out.json --save-solutions z.npz
    python -m fcc_qp_b200.replay --log /path/id_qp_log_walking.npz --convert compact.npz

Both modes run on the GPU (there is no CPU solve path); ``--convert`` and ``--describe`` only touch
the log file.  Solver settings default to the reference script's (rho 5e-5, eps 1e-6, max_iter 100,
``fcc_qp_test.py:78-83``).
"""
from __future__ import annotations

import argparse
import json
import sys
import time

import numpy as np

from .logdata import QPBatch, load_compact, load_walking_log, save_compact, stack_reference_log

# Cassie OSC variable layout used by the reference's plots (fcc_qp_test.py:52-56)
CASSIE_SLICES = {"vdot": (0, 22), "u": (22, 32), "lambda_h": (32, 38), "lambda_c": (38, 50)}


def load_log(path: str | None, nc: int, lambda_c_start: int) -> QPBatch:
    """Compact format (``logdata.save_compact``) or the reference's pickled object array."""
    if path is None:
        return load_walking_log()
    with np.load(path, allow_pickle=False) as d:
        compact = "dims" in d.files
    if compact:
        return load_compact(path)
    return stack_reference_log(path, nc=nc, lambda_c_start=lambda_c_start)  # needs allow_pickle (trusted file)


def describe(qp: QPBatch) -> dict:
    finite = np.isfinite(qp.lb) | np.isfinite(qp.ub)
    return {"qps": qp.batch, "num_vars": qp.n, "num_equality_constraints": qp.m, "nc": qp.nc,
            "lambda_c_start": qp.lambda_c_start, "bounded_variables": int(finite.any(0).sum()),
            "friction_coeffs_range": [float(qp.friction_coeffs.min()), float(qp.friction_coeffs.max())]
            if qp.nc else None,
            "bytes_dense_fp64": int(sum(a.nbytes for a in (qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs,
                                                            qp.lb, qp.ub)))}


def _percentiles(x) -> dict:
    x = np.asarray(x, dtype=np.float64)
    if x.size == 0:
        return {}
    return {"p50": float(np.percentile(x, 50)), "p90": float(np.percentile(x, 90)),
            "p99": float(np.percentile(x, 99)), "max": float(x.max()), "mean": float(x.mean())}


def summarize(qp: QPBatch, z, n_iter, status, res_b, res_f, bviol, fviol, max_iter: int,
              solve_times=None) -> dict:
    """The quantities ``make_plots`` draws (fcc_qp_test.py:44-69), as numbers."""
    z = np.asarray(z)
    n_iter = np.asarray(n_iter)
    vals, cnts = np.unique(n_iter, return_counts=True)
    out = {
        "qps": int(z.shape[0]),
        "iterations": {"histogram": {str(int(v)): int(c) for v, c in zip(vals, cnts)},
                       "mean_executed": float(np.where(n_iter == max_iter, max_iter, n_iter + 1).mean()),
                       "hit_max_iter": int((n_iter == max_iter).sum())},
        "status_counts": {str(int(v)): int(c) for v, c in zip(*np.unique(np.asarray(status), return_counts=True))},
        "admm_residual_bounds_max": float(np.max(res_b)), "admm_residual_friction_cone_max": float(np.max(res_f)),
        "bounds_viol": _percentiles(bviol), "friction_cone_viol": _percentiles(fviol),
        "objective": _percentiles(qp.objective(z)),
        "equality_residual_inf": float(np.abs(np.einsum("bij,bj->bi", qp.A_eq, z) - qp.b_eq).max()) if qp.m else 0.0,
        "z_abs_max": float(np.abs(z).max()),
    }
    if qp.n == 60 and qp.m == 38 and qp.nc == 12:   # the Cassie layout of the reference's plots
        out["slices"] = {k: {"min": float(z[:, a:b].min()), "max": float(z[:, a:b].max()),
                             "rms": float(np.sqrt((z[:, a:b] ** 2).mean()))} for k, (a, b) in CASSIE_SLICES.items()}
    if solve_times is not None:
        out["solve_time_s"] = _percentiles(solve_times)
    return out


def replay_sequential(qp: QPBatch, opts: dict, warm: bool = True, limit: int | None = None):
    """fcc_qp_test.py:77-89: one FCCQP object, ``set_warm_start(i > 0)``, Solve + GetSolution per QP."""
    from . import FCCQP, FCCQPOptions
    solver = FCCQP(qp.n, qp.m, qp.nc, qp.lambda_c_start)
    o = FCCQPOptions()
    o.rho, o.eps_fcone, o.eps_bound, o.max_iter = opts["rho"], opts["eps_fcone"], opts["eps_bound"], opts["max_iter"]
    solver.set_options(o)
    B = qp.batch if limit is None else min(limit, qp.batch)
    z = np.empty((B, qp.n))
    cols = {k: np.empty(B) for k in ("res_b", "res_f", "bviol", "fviol", "t", "tf")}
    n_iter, status = np.empty(B, np.int64), np.empty(B, np.int64)
    t0 = time.perf_counter()
    for i in range(B):
        solver.set_warm_start(warm and i > 0)
        q = qp.qp(i)
        solver.Solve(q["Q"], q["b"], q["A_eq"], q["b_eq"], q["friction_coeffs"], q["lb"], q["ub"])
        r = solver.GetSolution()
        d = r.details
        z[i] = r.z
        n_iter[i], status[i] = d.n_iter, getattr(d, "solve_status", 0)
        cols["res_b"][i], cols["res_f"][i] = d.eps_bounds, d.eps_friction_cone
        cols["bviol"][i], cols["fviol"][i] = d.bounds_viol, d.friction_cone_viol
        cols["t"][i], cols["tf"][i] = d.solve_time, d.factorization_time
    wall = time.perf_counter() - t0
    return dict(z=z, n_iter=n_iter, status=status, wall=wall, **cols)


def replay_batch(qp: QPBatch, opts: dict, device: int = 0, limit: int | None = None):
    """The whole log as ONE cold batched Solve (every QP solved like ``set_warm_start(False)``)."""
    from .batch import FCCQPBatch, FCCQPOptionsB
    if limit is not None:
        qp = qp.take(np.arange(min(limit, qp.batch)))
    s = FCCQPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, device=device)
    s.set_options(FCCQPOptionsB(**opts))
    t0 = time.perf_counter()
    s.Solve(qp.Q, qp.b, qp.A_eq, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub)
    sol = s.GetSolution()
    wall = time.perf_counter() - t0
    d = sol.details
    return dict(z=np.asarray(sol.z), n_iter=np.asarray(d.n_iter), status=np.asarray(d.solve_status),
                res_b=np.asarray(d.eps_bounds), res_f=np.asarray(d.eps_friction_cone),
                bviol=np.asarray(d.bounds_viol), fviol=np.asarray(d.friction_cone_viol), wall=wall,
                device_time=float(d.device_time))


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="python -m fcc_qp_b200.replay", description=__doc__.split("\n\n")[0])
    ap.add_argument("--log", default=None, help="log file: compact .npz or the reference's pickled id_qp_log_*.npz "
                                                 "(default: the shipped compact walking log)")
    ap.add_argument("--mode", choices=["sequential", "batch"], default="sequential")
    ap.add_argument("--cold", action="store_true", help="sequential mode: never warm start")
    ap.add_argument("--limit", type=int, default=None, help="only the first N QPs")
    ap.add_argument("--rho", type=float, default=5e-5)
    ap.add_argument("--eps-fcone", type=float, default=1e-6)
    ap.add_argument("--eps-bound", type=float, default=1e-6)
    ap.add_argument("--max-iter", type=int, default=100)
    ap.add_argument("--nc", type=int, default=12, help="contact force variables (reference-format logs only)")
    ap.add_argument("--lambda-c-start", type=int, default=38, help="first contact variable (reference-format logs only)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--json", default=None, help="write the summary here instead of stdout")
    ap.add_argument("--save-solutions", default=None, help="write z / n_iter / violations to this .npz")
    ap.add_argument("--convert", default=None, help="write the log in the compact format to this path and exit")
    ap.add_argument("--describe", action="store_true", help="print the log's dimensions and exit")
    return ap


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    qp = load_log(args.log, args.nc, args.lambda_c_start)
    if args.convert:
        save_compact(qp, args.convert)
        print(json.dumps({"converted": args.convert, **describe(qp)}))
        return 0
    if args.describe:
        print(json.dumps(describe(qp)))
        return 0
    opts = dict(max_iter=args.max_iter, rho=args.rho, eps_fcone=args.eps_fcone, eps_bound=args.eps_bound)
    if args.mode == "sequential":
        r = replay_sequential(qp, opts, warm=not args.cold, limit=args.limit)
        sub = qp if args.limit is None else qp.take(np.arange(r["z"].shape[0]))
        summary = summarize(sub, r["z"], r["n_iter"], r["status"], r["res_b"], r["res_f"], r["bviol"], r["fviol"],
                            args.max_iter, solve_times=r["t"])
        summary["factorization_time_s"] = _percentiles(r["tf"])
    else:
        r = replay_batch(qp, opts, device=args.device, limit=args.limit)
        sub = qp if args.limit is None else qp.take(np.arange(r["z"].shape[0]))
        summary = summarize(sub, r["z"], r["n_iter"], r["status"], r["res_b"], r["res_f"], r["bviol"], r["fviol"],
                            args.max_iter)
        summary["device_time_s"] = r["device_time"]
    summary.update({"mode": args.mode, "warm_start": args.mode == "sequential" and not args.cold, "solver": opts,
                    "wall_s": r["wall"], "qps_per_s_wall": r["z"].shape[0] / r["wall"], "log": describe(sub)})
    if args.save_solutions:
        np.savez_compressed(args.save_solutions, z=r["z"], n_iter=r["n_iter"], status=r["status"],
                            res_bounds=r["res_b"], res_fcone=r["res_f"], bounds_viol=r["bviol"], fcone_viol=r["fviol"])
    text = json.dumps(summary, indent=1)
    if args.json:
        with open(args.json, "w") as f:
            f.write(text + "\n")
    else:
        print(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
