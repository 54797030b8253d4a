"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends for the two CPU oracles.

* ``Oracle("port")``  -> oracle/libfccqp_oracle.so, the C restatement
  (oracle/fccqp_oracle.c) of the reference algorithm.
* ``Oracle("ref")``   -> oracle/_ref/libfccqp_ref.so, the UNMODIFIED reference
  (``/root/reference/src``) compiled by oracle/Makefile behind oracle/ref_shim.cpp.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.  The product (``fcc_qp_b200``) never does.

Both libraries take COLUMN-major matrices (Eigen layout, ``src/fcc_qp.hpp:114``);
the helpers below accept the row-major ``[B, m, n]`` stacks used everywhere else
in this repository and transpose once, outside any timed region.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "libfccqp_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libfccqp_ref.so")
_REF_AVX2_SO = os.path.join(_HERE, "_ref", "libfccqp_ref_avx2.so")   # same sources, -march=x86-64-v3 (bench.py extra)
REFERENCE_ROOT = "/root/reference"

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(ref: bool = True, quiet: bool = True) -> None:
    """Compile the C restatement, and oracle/_ref when the reference tree exists."""
    kw = dict(stdout=subprocess.DEVNULL) if quiet else {}
    subprocess.check_call(["make", "-C", _HERE, "port"], **kw)
    if ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.check_call(["make", "-C", _HERE, "ref"], **kw)


def _so(kind: str) -> str:
    return {"port": _PORT_SO, "ref": _REF_SO, "ref_avx2": _REF_AVX2_SO}[kind]


def have(kind: str) -> bool:
    return os.path.exists(_so(kind))


def _ptr(a, t=_dp):
    return a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def colmajor_stack(M: np.ndarray) -> np.ndarray:
    """[B, r, c] row-major stack -> buffer whose i-th slab is the column-major r x c matrix."""
    return np.ascontiguousarray(np.swapaxes(M, -1, -2))


class Oracle:
    def __init__(self, kind: str = "port"):
        assert kind in ("port", "ref", "ref_avx2")
        self.kind = kind
        path = _so(kind)
        if not os.path.exists(path):
            build(ref=(kind != "port"))
        self.lib = C.CDLL(path)
        self.p = "fccqp_oracle_" if kind == "port" else "fccqp_ref_"
        f = lambda name: getattr(self.lib, self.p + name)
        f("create").restype = C.c_void_p
        f("create").argtypes = [C.c_int] * 4
        f("destroy").argtypes = [C.c_void_p]
        f("set_options").argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        f("set_warm_start").argtypes = [C.c_void_p, C.c_int]
        f("solve").argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp]
        if kind == "port":
            f("solve").restype = C.c_int
        f("get_solution").argtypes = [C.c_void_p, _dp, _ip, _ip, _dp]
        f("solve_batch").restype = C.c_double
        f("solve_batch").argtypes = ([C.c_int] * 6 + [C.c_double] * 3 + [C.c_int] * 2 +
                                     [_dp] * 5 + [C.c_long, _dp, _dp, C.c_long, _dp, _ip, _ip, _dp])
        f("solve_lanes").argtypes = ([C.POINTER(C.c_void_p), C.c_int, C.c_int] + [_dp] * 5 +
                                     [C.c_long, _dp, _dp, C.c_long, _dp, _ip, _ip, _dp])
        f("hardware_threads").restype = C.c_int
        if kind == "port":
            f("presolve_path").argtypes = [C.c_void_p]
            f("presolve_path").restype = C.c_int
            f("get_state").argtypes = [C.c_void_p, _dp, _dp, _dp]
            f("set_state").argtypes = [C.c_void_p, _dp, _dp, _dp]
            self.lib.fccqp_oracle_project_cone3.argtypes = [_dp, C.c_double, _dp]
            self.lib.fccqp_oracle_set_relaxation.argtypes = [C.c_double]
            self.lib.fccqp_oracle_set_trace.argtypes = [C.c_void_p, _dp, C.c_int]
            self.lib.fccqp_oracle_set_adaptive_rho.argtypes = [C.c_int]
            self.lib.fccqp_oracle_cone_violation.restype = C.c_double
            self.lib.fccqp_oracle_cone_violation.argtypes = [_dp, C.c_int, _dp]
            self.lib.fccqp_oracle_bound_violation.restype = C.c_double
            self.lib.fccqp_oracle_bound_violation.argtypes = [_dp, _dp, _dp, C.c_int]

    def fn(self, name):
        return getattr(self.lib, self.p + name)

    def set_relaxation(self, alpha: float) -> None:
        """Over-relaxation of the C restatement (port only; checks the product's opt-in extension).
        Process-wide; reset to 1.0 (the reference's iteration) when done."""
        assert self.kind == "port"
        self.lib.fccqp_oracle_set_relaxation(float(alpha))

    def set_adaptive_rho(self, interval: int) -> None:
        """Adaptive rho of the C restatement (port only; checks the product's opt-in extension).  Process-wide;
        reset to 0 (the reference's fixed rho) when done."""
        assert self.kind == "port"
        self.lib.fccqp_oracle_set_adaptive_rho(int(interval))

    def hardware_threads(self) -> int:
        return int(self.fn("hardware_threads")())

    # ---- single-solver object, mirrors the FCCQP class ----
    def solver(self, n, m, nc, lcs) -> "OracleSolver":
        return OracleSolver(self, n, m, nc, lcs)

    # ---- batch over stacked row-major arrays ----
    def solve_batch(self, qp, max_iter=1000, rho=1e-6, eps_fcone=1e-3, eps_bound=1e-6,
                    warm_mode=0, nthreads=1, prepared=None) -> dict:
        """Solve every QP of ``qp`` (a fcc_qp_b200.logdata.QPBatch).

        warm_mode 0: all cold.  warm_mode 1: warm-sequential inside each thread chunk
        (``fcc_qp_test.py:86-89`` when nthreads == 1).  Returns arrays plus ``elapsed``
        (seconds, slowest thread, Solve+GetSolution only).
        """
        B, n, m = qp.batch, qp.n, qp.m
        Qc, Ac = prepared if prepared is not None else (colmajor_stack(qp.Q), colmajor_stack(qp.A_eq))
        b, beq, mu, lb, ub = map(_f64, (qp.b, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub))
        z = np.empty((B, n)); it = np.empty(B, np.int32); st = np.empty(B, np.int32)
        det = np.empty((B, 6))
        el = self.fn("solve_batch")(B, n, m, qp.nc, qp.lambda_c_start, max_iter, rho, eps_fcone,
                                    eps_bound, warm_mode, nthreads, _ptr(Qc), _ptr(b), _ptr(Ac),
                                    _ptr(beq), _ptr(mu), mu.shape[1] if mu.ndim == 2 else 0,
                                    _ptr(lb), _ptr(ub), n, _ptr(z), _ptr(it, _ip), _ptr(st, _ip),
                                    _ptr(det))
        return dict(z=z, n_iter=it, status=st, res_bounds=det[:, 0].copy(),
                    res_fcone=det[:, 1].copy(), bounds_viol=det[:, 2].copy(),
                    fcone_viol=det[:, 3].copy(), solve_time=det[:, 4].copy(),
                    factorization_time=det[:, 5].copy(), elapsed=float(el))

    def lanes(self, B, n, m, nc, lcs) -> "OracleLanes":
        return OracleLanes(self, B, n, m, nc, lcs)


class OracleSolver:
    """One solver object with the reference's method names (src/fcc_qp.hpp:73-121)."""

    def __init__(self, orc: Oracle, n, m, nc, lcs):
        self.o, self.n, self.m, self.nc, self.lcs = orc, n, m, nc, lcs
        self.h = C.c_void_p(orc.fn("create")(n, m, nc, lcs))
        if not self.h:
            raise ValueError("invalid dimensions")

    def __del__(self):
        try:
            self.o.fn("destroy")(self.h)
        except Exception:
            pass

    def set_options(self, max_iter, rho, eps_fcone, eps_bound):
        self.o.fn("set_options")(self.h, int(max_iter), float(rho), float(eps_fcone), float(eps_bound))

    def set_warm_start(self, warm):
        self.o.fn("set_warm_start")(self.h, int(bool(warm)))

    def Solve(self, Q, b, A_eq, b_eq, friction_coeffs, lb, ub):
        Qc = np.asfortranarray(Q, dtype=np.float64)
        Ac = np.asfortranarray(np.asarray(A_eq, dtype=np.float64).reshape(self.m, self.n))
        mu = _f64(np.asarray(friction_coeffs, dtype=np.float64).reshape(-1))
        b, b_eq, lb, ub = map(_f64, (b, b_eq, lb, ub))
        if self.o.kind == "ref" and mu.size < self.nc // 3:
            raise IndexError("friction_coeffs too short")  # reference: std::out_of_range
        r = self.o.fn("solve")(self.h, _ptr(Qc), _ptr(b), _ptr(Ac), _ptr(b_eq), _ptr(mu), mu.size,
                               _ptr(lb), _ptr(ub))
        if self.o.kind == "port" and r != 0:
            raise IndexError("friction_coeffs too short")

    def GetSolution(self) -> dict:
        z = np.empty(self.n); it = C.c_int(); st = C.c_int(); det = np.empty(6)
        self.o.fn("get_solution")(self.h, _ptr(z), C.byref(it), C.byref(st), _ptr(det))
        return dict(z=z, n_iter=it.value, status=st.value, res_bounds=det[0], res_fcone=det[1],
                    bounds_viol=det[2], fcone_viol=det[3], solve_time=det[4],
                    factorization_time=det[5])

    def trace_residuals(self, cap: int) -> np.ndarray:
        """Port only: the following Solve calls record (bound residual, cone residual) per ADMM iteration
        -- the two numbers the exit test of fcc_qp.cpp:105 compares with eps -- into the returned [cap, 2] array."""
        assert self.o.kind == "port"
        self._trace = np.full((cap, 2), np.nan)
        self.o.lib.fccqp_oracle_set_trace(self.h, _ptr(self._trace), cap)
        return self._trace

    def presolve_path(self) -> int:
        return int(self.o.fn("presolve_path")(self.h))

    def get_state(self):
        x, mx, mc = np.empty(self.n), np.empty(self.n), np.empty(max(self.nc, 1))
        self.o.fn("get_state")(self.h, _ptr(x), _ptr(mx), _ptr(mc))
        return x, mx, mc[: self.nc]


class OracleLanes:
    """B persistent solver objects, one per lane."""

    def __init__(self, orc: Oracle, B, n, m, nc, lcs):
        self.o, self.B, self.n, self.m, self.nc = orc, B, n, m, nc
        self.handles = (C.c_void_p * B)(*[orc.fn("create")(n, m, nc, lcs) for _ in range(B)])

    def __del__(self):
        try:
            for h in self.handles:
                self.o.fn("destroy")(C.c_void_p(h))
        except Exception:
            pass

    def set_options(self, max_iter, rho, eps_fcone, eps_bound):
        for h in self.handles:
            self.o.fn("set_options")(C.c_void_p(h), int(max_iter), float(rho), float(eps_fcone),
                                     float(eps_bound))

    def solve(self, qp, warm: bool) -> dict:
        B, n = self.B, self.n
        Qc, Ac = colmajor_stack(qp.Q), colmajor_stack(qp.A_eq)
        b, beq, mu, lb, ub = map(_f64, (qp.b, qp.b_eq, qp.friction_coeffs, qp.lb, qp.ub))
        z = np.empty((B, n)); it = np.empty(B, np.int32); st = np.empty(B, np.int32)
        det = np.empty((B, 6))
        self.o.fn("solve_lanes")(self.handles, B, int(warm), _ptr(Qc), _ptr(b), _ptr(Ac), _ptr(beq),
                                 _ptr(mu), mu.shape[1], _ptr(lb), _ptr(ub), n, _ptr(z),
                                 _ptr(it, _ip), _ptr(st, _ip), _ptr(det))
        return dict(z=z, n_iter=it, status=st, res_bounds=det[:, 0].copy(),
                    res_fcone=det[:, 1].copy(), bounds_viol=det[:, 2].copy(),
                    fcone_viol=det[:, 3].copy())
