"""Compact, pickle-free storage of logged WBC QPs and a loader that stacks them
into dense ``[B, ...]`` arrays for the batched solver.

The reference ships its only workload as a 99 MB pickled object array of 2019
dicts (``/root/reference/test_data/id_qp_log_walking.npz``, loaded at
``fcc_qp_test.py:22-24``).  Q and A_eq have a fixed sparsity pattern over the
whole log and lb/ub/friction_coeffs are constant, so the same information is
stored here as the union pattern plus one row of values per QP (bit-exact,
~9 MB compressed, no ``allow_pickle``).
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np


@dataclasses.dataclass
class QPBatch:
    """Stacked problem data for B QPs of identical dimensions (C-contiguous f64).

    Field meaning follows ``FCCQP::Solve`` (reference ``src/fcc_qp.hpp:114-117``):
    ``Q[B,n,n]`` symmetric PSD, ``b[B,n]``, ``A_eq[B,m,n]``, ``b_eq[B,m]``,
    ``friction_coeffs[B,nc/3]``, ``lb[B,n]``, ``ub[B,n]``.
    """

    n: int
    m: int
    nc: int
    lambda_c_start: int
    Q: np.ndarray
    b: np.ndarray
    A_eq: np.ndarray
    b_eq: np.ndarray
    friction_coeffs: np.ndarray
    lb: np.ndarray
    ub: np.ndarray

    @property
    def batch(self) -> int:
        return int(self.Q.shape[0])

    def take(self, idx) -> "QPBatch":
        idx = np.asarray(idx)
        f = lambda a: np.ascontiguousarray(a[idx])
        return QPBatch(self.n, self.m, self.nc, self.lambda_c_start, f(self.Q), f(self.b),
                       f(self.A_eq), f(self.b_eq), f(self.friction_coeffs), f(self.lb), f(self.ub))

    def tile(self, batch: int) -> "QPBatch":
        """Repeat the QPs cyclically up to ``batch`` entries."""
        idx = np.arange(batch) % self.batch
        return self.take(idx)

    def objective(self, z: np.ndarray) -> np.ndarray:
        """0.5 z'Qz + b'z per QP (the reference never evaluates it; parity metric)."""
        return 0.5 * np.einsum("bi,bij,bj->b", z, self.Q, z) + np.einsum("bi,bi->b", self.b, z)

    def qp(self, i: int) -> dict:
        """One QP as the dict layout of the reference log (``fcc_qp_test.py:27-30``)."""
        return dict(Q=self.Q[i], b=self.b[i], A_eq=self.A_eq[i], b_eq=self.b_eq[i],
                    friction_coeffs=tuple(float(x) for x in self.friction_coeffs[i]),
                    lb=self.lb[i], ub=self.ub[i])


def stack_reference_log(path: str, nc: int = 12, lambda_c_start: int = 38) -> QPBatch:
    """Load the reference's pickled log (object array of dicts) and stack it."""
    qps = np.load(path, allow_pickle=True)["qps"]
    st = lambda k: np.ascontiguousarray(np.stack([np.asarray(q[k], dtype=np.float64) for q in qps]))
    Q, A = st("Q"), st("A_eq")
    return QPBatch(Q.shape[1], A.shape[1], nc, lambda_c_start, Q, st("b"), A, st("b_eq"),
                   st("friction_coeffs"), st("lb"), st("ub"))


def save_compact(batch: QPBatch, path: str) -> None:
    """Write ``batch`` as union-sparsity values (bit-exact round trip)."""
    n = batch.n
    iu = np.triu_indices(n)
    if not np.array_equal(batch.Q, batch.Q.transpose(0, 2, 1)):
        raise ValueError("compact format requires exactly symmetric Q")
    qmask = (batch.Q != 0).any(0)[iu]
    q_rows, q_cols = iu[0][qmask].astype(np.int16), iu[1][qmask].astype(np.int16)
    amask = (batch.A_eq != 0).any(0)
    a_rows, a_cols = [x.astype(np.int16) for x in np.nonzero(amask)]
    const = lambda a: bool((a == a[:1]).all())
    out = dict(
        dims=np.array([batch.n, batch.m, batch.nc, batch.lambda_c_start, batch.batch], np.int64),
        q_rows=q_rows, q_cols=q_cols, q_vals=batch.Q[:, q_rows, q_cols],
        a_rows=a_rows, a_cols=a_cols, a_vals=batch.A_eq[:, a_rows, a_cols],
        b=batch.b, b_eq=batch.b_eq,
    )
    for k in ("friction_coeffs", "lb", "ub"):
        a = getattr(batch, k)
        out[k] = a[:1] if const(a) else a
    np.savez_compressed(path, **out)


def load_compact(path: str) -> QPBatch:
    d = np.load(path, allow_pickle=False)
    n, m, nc, lcs, B = (int(x) for x in d["dims"])
    Q = np.zeros((B, n, n))
    qr, qc = d["q_rows"].astype(np.int64), d["q_cols"].astype(np.int64)
    Q[:, qr, qc] = d["q_vals"]
    Q[:, qc, qr] = d["q_vals"]
    A = np.zeros((B, m, n))
    A[:, d["a_rows"].astype(np.int64), d["a_cols"].astype(np.int64)] = d["a_vals"]
    bc = lambda a: np.ascontiguousarray(np.broadcast_to(a, (B,) + a.shape[1:]))
    return QPBatch(n, m, nc, lcs, Q, np.ascontiguousarray(d["b"]), A,
                   np.ascontiguousarray(d["b_eq"]), bc(d["friction_coeffs"]), bc(d["lb"]), bc(d["ub"]))


_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_walking_log() -> QPBatch:
    """The 2019 logged Cassie OSC QPs (n=60, m=38, nc=12, lambda_c_start=38)."""
    return load_compact(os.path.join(_GOLDEN, "walking_log_compact.npz"))
