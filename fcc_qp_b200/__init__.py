"""fcc_qp_b200 -- B200-native batched ADMM solver for whole-body-control QPs.

Drop-in surface of the reference package ``fcc_qp`` (``src/main.cpp:19-56``):
``FCCQP``, ``FCCQPOptions``, ``FCCQPSolution``, ``FCCQPDetails`` come from the
pybind11 module ``fcc_qp_solver`` built over the CUDA C ABI; ``FCCQPBatch`` /
``solve_batch`` are the batched entry points.  Native pieces are imported
lazily so that pure-Python helpers (``logdata``, ``synthetic``) work without a
build; using a solver without the CUDA library raises ImportError.
"""
from __future__ import annotations

__all__ = ["FCCQP", "FCCQPOptions", "FCCQPSolution", "FCCQPDetails", "FCCQPBatch", "FCCQPBatchCpp", "BatchSolution",
           "solve_batch", "FCCQPError"]

_PYBIND = ("FCCQP", "FCCQPOptions", "FCCQPSolution", "FCCQPDetails", "FCCQPSolveStatus")
_BATCH = ("FCCQPBatch", "FCCQPBatchCpp", "BatchSolution", "BatchDetails", "FCCQPOptionsB", "solve_batch")


def __getattr__(name):
    if name in _PYBIND:
        from . import fcc_qp_solver  # built by fcc_qp_b200.build
        return getattr(fcc_qp_solver, name)
    if name in _BATCH:
        from . import batch
        return getattr(batch, name)
    if name == "FCCQPError":
        from ._native import FCCQPError
        return FCCQPError
    raise AttributeError(name)
