"""On-device assembly of whole-body-control QPs from robot quantities.

The reference's callers hand ``FCCQP::Solve`` finished ``Q, b, A_eq, b_eq``
(``src/fcc_qp.hpp:114-117``); in an operational-space controller those are assembled from the mass
matrix, Jacobians, bias forces and task commands of the current state (``fccqp.pdf`` section 4).
For a batch that lives on the GPU it is cheaper to ship those terms (``nv^2 + (nh+nc+ny) nv``
doubles per QP) and assemble there than to ship ``n^2 + m n`` doubles of mostly structural zeros:
``assemble`` calls ``fccqp_wbc_assemble`` (``include/fccqp.h``) and returns torch CUDA tensors that
``FCCQPBatch.Solve`` consumes in place.  ``synthetic.assemble_numpy`` is the CPU restatement the
tests compare against.  No CPU path: without the CUDA library / a device this raises.
"""
from __future__ import annotations

import ctypes as C

from . import _native as nat
from .synthetic import WBCTerms

_FIELDS = ("M", "Jh", "Jc", "Jy", "W", "ydd_cmd", "bias", "gamma_h", "gamma_c")


def to_device(terms: WBCTerms, device, pinned: dict | None = None) -> dict:
    """Host terms -> CUDA tensors (asynchronous from page-locked staging when ``pinned`` is given)."""
    import torch
    out = {}
    for k in _FIELDS + ("friction_coeffs",):
        a = getattr(terms, k)
        if pinned is not None:
            h = pinned.get(k)
            if h is None or tuple(h.shape) != a.shape:
                h = pinned[k] = torch.empty(a.shape, dtype=torch.float64).pin_memory()
            h.numpy()[...] = a
            out[k] = h.to(device, non_blocking=True)
        else:
            out[k] = torch.as_tensor(a, dtype=torch.float64, device=device)
    return out


def assemble(terms, shape=None, device=0, weights=None, u_max=None):
    """``(Q[B,n,n], b[B,n], A_eq[B,m,n], b_eq[B,m], friction_coeffs, lb[n], ub[n])`` as CUDA tensors.

    ``terms``: a ``WBCTerms`` (numpy, copied to the device) or a dict of CUDA tensors with the keys
    ``M, Jh, Jc, Jy, W, ydd_cmd, bias, gamma_h, gamma_c, friction_coeffs`` (then ``shape`` is required).
    """
    import torch
    dev = torch.device("cuda", device) if isinstance(device, int) else device
    if isinstance(terms, WBCTerms):
        shape = terms.shape
        weights = terms.weights if weights is None else weights
        u_max = terms.u_max if u_max is None else u_max
        t = to_device(terms, dev)
    else:
        t = terms
        if shape is None:
            raise ValueError("shape is required with tensor inputs")
    weights = (1e-5, 1e-4, 1e-6, 80.0) if weights is None else weights
    u_max = 300.0 if u_max is None else u_max
    nv, nu, nh, nc = shape.nv, shape.nu, shape.nh, shape.nc
    B, ny = int(t["M"].shape[0]), int(t["Jy"].shape[1])
    n, m = shape.n, shape.m
    for k in _FIELDS:
        if not t[k].is_cuda or t[k].dtype != torch.float64:
            raise ValueError(f"{k} must be a float64 CUDA tensor")
        t[k] = t[k].contiguous()
    Q = torch.empty((B, n, n), dtype=torch.float64, device=dev)
    b = torch.empty((B, n), dtype=torch.float64, device=dev)
    A = torch.empty((B, m, n), dtype=torch.float64, device=dev)
    beq = torch.empty((B, m), dtype=torch.float64, device=dev)
    d = nat.WbcDesc()
    d.abi_version, d.batch = nat.ABI_VERSION, B
    d.nv, d.nu, d.nh, d.nc, d.ny = nv, nu, nh, nc, ny
    d.device = dev.index if dev.index is not None else torch.cuda.current_device()
    d.w_vdot, d.w_u, d.w_lambda_c, d.w_eps = (float(w) for w in weights)
    bs = lambda a: int(a.stride(0)) if a.shape[0] > 1 else int(a[0].numel())
    for name, key, stride in (("M", "M", "M_batch_stride"), ("Jh", "Jh", "Jh_batch_stride"), ("Jc", "Jc", "Jc_batch_stride"),
                              ("Jy", "Jy", "Jy_batch_stride"), ("W", "W", "W_batch_stride"),
                              ("ydd_cmd", "ydd_cmd", "ydd_batch_stride"), ("bias", "bias", "bias_batch_stride"),
                              ("gamma_h", "gamma_h", "gh_batch_stride"), ("gamma_c", "gamma_c", "gc_batch_stride")):
        a = t[key]
        setattr(d, name, a.data_ptr() if a.numel() else None)
        setattr(d, stride, bs(a) if a.numel() else 0)
    d.Q, d.b, d.A_eq, d.b_eq = Q.data_ptr(), b.data_ptr(), A.data_ptr(), beq.data_ptr()
    d.stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        nat.check(nat.lib().fccqp_wbc_assemble(C.byref(d)))
    lb = torch.full((n,), -float("inf"), dtype=torch.float64, device=dev)
    ub = torch.full((n,), float("inf"), dtype=torch.float64, device=dev)
    lb[nv:nv + nu] = -u_max
    ub[nv:nv + nu] = u_max
    assemble._keepalive = t   # inputs stay alive until the asynchronous launch has consumed them
    return Q, b, A, beq, t["friction_coeffs"], lb, ub
