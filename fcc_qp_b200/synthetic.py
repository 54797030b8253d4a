"""Synthetic whole-body-control QPs of the shapes named in BASELINE.json.

The generator mirrors the structure of the logged Cassie OSC QPs (SURVEY.md 8d;
``fccqp.pdf`` section 4): decision vector ``x = [vdot(nv) u(nu) lambda_h(nh)
lambda_c(nc) eps(ne)]``, cost on task-space acceleration error plus small
regularisation, dynamics / holonomic / soft-contact equality rows, torque box
bounds and one Coulomb cone per 3-D contact force.

    Q = blkdiag(Jy' W Jy + 1e-5 I, 1e-4 I, 0, 1e-6 I, 80 I)
    b = [-Jy' W ydd_cmd; 0]
    A = [[M, -B, -Jh', -Jc', 0], [Jh, 0, 0, 0, 0], [Jc, 0, 0, 0, I]]
    b_eq = [-C; -gamma_h; -gamma_c]

All arrays are float64, C-contiguous and deterministic in ``seed``.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .logdata import QPBatch


@dataclasses.dataclass(frozen=True)
class Shape:
    name: str
    nv: int
    nu: int
    nh: int
    nc: int
    seed: int
    joint_scale: float = 0.1  # size of the joint columns of the contact Jacobian (tunes how often cones bind)

    @property
    def ne(self) -> int:
        return self.nc

    @property
    def n(self) -> int:
        return self.nv + self.nu + self.nh + self.nc + self.ne

    @property
    def m(self) -> int:
        return self.nv + self.nh + self.ne

    @property
    def lambda_c_start(self) -> int:
        return self.nv + self.nu + self.nh


# SURVEY.md 8d configs 3, 4, 5 (+ the Cassie shape of the walking log for reference)
HUMANOID = Shape("humanoid", nv=36, nu=30, nh=0, nc=12, seed=20241, joint_scale=0.1)        # n=90  m=48 N=138
QUADRUPED = Shape("quadruped", nv=18, nu=12, nh=0, nc=12, seed=20242, joint_scale=0.05)      # n=54  m=30 N=84
MULTICONTACT = Shape("multicontact", nv=36, nu=30, nh=6, nc=24, seed=20243, joint_scale=0.03)  # n=120 m=66 N=186
CASSIE_LIKE = Shape("cassie_like", nv=22, nu=10, nh=6, nc=12, seed=20240, joint_scale=0.05)  # n=60  m=38 N=98
SHAPES = {s.name: s for s in (HUMANOID, QUADRUPED, MULTICONTACT, CASSIE_LIKE)}


@dataclasses.dataclass
class WBCTerms:
    """Robot quantities one WBC QP is assembled from (``fccqp.pdf`` section 4; ``include/fccqp.h``,
    ``fccqp_wbc_assemble``).  All arrays carry the batch as leading dimension, float64, C-contiguous."""
    shape: Shape
    M: np.ndarray          # [B,nv,nv] mass matrix
    Jh: np.ndarray         # [B,nh,nv] holonomic Jacobian
    Jc: np.ndarray         # [B,nc,nv] contact Jacobian (zero rows for feet off the ground)
    Jy: np.ndarray         # [B,ny,nv] task Jacobian
    W: np.ndarray          # [B,ny]    task weights
    ydd_cmd: np.ndarray    # [B,ny]    commanded task accelerations
    bias: np.ndarray       # [B,nv]    Coriolis + gravity
    gamma_h: np.ndarray    # [B,nh]
    gamma_c: np.ndarray    # [B,nc]
    friction_coeffs: np.ndarray  # [B,nc/3]
    u_max: float = 300.0
    weights: tuple = (1e-5, 1e-4, 1e-6, 80.0)   # w_vdot, w_u, w_lambda_c, w_eps

    @property
    def batch(self) -> int:
        return int(self.M.shape[0])

    def nbytes(self) -> int:
        return int(sum(a.nbytes for a in (self.M, self.Jh, self.Jc, self.Jy, self.W, self.ydd_cmd, self.bias,
                                          self.gamma_h, self.gamma_c, self.friction_coeffs)))


def make_terms(shape: Shape, batch: int, seed: int | None = None, u_max: float = 300.0,
               ydd_sigma: float = 1.0, bias_sigma: float = 1.0, joint_scale: float | None = None,
               height: float = 0.75, weight: float = 300.0) -> WBCTerms:
    """Robot quantities of B synthetic QPs of ``shape`` (see ``make_batch`` for the tuning)."""
    joint_scale = shape.joint_scale if joint_scale is None else joint_scale
    rng = np.random.default_rng(shape.seed if seed is None else seed)
    B, nv, nu, nh, nc, ne = batch, shape.nv, shape.nu, shape.nh, shape.nc, shape.ne
    ncon = nc // 3
    ny = max(nv - 4, 1)

    G = rng.standard_normal((B, nv, nv))
    Mass = G @ G.transpose(0, 2, 1) / nv
    Mass[:, np.arange(nv), np.arange(nv)] += rng.uniform(0.05, 3.0, (B, nv))

    # Point contacts around the floating base: J_c = [I3, -[p]x, J_joints].  With the gravity-like
    # bias below, the (almost cost-free) contact forces mostly come out inside their cones, like
    # the logged walking QPs; larger commands / fewer stance feet make the cones active.
    # Contacts sit at the corners of a support rectangle (two rings when ncon > 4); a QP is in
    # full support or on one diagonal pair, so the weight can be carried by normal forces alone.
    corner = np.array([[1.0, 1.0], [1.0, -1.0], [-1.0, 1.0], [-1.0, -1.0]])
    pos = np.zeros((B, ncon, 3))
    for c in range(ncon):
        ring = 1.0 + 0.5 * (c // 4)
        pos[:, c, :2] = corner[c % 4] * np.array([0.25, 0.15]) * ring + rng.uniform(-0.03, 0.03, (B, 2))
    pos[:, :, 2] = -rng.uniform(0.8, 1.2, (B, ncon)) * height
    pattern = rng.integers(0, 4, B)  # 0,1: all feet; 2: diagonal {0,3}; 3: diagonal {1,2}
    stance = np.ones((B, ncon), dtype=bool)
    for c in range(ncon):
        stance[:, c] &= ~((pattern == 2) & (c % 4 in (1, 2))) & ~((pattern == 3) & (c % 4 in (0, 3)))
    Jc = np.zeros((B, ncon, 3, nv))
    Jc[:, :, np.arange(3), np.arange(3)] = 1.0
    px, py, pz = pos[..., 0], pos[..., 1], pos[..., 2]
    zero = np.zeros_like(px)
    skew = np.stack([np.stack([zero, -pz, py], -1), np.stack([pz, zero, -px], -1),
                     np.stack([-py, px, zero], -1)], -2)  # [p]x
    Jc[:, :, :, 3:6] = -skew
    if nv > 6:
        Jj = rng.standard_normal((B, ncon, 3, nv - 6)) * joint_scale
        Jj *= rng.random((B, ncon, 3, nv - 6)) < 0.5
        Jc[:, :, :, 6:] = Jj
    Jc = Jc.reshape(B, nc, nv) * np.repeat(stance, 3, axis=1)[:, :, None]
    Jh = rng.standard_normal((B, nh, nv)) * 0.5 if nh else np.zeros((B, 0, nv))
    if nh:
        Jh[:, :, :6] = 0.0  # hand/loop constraints act on joints only

    Jy = rng.standard_normal((B, ny, nv))
    W = rng.uniform(0.1, 20.0, (B, ny))
    ydd = rng.standard_normal((B, ny)) * ydd_sigma

    Cg = rng.standard_normal((B, nv)) * bias_sigma
    Cg[:, :6] *= 0.25
    Cg[:, 2] += weight  # weight on the vertical floating-base coordinate
    gamma_h = rng.standard_normal((B, nh)) * 0.1
    gamma_c = rng.standard_normal((B, ne)) * 0.1
    mu = rng.uniform(0.4, 1.0, (B, ncon))
    c = np.ascontiguousarray
    return WBCTerms(shape, c(Mass), c(Jh), c(Jc), c(Jy), c(W), c(ydd), c(Cg), c(gamma_h), c(gamma_c), c(mu), u_max)


def assemble_numpy(t: WBCTerms) -> QPBatch:
    """CPU restatement of ``fccqp_wbc_assemble`` (the parity reference of the device kernel):
    ``Q, b, A_eq, b_eq, lb, ub`` of the QPs described by ``t`` (formulas in the module docstring)."""
    shape = t.shape
    B, nv, nu, nh, nc, ne = t.batch, shape.nv, shape.nu, shape.nh, shape.nc, shape.ne
    n, m = shape.n, shape.m
    w_v, w_u, w_lc, w_eps = t.weights
    Bsel = np.zeros((nv, nu))
    Bsel[nv - nu:, :] = np.eye(nu)  # floating base (first nv - nu dofs) is unactuated

    Q = np.zeros((B, n, n))
    Q[:, :nv, :nv] = np.einsum("bki,bk,bkj->bij", t.Jy, t.W, t.Jy)
    d = np.concatenate([np.full(nv, w_v), np.full(nu, w_u), np.zeros(nh), np.full(nc, w_lc), np.full(ne, w_eps)])
    Q[:, np.arange(n), np.arange(n)] += d
    Q = 0.5 * (Q + Q.transpose(0, 2, 1))
    b = np.zeros((B, n))
    b[:, :nv] = -np.einsum("bki,bk->bi", t.Jy, t.W * t.ydd_cmd)

    o_u, o_h, o_c, o_e = nv, nv + nu, nv + nu + nh, nv + nu + nh + nc
    A = np.zeros((B, m, n))
    A[:, :nv, :nv] = t.M
    A[:, :nv, o_u:o_h] = -Bsel
    if nh:
        A[:, :nv, o_h:o_c] = -t.Jh.transpose(0, 2, 1)
        A[:, nv:nv + nh, :nv] = t.Jh
    A[:, :nv, o_c:o_e] = -t.Jc.transpose(0, 2, 1)
    A[:, nv + nh:, :nv] = t.Jc
    A[:, nv + nh:, o_e:] = np.eye(ne)
    beq = np.concatenate([-t.bias, -t.gamma_h, -t.gamma_c], axis=1)

    lb = np.full((B, n), -np.inf)
    ub = np.full((B, n), np.inf)
    lb[:, o_u:o_h] = -t.u_max
    ub[:, o_u:o_h] = t.u_max
    c = np.ascontiguousarray
    return QPBatch(n, m, nc, shape.lambda_c_start, c(Q), c(b), c(A), c(beq), c(t.friction_coeffs), c(lb), c(ub))


def make_batch(shape: Shape, batch: int, seed: int | None = None, **kw) -> QPBatch:
    """B synthetic QPs of ``shape``.  Defaults are tuned (against the reference solver, rho=5e-5,
    eps=1e-6, max_iter=100) so that most pre-solve points are feasible, 15-30 % of the QPs need
    ADMM iterations and a few per cent run to max_iter -- the mix seen on the walking log."""
    return assemble_numpy(make_terms(shape, batch, seed, **kw))


def random_walk(qp: QPBatch, rng: np.random.Generator, sigma: float = 0.02) -> QPBatch:
    """Next step of a sequential warm-started scenario: b and b_eq
    take a 2 % multiplicative random-walk step, everything else is carried over."""
    b = qp.b * (1.0 + sigma * rng.standard_normal(qp.b.shape))
    beq = qp.b_eq * (1.0 + sigma * rng.standard_normal(qp.b_eq.shape))
    return dataclasses.replace(qp, b=np.ascontiguousarray(b), b_eq=np.ascontiguousarray(beq))


def random_qps(rng, B, n, m, nc, lcs):
    """Well-conditioned convex QPs of arbitrary small shape with full-row-rank A_eq, a few active bounds and cones (the
    randomised-shape parity tests and the small-QP benchmarks draw from here)."""
    from .logdata import QPBatch
    G = rng.standard_normal((B, n, n))
    Q = G @ G.transpose(0, 2, 1) / n + np.eye(n) * rng.uniform(0.05, 1.0, (B, 1, 1))
    Q = 0.5 * (Q + Q.transpose(0, 2, 1))
    A = rng.standard_normal((B, m, n)) * (rng.random((B, m, n)) < 0.6)
    A[:, np.arange(m), np.arange(m)] += 2.0                    # keeps the rows independent
    b = rng.standard_normal((B, n)) * 2.0
    beq = rng.standard_normal((B, m))
    lb = np.full((B, n), -np.inf); ub = np.full((B, n), np.inf)
    nb = max(1, n // 4)
    idx = rng.choice(n, nb, replace=False)
    lb[:, idx] = -rng.uniform(0.05, 0.5, (B, nb)); ub[:, idx] = rng.uniform(0.05, 0.5, (B, nb))
    mu = rng.uniform(0.3, 1.0, (B, max(nc // 3, 0)))
    c = np.ascontiguousarray
    return QPBatch(n, m, nc, lcs, c(Q), c(b), c(A), c(beq), c(mu), c(lb), c(ub))


def scale_constraint_rows(qp, rng, spread):
    """The same QPs with every row of A_eq (and b_eq) scaled by 10^u, u uniform in [-spread, spread]: the solutions do not
    change, the conditioning of the KKT matrix does (FP32-mode accuracy against conditioning, tools/fp32_run.py)."""
    from .logdata import QPBatch
    sc = 10.0 ** rng.uniform(-spread, spread, (qp.batch, qp.m))
    return QPBatch(qp.n, qp.m, qp.nc, qp.lambda_c_start, qp.Q, qp.b, np.ascontiguousarray(qp.A_eq * sc[:, :, None]),
                   np.ascontiguousarray(qp.b_eq * sc), qp.friction_coeffs, qp.lb, qp.ub)
