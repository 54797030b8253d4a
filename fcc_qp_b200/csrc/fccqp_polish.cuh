// fccqp_polish.cuh -- opt-in solution polish (the reference has no counterpart: src/fcc_qp.cpp
// returns the ADMM iterate as it is, fccqp.pdf Table 2 lists polish as an OSQP feature FCCQP lacks).
//
// After an ADMM solve the constraints that hold with equality are guessed from the solver's own state (x, mu_x, mu_c): a
// bounded variable with x + mu_x at or beyond a bound sits ON that bound; a contact whose argument of the cone projection
// f = x_c + mu_c (constraint_utils.cpp:5-25) falls in the polar cone sits at the APEX (lambda_c = 0), one that falls
// outside both cones sits on the cone's BOUNDARY: it may move in the tangent plane of the cone at its projection o,
// lambda_c = alpha o^ + beta t1 (o^ = o / |o| the ray, t1 the horizontal tangent), not along the normal.
// With that active set the QP is an equality-constrained one of the SAME size:
//   * a boundary contact is rotated into the basis R_c = [o^, t1, t2]: (alpha, beta, gamma), gamma (the normal) fixed to 0
//     -- the tangent plane leaves the cone only to second order in beta, so a good guess lands within the cone tolerance;
//   * fixed variables (bounds, apex forces, the gammas) are decoupled: unit diagonal, their coupling moved to the right-hand
//     sides, their columns of A_eq zeroed;
// and the existing solver takes it as a QP with nc = 0 and no bounds (its pre-solve IS the KKT solve, src/fcc_qp.cpp:159-178).
// polish_finish rotates the answer back and ACCEPTS it per QP only if that solve succeeded and the point satisfies every bound
// and every friction cone to the solver's tolerances; otherwise the ADMM result stays.
#pragma once
#include "fccqp_kernel.cuh"

namespace fccqp {

struct PolishParams {
  int B, n, m, nc, lcs;
  double eps_fcone, eps_bound, eps_objective;
  // the QP (dense row-major per QP; batch strides in elements, 0 = shared by all QPs)
  const double* Q;   long long q_bs;
  const double* b;   long long b_bs;
  const double* A;   long long a_bs;
  const double* beq; long long beq_bs;
  const double* mu;  long long mu_bs;
  const double* lb;  long long lb_bs;
  const double* ub;  long long ub_bs;
  // the ADMM result
  const double* x; const double* mu_x; const double* mu_c;
  // the polished QP, contiguous [B, ...]
  double* Qp; double* bp; double* Ap; double* beqp;
  double* rot;          // [B, nc/3, 4]: kind (0 interior, 1 apex, 2 boundary), o^ (3)
  // finish
  const double* y; const int* y_status;
  double* z; double* bviol; double* fviol; int* polished;
};

// One CTA per QP (grid-stride).  Shared memory: w[n][3] weights, g[n], val[n] doubles; cnt[n], kind[n], base[n] ints.
__global__ void __launch_bounds__(128) polish_prepare_kernel(const PolishParams p) {
  extern __shared__ __align__(16) double sm[];
  const int n = p.n, m = p.m, nc = p.nc, lcs = p.lcs, ncon = nc / 3;
  double* const w = sm;                  // [n][3] weights of variable j over the original variables base[j] + k
  double* const g = w + 3 * n;           // [n]   b + Q[:, Fb] v
  double* const val = g + n;             // [n]   value of a fixed variable
  int* const cnt = reinterpret_cast<int*>(val + n);    // [n]   1, or 3 for the alpha of a boundary contact
  int* const kind = cnt + n;            // [n]   0 free, 1 fixed at val (a bound), 2 fixed at 0 (apex force / gamma)
  int* const base = kind + n;           // [n]   first original variable the weights of this variable refer to
  const int t = threadIdx.x, nt = blockDim.x;
  for (int qp = blockIdx.x; qp < p.B; qp += gridDim.x) {
    const double* Q = p.Q + (size_t)qp * p.q_bs;
    const double* bv = p.b + (size_t)qp * p.b_bs;
    const double* A = p.A + (size_t)qp * p.a_bs;
    const double* beq = p.beq + (size_t)qp * p.beq_bs;
    const double* x = p.x + (size_t)qp * n;
    const double* mux = p.mu_x + (size_t)qp * n;
    __syncthreads();   // previous QP done with the shared arrays
    // ---- active set.  Variables outside the contact block: bounds.
    for (int i = t; i < n; i += nt) {
      int k = 0; double v = 0.0;
      if (i < lcs || i >= lcs + nc) {
        const double lo = p.lb[(size_t)qp * p.lb_bs + i], hi = p.ub[(size_t)qp * p.ub_bs + i];
        const double a = x[i] + mux[i];
        if (a <= lo) { k = 1; v = lo; }
        else if (a >= hi) { k = 1; v = hi; }
      }
      kind[i] = k; val[i] = v; cnt[i] = 1; base[i] = i;
      w[3 * i] = 1.0; w[3 * i + 1] = 0.0; w[3 * i + 2] = 0.0;
    }
    __syncthreads();
    // Contacts: the branch of the cone projection taken by f = x_c + mu_c (constraint_utils.cpp:5-25)
    for (int c = t; c < ncon; c += nt) {
      const int o = lcs + 3 * c;
      const double mu = p.mu[(size_t)qp * p.mu_bs + c];
      const double* muc = p.mu_c + (size_t)qp * nc + 3 * c;
      const double f0 = x[o] + muc[0], f1 = x[o + 1] + muc[1], f2 = x[o + 2] + muc[2];
      double o0, o1, o2;
      project_cone3(f0, f1, f2, mu, o0, o1, o2);
      const double r = sqrt(f0 * f0 + f1 * f1);
      double* rot = p.rot + ((size_t)qp * ncon + c) * 4;
      const double no = sqrt(o0 * o0 + o1 * o1 + o2 * o2);
      if (mu * f2 >= r) {                                   // inside the cone: free
        rot[0] = 0.0; rot[1] = rot[2] = rot[3] = 0.0;
      } else if (f2 < -mu * r || !(no > 0.0)) {             // polar cone (or projected to the origin): apex
        rot[0] = 1.0; rot[1] = rot[2] = rot[3] = 0.0;
        kind[o] = kind[o + 1] = kind[o + 2] = 2;
      } else {                                              // boundary: lambda = alpha o^ + beta t1 in the slots of o, o + 1
        const double d0 = o0 / no, d1 = o1 / no, d2 = o2 / no;
        rot[0] = 2.0; rot[1] = d0; rot[2] = d1; rot[3] = d2;
        w[3 * o] = d0; w[3 * o + 1] = d1; w[3 * o + 2] = d2; cnt[o] = 3;
        const double h = sqrt(d0 * d0 + d1 * d1);
        if (h > 0.0) {                                       // t1 = (-d1, d0, 0) / h
          w[3 * (o + 1)] = -d1 / h; w[3 * (o + 1) + 1] = d0 / h; w[3 * (o + 1) + 2] = 0.0; cnt[o + 1] = 3; base[o + 1] = o;
        } else kind[o + 1] = 2;                              // (mu = 0: the cone is a ray, nothing tangential)
        kind[o + 2] = 2;
      }
    }
    __syncthreads();
    // ---- g = b + Q[:, Fb] v over the bound-fixed variables (the only fixed ones with a nonzero value)
    for (int a = t; a < n; a += nt) {
      double s = bv[a];
      for (int i = 0; i < n; ++i)
        if (kind[i] == 1) s = fma(Q[(size_t)a * n + i], val[i], s);
      g[a] = s;
    }
    __syncthreads();
    // ---- polished QP
    double* Qp = p.Qp + (size_t)qp * n * n;
    for (int e = t; e < n * n; e += nt) {
      const int i = e / n, j = e - i * n;
      double q = 0.0;
      if (kind[i] != 0 || kind[j] != 0) q = (i == j) ? 1.0 : 0.0;
      else {
        const int ni = cnt[i], nj = cnt[j];
        for (int k = 0; k < ni; ++k)
          for (int l = 0; l < nj; ++l) q = fma(w[3 * i + k] * w[3 * j + l], Q[(size_t)(base[i] + k) * n + (base[j] + l)], q);
      }
      Qp[e] = q;
    }
    double* bp = p.bp + (size_t)qp * n;
    for (int j = t; j < n; j += nt) {
      double s;
      if (kind[j] != 0) s = -val[j];
      else {
        const int nj = cnt[j];
        s = 0.0;
        for (int l = 0; l < nj; ++l) s = fma(w[3 * j + l], g[base[j] + l], s);
      }
      bp[j] = s;
    }
    double* Ap = p.Ap + (size_t)qp * m * n;
    for (int e = t; e < m * n; e += nt) {
      const int k = e / n, j = e - k * n;
      double s = 0.0;
      if (kind[j] == 0) {
        const int nj = cnt[j];
        for (int l = 0; l < nj; ++l) s = fma(w[3 * j + l], A[(size_t)k * n + base[j] + l], s);
      }
      Ap[e] = s;
    }
    double* beqp = p.beqp + (size_t)qp * m;
    for (int k = t; k < m; k += nt) {
      double s = beq[k];
      for (int i = 0; i < n; ++i)
        if (kind[i] == 1) s = fma(-A[(size_t)k * n + i], val[i], s);
      beqp[k] = s;
    }
  }
}

// One warp per QP: rotate back, accept or leave.  Accepted = the inner solve succeeded, the point satisfies A_eq z = b_eq,
// every bound and every friction cone to the tolerances, AND its objective is not above the ADMM iterate's by more than eps_objective
// (relative): a feasible point of a WRONG active-set guess is optimal for the wrong problem, and this is what catches it
// (the ADMM iterate satisfies A_eq z = b_eq exactly and violates the inequalities by at most its residuals, so its objective
// is a lower estimate of the optimum).  Shared memory: two vectors of n doubles per warp.
__global__ void __launch_bounds__(128) polish_finish_kernel(const PolishParams p) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.n, nc = p.nc, lcs = p.lcs, ncon = nc / 3;
  double* const xp = sm + (size_t)warp * 2 * n;     // polished point
  double* const za = xp + n;                         // ADMM point
  for (int qp = blockIdx.x * (blockDim.x >> 5) + warp; qp < p.B; qp += gridDim.x * (blockDim.x >> 5)) {
    const double* y = p.y + (size_t)qp * n;
    const double* rot = p.rot + (size_t)qp * ncon * 4;
    const double* Q = p.Q + (size_t)qp * p.q_bs;
    const double* bv = p.b + (size_t)qp * p.b_bs;
    bool ok = p.y_status[qp] == 0;
    double bviol = 0.0, fviol = 0.0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      double xi = y[i];
      if (i >= lcs && i < lcs + nc) {
        const int c = (i - lcs) / 3, k = (i - lcs) - 3 * c;
        const double* r = rot + 4 * c;
        if (r[0] == 1.0) xi = 0.0;
        else if (r[0] == 2.0) {
          const double h = sqrt(r[1] * r[1] + r[2] * r[2]);
          const double t1k = h > 0.0 ? (k == 0 ? -r[2] / h : (k == 1 ? r[1] / h : 0.0)) : 0.0;
          xi = fma(y[lcs + 3 * c + 1], t1k, y[lcs + 3 * c] * r[1 + k]);
        }
      }
      const double lo = p.lb[(size_t)qp * p.lb_bs + i], hi = p.ub[(size_t)qp * p.ub_bs + i];
      if (!isfinite(xi) || xi < lo - p.eps_bound || xi > hi + p.eps_bound) ok = false;
      const double d = xi - clampd(xi, lo, hi);
      bviol = fma(d, d, bviol);
      xp[i] = xi;
      za[i] = p.z[(size_t)qp * n + i];
    }
    __syncwarp();
    for (int c = lane; c < ncon; c += 32) {
      const int o = lcs + 3 * c;
      const double viol = sqrt(xp[o] * xp[o] + xp[o + 1] * xp[o + 1]) - p.mu[(size_t)qp * p.mu_bs + c] * xp[o + 2];   // constraint_utils.cpp:48-59
      if (!(viol <= p.eps_fcone)) ok = false;
      fviol += viol > 0.0 ? viol : 0.0;
    }
    // A_eq z = b_eq must hold on the polished point (a guess that fixes too much leaves an inconsistent system; the same
    // test as the solver's own after a regularised retry: 1e-7 relative to the magnitude of the terms)
    {
      const double* A = p.A + (size_t)qp * p.a_bs;
      const double* beq = p.beq + (size_t)qp * p.beq_bs;
      double xmax = 1.0;
      for (int i = lane; i < n; i += 32) xmax = fmax(xmax, fabs(xp[i]));
      xmax = warp_max(xmax);
      for (int k = lane; k < p.m; k += 32) {
        const double* arow = A + (size_t)k * n;
        double ax = 0.0, mag = fabs(beq[k]), asum = 0.0;
        for (int j = 0; j < n; ++j) { const double term = arow[j] * xp[j]; ax += term; mag += fabs(term); asum += fabs(arow[j]); }
        // (+ an absolute floor at rounding level of the row: rows like eps_k = 0 hold to 1e-15, not to 1e-7 of nothing)
        if (!(fabs(ax - beq[k]) <= 1e-7 * mag + 1e-10 * asum * xmax)) ok = false;
      }
    }
    // objectives 1/2 z'Qz + b'z of both points
    double fp = 0.0, fa = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double* qrow = Q + (size_t)i * n;
      double sp = 0.0, sa = 0.0;
      for (int j = 0; j < n; ++j) { const double q = qrow[j]; sp = fma(q, xp[j], sp); sa = fma(q, za[j], sa); }
      fp = fma(xp[i], fma(0.5, sp, bv[i]), fp);
      fa = fma(za[i], fma(0.5, sa, bv[i]), fa);
    }
    fp = warp_sum(fp); fa = warp_sum(fa);
    if (!(fp <= fa + p.eps_objective * fmax(1.0, fabs(fa)))) ok = false;
    ok = __all_sync(0xffffffffu, ok);
    bviol = warp_sum(bviol); fviol = warp_sum(fviol);
    if (ok)
      for (int i = lane; i < n; i += 32) p.z[(size_t)qp * n + i] = xp[i];
    if (lane == 0) {
      p.polished[qp] = ok ? 1 : 0;
      if (ok) { p.bviol[qp] = sqrt(bviol); p.fviol[qp] = fviol; }
    }
  }
}

}  // namespace fccqp
