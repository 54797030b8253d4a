// fccqp_warp.cuh -- one QP per WARP: the mapping for small problems (n + m <= 32).
//
// Same contract as fccqp_solve_kernel (fccqp_kernel.cuh): FCCQP::Solve + GetSolution of the reference
// (src/fcc_qp.cpp:114-207) per QP, FP64 throughout, same pre-solve (augmented-Lagrangian quasi-definite KKT matrix,
// unpivoted LDL^T), same lazy rho-KKT factorization, same order of operations in the z-update / residuals / duals /
// exit test.  What changes is the mapping: a CTA-per-QP kernel spends a 5-variable QP's time in block-wide barriers
// and 8x8-tile bookkeeping for a matrix that is smaller than one tile row.  Here lane i of a warp owns row i of the
// (n + m) x (n + m) KKT matrix and every per-row scalar (bounds, duals, right-hand side, pivot) in registers; the matrix
// itself sits in shared memory, one private slab per warp with an odd row stride (row-per-lane and column-per-lane
// accesses are both conflict-free, reads of the pivot row are broadcasts); lanes talk through shuffles; nothing but
// __syncwarp orders anything, so the warps of a CTA -- and of an SM -- run fully decoupled, each pulling QPs from the
// work counter on its own.
//
//   factorization   right-looking LDL^T on the full symmetric square: step k scales column k below the pivot and
//                   updates rows k+1.. with the (untouched) pivot row -- no divergence, (N-k) fused multiply-adds per lane
//   solves          forward / backward substitution, one shuffle broadcast per step
//   long-running    after SolveParams::full_inverse_at iterations the x-block of the inverse, G = [K^-1]_xx, replaces the
//   QPs             factors (n solves for unit vectors, lanes keep one ROW of G each): every later x-update is
//                   x = x_base + rho G (x_bar - mu), n shuffle-broadcast multiply-adds with no dependent chain
#pragma once
#include "fccqp_kernel.cuh"

namespace fccqp {

constexpr unsigned kFullMask = 0xffffffffu;
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFullMask, v, src); }

// LDL^T of the N x N symmetric matrix in K (row stride ld), in place: strictly-lower part = L, diagonal = D.
__device__ __forceinline__ void warp_ldlt(double* __restrict__ K, const int N, const int ld, const int lane) {
  double* const Krow = K + lane * ld;
#pragma unroll 1
  for (int k = 0; k + 1 < N; ++k) {
    const double* const prow = K + k * ld;
    const double rk = 1.0 / prow[k];
    if (lane > k && lane < N) {
      const double l = Krow[k] * rk;
#pragma unroll 4
      for (int j = k + 1; j < N; ++j) Krow[j] = fma(-l, prow[j], Krow[j]);
      Krow[k] = l;
    }
    __syncwarp();
  }
}

// x = K^-1 rhs with the factors of warp_ldlt; lane i holds rhs_i / returns x_i (lanes >= N: 0).
// (Not inlined: called from three places, 32 times over when the inverse of a long-running QP is formed.)
__device__ __noinline__ double warp_solve(const double* __restrict__ K, const int N, const int ld, const int lane,
                                          const double dinv, double y) {
  const double* const Krow = K + lane * ld;
#pragma unroll 4
  for (int k = 0; k + 1 < N; ++k) {
    const double yk = shfl_d(y, k);
    if (lane > k && lane < N) y = fma(-Krow[k], yk, y);
  }
  y *= dinv;
#pragma unroll 4
  for (int k = N - 1; k > 0; --k) {
    const double xk = shfl_d(y, k);
    if (lane < k) y = fma(-K[k * ld + lane], xk, y);
  }
  return y;
}

template <int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(32 * kWarps, kMinBlocks) fccqp_warp_kernel(const SolveParams p) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.n, m = p.m, nc = p.nc, lcs = p.lcs, N = n + m;
  const int ld = N | 1;
  double* const K = smem + (size_t)warp * (size_t)(N * ld);
  double* const Krow = K + (lane < N ? lane : 0) * ld;
  const bool is_x = lane < n, is_c = lane >= n && lane < N;
  const bool in_cone = is_x && lane >= lcs && lane < lcs + nc;
  const int ck = in_cone ? (lane - lcs) % 3 : 0;     // component inside the contact triple
  const int cbase = in_cone ? lane - ck : 0;         // first lane of the triple
  const long long q_slow = p.q_cs <= p.q_rs ? p.q_rs : p.q_cs;
  const long long q_fast = p.q_cs <= p.q_rs ? p.q_cs : p.q_rs;

  for (;;) {
    int qp = 0;
    if (lane == 0) {
      qp = (int)atomicAdd(p.work_counter, 1u);
      if (p.index_list && qp < p.B) qp = p.index_list[qp];    // processing order (FCCQP_SCHEDULE_LPT)
    }
    qp = __shfl_sync(kFullMask, qp, 0);
    if (qp >= p.B) break;
    const long long t_start = clock64();
    const double* Qg = p.Q + (size_t)qp * p.q_bs;
    const double* Ag = p.A + (size_t)qp * p.a_bs;

    // ---------------- vectors: one register per row and vector ----------------
    double v_b = 0.0, v_lb = 0.0, v_ub = 0.0, v_mux = 0.0, v_xbar = 0.0, v_x = 0.0;
    double v_muc = 0.0, v_lcbar = 0.0, v_fric = 0.0;
    int finite_bounds = 0;
    if (is_x) {
      v_b = p.b[(size_t)qp * p.b_bs + lane];
      v_lb = p.lb[(size_t)qp * p.lb_bs + lane];
      v_ub = p.ub[(size_t)qp * p.ub_bs + lane];
      if (p.warm) { v_x = p.x[(size_t)qp * n + lane]; v_mux = p.mu_x[(size_t)qp * n + lane]; }
      if (!isinf(v_lb) || !isinf(v_ub)) finite_bounds = 1;
    } else if (is_c) {
      v_b = p.beq[(size_t)qp * p.beq_bs + (lane - n)];
    }
    if (in_cone) {
      v_fric = p.mu[(size_t)qp * p.mu_bs + (lane - lcs) / 3];
      if (p.warm) v_muc = p.mu_c[(size_t)qp * nc + (lane - lcs)];
    }
    const bool eqc = (nc == 0) && !__any_sync(kFullMask, finite_bounds);   // fcc_qp.cpp:132-133
    const bool presolve = eqc || !p.warm;                                  // fcc_qp.cpp:159

    int status_flag = 0, n_iter = 0;
    bool reg_used = false, reg_mine = false;   // regularised retry for dependent constraint rows (fccqp_kernel.cuh)
    double reg_delta = 0.0;
    int reg_tries = 0;
    double res_x = 0.0, res_c = 0.0;
    unsigned long long fact_cycles = 0;

    for (int pass = presolve ? 0 : 1; pass < 2; ++pass) {
      if (pass == 1 && eqc) break;
      if (pass == 1) {
        v_xbar = v_x;            // fcc_qp.cpp:74-75
        v_lcbar = v_x;
        n_iter = p.max_iter;
      }
      const int iters = pass == 0 ? 1 : p.max_iter;
      bool factored = false, have_g = false;
      double dinv = 0.0, rhs0 = 0.0, v_xbase = 0.0;
      double rho_cur = p.rho;   // (changes only with the adaptive-rho extension)

#pragma unroll 1
      for (int iter = 0; iter < iters; ++iter) {
        double val = v_x;
        // x-update 0 of a cold solve is the identity (x_bar = x0, zero duals; fccqp_kernel.cuh): the rho-KKT system is
        // only factored for QPs that go on iterating
        if (!(pass == 1 && iter == 0 && presolve && p.first_update_identity)) {
          while (!factored) {
            factored = true;
            const long long t_f0 = clock64();
            __syncwarp();
            // ---- assemble [[Q (+ rho I), A'], [A, 0]]: coalesced rows of Q and A_eq, A_eq also transposed
#pragma unroll 2
            for (int i = 0; i < n; ++i)
              if (is_x) K[i * ld + lane] = Qg[(long long)i * q_slow + (long long)lane * q_fast];
#pragma unroll 2
            for (int r = 0; r < m; ++r) {
              if (is_x) {
                const double a = Ag[(long long)r * p.a_rs + (long long)lane * p.a_cs];
                K[(n + r) * ld + lane] = a;
                Krow[n + r] = a;
              } else if (is_c) {
                K[(n + r) * ld + lane] = 0.0;
              }
            }
            __syncwarp();
            if (reg_mine) Krow[lane] = -reg_delta;
            if (pass == 1) {
              if (is_x) Krow[lane] += rho_cur;
            } else {
              // sigma = trace(Q) / ||A||_F^2 balances the two terms of Q + sigma A'A (fccqp_kernel.cuh, pre-solve)
              double trq = is_x ? Krow[lane] : 0.0, fro = 0.0;
              if (is_c)
                for (int j = 0; j < n; ++j) fro = fma(Krow[j], Krow[j], fro);
              trq = warp_sum(trq); fro = warp_sum(fro);
              const double sigma = (trq > 0.0 && fro > 0.0 && isfinite(trq / fro)) ? trq / fro : 1.0;
              // rhs_x = -b + sigma A' b_eq, rhs_c = b_eq; Q += sigma A'A
              double s = 0.0;
              for (int r = 0; r < m; ++r) {
                const double br = shfl_d(v_b, n + r);
                if (is_x) s = fma(Krow[n + r], br, s);
              }
              rhs0 = is_x ? fma(sigma, s, -v_b) : (is_c ? v_b : 0.0);
              if (is_x) {
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                  const double* cj = K + j * ld + n;      // column j of A_eq = row j of the transposed copy (broadcast)
                  double t = 0.0;
                  for (int r = 0; r < m; ++r) t = fma(Krow[n + r], cj[r], t);
                  Krow[j] = fma(sigma, t, Krow[j]);
                }
              }
            }
            __syncwarp();
            warp_ldlt(K, N, ld, lane);
            const double d = lane < N ? Krow[lane] : 1.0;
            dinv = 1.0 / d;
            {
              // inertia (+ on the variable rows, - on the constraint rows) and pivot size (kPivotRatio), as in the
              // CTA kernels: anything else is FCCQP_STATUS_NUMERICAL_ISSUE
              bool badp = lane < N && (!isfinite(d) || (is_c ? !(d < 0.0) : !(d > 0.0)));
              const double pa = warp_max((is_x && pass == 0) ? fabs(d) : 0.0), pc = warp_max(is_c ? fabs(d) : 0.0);
              if (lane < N && fabs(d) < kPivotRatio * (is_c ? pc : pa)) badp = true;
              const unsigned fmask = __ballot_sync(kFullMask, badp);
              const int row = fmask ? __ffs(fmask) - 1 : -1;       // first failing row
              if (fmask && row >= n && row < N && reg_tries < kRegTries && isfinite(pc) && pc > 0.0) {
                // a dependent constraint row: -delta on its diagonal and one more attempt (fccqp_kernel.cuh)
                if (reg_delta == 0.0) reg_delta = kRegDelta * pc;
                if (lane == row) reg_mine = true;
                reg_used = true;
                ++reg_tries;
                factored = false;
              } else if (fmask) {
                status_flag = 2;
              }
            }
            fact_cycles += (unsigned long long)(clock64() - t_f0);
          }

          const double w = (pass == 1 && is_x) ? (in_cone ? (v_lcbar - v_muc) : (v_xbar - v_mux)) : 0.0;
          if (pass == 1 && !have_g && iter >= p.full_inverse_at) {
            // long-running QP: x_base = [K^-1 (-b; b_eq)]_x, then G = [K^-1]_xx column by column (one solve per unit
            // vector; lane i keeps G[i][c], which is G[c][i]).  The solves read all of L, so G cannot be built in
            // place: it is collected in registers (n <= 32 columns) and written over the factors afterwards.
            v_xbase = warp_solve(K, N, ld, lane, dinv, is_x ? -v_b : (is_c ? v_b : 0.0));
            double g[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) g[c] = 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < n) g[c] = warp_solve(K, N, ld, lane, dinv, lane == c ? 1.0 : 0.0);   // column c of K^-1; keeps G[lane][c]
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < n && is_x) Krow[c] = g[c];
            __syncwarp();
            have_g = true;
          }
          if (have_g) {
            // x = x_base + rho G w
            double s0 = 0.0, s1 = 0.0;
            int c = 0;
#pragma unroll 2
            for (; c + 1 < n; c += 2) {
              s0 = fma(Krow[c], shfl_d(w, c), s0);
              s1 = fma(Krow[c + 1], shfl_d(w, c + 1), s1);
            }
            if (c < n) s0 = fma(Krow[c], shfl_d(w, c), s0);
            val = is_x ? fma(rho_cur, s0 + s1, v_xbase) : 0.0;
          } else {
            // -(b + q_rho), q_rho = -rho (xbar - mu_x) with the cone segment overwritten (fcc_qp.cpp:81-83)
            const double acc = pass == 0 ? rhs0 : (is_x ? -(v_b - rho_cur * w) : (is_c ? v_b : 0.0));
            val = warp_solve(K, N, ld, lane, dinv, acc);
            if (!is_x) val = 0.0;
          }
        }

        if (pass == 0) {
          v_x = val;
          if (p.dbg_x0 && is_x) p.dbg_x0[(size_t)qp * n + lane] = val;
          continue;
        }

        // ---- z-update, residuals, duals (fcc_qp.cpp:88-103), exit test (:105-109)
        v_x = val;
        double rx = 0.0, rc = 0.0, dz = 0.0;
        const bool relax = p.alpha != 1.0;
        if (is_x) {
          const double xh = relax ? fma(p.alpha, val, (1.0 - p.alpha) * v_xbar) : val;
          const double xb = clampd(xh + v_mux, v_lb, v_ub);
          dz = fabs(xb - v_xbar);
          v_xbar = xb;
          const double r = xh - xb;
          v_mux += r;
          rx = fabs(r);
        }
        {
          // every lane of a contact triple projects the whole triple and keeps its own component
          const double xk = relax ? fma(p.alpha, val, (1.0 - p.alpha) * v_lcbar) : val;
          const double fk = xk + v_muc;
          const double f0 = shfl_d(fk, cbase), f1 = shfl_d(fk, cbase + 1), f2 = shfl_d(fk, cbase + 2);
          if (in_cone) {
            double o0, o1, o2;
            project_cone3(f0, f1, f2, v_fric, o0, o1, o2);
            const double o = ck == 0 ? o0 : (ck == 1 ? o1 : o2);
            dz = fmax(dz, fabs(o - v_lcbar));
            v_lcbar = o;
            const double r = xk - o;
            v_muc += r;
            rc = fabs(r);
          }
        }
        if (rx != rx || rc != rc) status_flag = 2;   // (ballot below: any lane)
        const bool conv = __all_sync(kFullMask, (rc < p.eps_fcone) && (rx < p.eps_bound));
        if (conv || iter + 1 == iters) {
          res_x = warp_max(rx); res_c = warp_max(rc);
          if (conv) { n_iter = iter; break; }
        } else if (p.adapt_k > 0 && (iter + 1) % p.adapt_k == 0) {
          // adaptive rho (extension; see fccqp_kernel.cuh and oracle/fccqp_oracle.c, do_admm)
          const double rp = warp_max(fmax(rx, rc)), rd = rho_cur * warp_max(dz);
          double ratio = sqrt(rp / (rd > 1e-300 ? rd : 1e-300));
          ratio = fmin(fmax(ratio, 0.1), 10.0);
          if (ratio > 5.0 || ratio < 0.2) {
            const double rho_new = fmin(fmax(rho_cur * ratio, 1e-9), 1e9);
            const double sc = rho_cur / rho_new;
            v_mux *= sc; v_muc *= sc;
            rho_cur = rho_new;
            factored = false;
            have_g = false;
          }
        }
      }
    }

    // ---------------- epilogue: violations (constraint_utils.cpp:48-65), outputs ----------------
    double bv = 0.0, fv = 0.0;
    if (is_x) { const double d = v_x - clampd(v_x, v_lb, v_ub); bv = d * d; }
    {
      const double x1 = shfl_d(v_x, cbase + 1), x2 = shfl_d(v_x, cbase + 2);
      if (in_cone && ck == 0) {
        const double r = sqrt(v_x * v_x + x1 * x1) - v_fric * x2;
        fv = r > 0.0 ? r : 0.0;
      }
    }
    bv = warp_sum(bv); fv = warp_sum(fv);
    bool incons = false;
    if (reg_used) {
      // regularised retry taken: the answer is the reference's only if A_eq x = b_eq still holds (consistent dependent rows)
      double ax = 0.0, mag = fabs(v_b);
      for (int j = 0; j < n; ++j) {
        const double xj = shfl_d(v_x, j);
        if (is_c) { const double term = Ag[(long long)(lane - n) * p.a_rs + (long long)j * p.a_cs] * xj; ax += term; mag += fabs(term); }
      }
      incons = is_c && !(fabs(ax - v_b) <= 1e-7 * mag + 1e-300);
    }
    const bool bad = __any_sync(kFullMask, (is_x && !isfinite(v_x)) || status_flag == 2 || incons);
    if (is_x) {
      p.x[(size_t)qp * n + lane] = v_x;
      if (p.mu_x) p.mu_x[(size_t)qp * n + lane] = v_mux;
    }
    if (p.mu_c && in_cone) p.mu_c[(size_t)qp * nc + (lane - lcs)] = v_muc;
    if (lane == 0) {
      if (p.n_iter) p.n_iter[qp] = n_iter;
      if (p.status) p.status[qp] = bad ? 2 : (n_iter == p.max_iter ? 1 : 0);   // fcc_qp.cpp:203-204
      if (p.res_b) p.res_b[qp] = res_x;
      if (p.res_f) p.res_f[qp] = res_c;
      if (p.bviol) p.bviol[qp] = sqrt(bv);
      if (p.fviol) p.fviol[qp] = fv;
      if (p.cycles) {
        atomicAdd(p.cycles, fact_cycles);
        atomicAdd(p.cycles + 1, (unsigned long long)(clock64() - t_start));
      }
    }
  }
}

}  // namespace fccqp
