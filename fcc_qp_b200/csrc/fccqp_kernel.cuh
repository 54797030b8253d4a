// fccqp_kernel.cuh -- sm_100a device code of the batched FCCQP solve.
//
// One CTA solves one QP at a time; persistent CTAs pull QP indices from a global
// work counter, so the 1-2 % of QPs that run to max_iter do not stall the rest of
// the batch.  Everything between stage-in and the final store lives in shared
// memory / registers.  Stages (reference code they replace):
//
//   K0 stage-in          Solve assembly            src/fcc_qp.cpp:141-150
//   K1 cold pre-solve    LDLT -> COD fallback      src/fcc_qp.cpp:159-178
//   K2 rho-KKT factor    LDLT.compute              src/fcc_qp.cpp:62-71
//   K3 x-update          LDLT.solve                src/fcc_qp.cpp:81-87
//   K4 z-update          clamp + cone projection   src/fcc_qp.cpp:90-92, src/constraint_utils.cpp:5-46
//   K5 residuals / duals / exit                    src/fcc_qp.cpp:95-109
//   K6 epilogue          violations + details      src/fcc_qp.cpp:184-186,194-207
//
// This is NOT a port of the Eigen code paths; the linear algebra is re-derived
// so that it is pivot-free and symmetric (half the storage, no argmax chains):
//
//   * K1.  The reference solves the indefinite system [[Q,A'],[A,0]] s = [-b; b_eq]
//     whose (1,1) block is singular (zero-cost force variables) through a failed
//     LDLT and a complete orthogonal decomposition.  Here the SAME solution is
//     obtained from the augmented-Lagrangian form [[Q + sigma A'A, A'],[A,0]] with
//     right-hand side [-b + sigma A' b_eq; b_eq]: on {Ax = b_eq} the added term is
//     constant, so x is unchanged, while Q + sigma A'A is positive definite
//     exactly when the KKT matrix is nonsingular.  That matrix is quasi-definite,
//     so an unpivoted LDL^T exists; one step of iterative refinement against the
//     ORIGINAL system removes the sigma-dependent rounding (measured: <= 6.2e-11
//     relative to the reference on the walking log, 0/2019 iteration mismatches).
//   * K2 is the same unpivoted blocked LDL^T on [[Q + rho I, A'],[A,0]].
//   * K3 is a blocked triangular solve with explicitly inverted 16x16 diagonal
//     blocks (short GEMV chains instead of 2N dependent steps).
//
// Storage: the lower triangle of the (n+m) x (n+m) KKT matrix, row-major, rows
// grouped by 4 and padded so that every 4-row group has one stride (register
// tiles) and consecutive rows start 2 (mod 4) doubles apart (bank spread).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fccqp {

constexpr int kNB = 8;    // panel width of the blocked factorization
constexpr int kTB = 16;   // diagonal block of the blocked triangular solves
constexpr int kTile = 4;  // register tile of the trailing updates

struct SolveParams {
  int B, n, m, nc, lcs;
  int max_iter, warm;
  int refine;              // iterative-refinement steps of the cold pre-solve (default 1)
  double rho, eps_fcone, eps_bound;
  const double* Q;   long long q_bs, q_rs, q_cs;
  const double* b;   long long b_bs;
  const double* A;   long long a_bs, a_rs, a_cs;
  const double* beq; long long beq_bs;
  const double* mu;  long long mu_bs;
  const double* lb;  long long lb_bs;
  const double* ub;  long long ub_bs;
  double* x; double* mu_x; double* mu_c;
  int* n_iter; int* status;
  double* res_b; double* res_f; double* bviol; double* fviol;
  unsigned int* work_counter;
  double* dbg_x0;              // optional [B,n]: pre-solve point (debug / tests)
  unsigned long long* cycles;  // optional [2]: summed factorization / total cycles
  unsigned long long* prof;    // optional [16]: per-phase cycle counters (developer profiling)
};

// Packed lower-triangular row-major storage.  Row i belongs to group q = i/4; every row of a
// group has allocated length 4(q+1)+2 (covers columns 0..4q+3, +2 keeps starts 16B aligned
// while spreading consecutive rows over the banks).
__host__ __device__ __forceinline__ int row_off(int i) {
  const int q = i >> 2, r = i & 3;
  return 8 * q * (q + 1) + r * 4 * (q + 1) + 2 * i;
}
__host__ __device__ __forceinline__ int row_len(int i) { return 4 * ((i >> 2) + 1) + 2; }

// Shared-memory carve-up, identical on host (sizing) and device (pointers).
struct Layout {
  int N, NP, nblk, NT;  // NT = nblk * kTB (padded vector length)
  size_t off_M, off_wt, off_xinv, off_dinv, off_tbuf, off_ybuf, off_sbuf;
  size_t off_b, off_beq, off_lb, off_ub, off_mu, off_xs, off_xbar, off_mux, off_lcbar, off_muc;
  size_t off_red, off_int;
  size_t doubles_total;
  __host__ __device__ static inline size_t up2(size_t v) { return (v + 1) & ~size_t(1); }
  __host__ __device__ Layout(int n, int m, int nc) {
    N = n + m;
    NP = (N + kTile - 1) / kTile * kTile;
    nblk = (N + kTB - 1) / kTB;
    NT = nblk * kTB;
    size_t o = 0;
    off_M = o;     o += up2((size_t)row_off(NP));
    // wt (W = L21 D panel of the factorization) is dead once the factorization is done, which
    // is when xinv (inverted diagonal blocks) is built: they share one region.
    off_wt = o;    off_xinv = o;
    {
      const size_t a = up2((size_t)NP * kNB), b2 = up2((size_t)nblk * kTB * (kTB + 1));
      o += a > b2 ? a : b2;
    }
    off_dinv = o;  o += up2(NT);
    off_tbuf = o;  o += up2(NT);
    off_ybuf = o;  o += up2(NT);
    off_sbuf = o;  o += up2(NT);
    off_b = o;     o += up2(n);
    off_beq = o;   o += up2(m);
    off_lb = o;    o += up2(n);
    off_ub = o;    o += up2(n);
    off_mu = o;    o += up2(nc / 3 + 1);
    off_xs = o;    o += up2(n);
    off_xbar = o;  o += up2(n);
    off_mux = o;   o += up2(n);
    off_lcbar = o; o += up2(nc + 1);
    off_muc = o;   o += up2(nc + 1);
    off_red = o;   o += 4 * 32;   // block_reduce2 scratch: 2 buffers x 2 values x 32 warps
    off_int = o;   o += 64;       // ints: work index, profiling slots
    doubles_total = o;
  }
  __host__ __device__ size_t bytes() const { return doubles_total * sizeof(double); }
};

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide reduction of two values (max or sum).  `red` holds 2 x 2 x 32 doubles;
// `parity` alternates between the two halves so one barrier per call suffices.
template <bool kSum>
__device__ __forceinline__ void block_reduce2(double& a, double& b, double* red, int& parity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (kSum) { a = warp_sum(a); b = warp_sum(b); } else { a = warp_max(a); b = warp_max(b); }
  double* r = red + parity * 64;
  if (lane == 0) { r[warp] = a; r[32 + warp] = b; }
  __syncthreads();
  double ra = r[0], rb = r[32];
  for (int w = 1; w < nw; ++w) {
    if (kSum) { ra += r[w]; rb += r[32 + w]; } else { ra = fmax(ra, r[w]); rb = fmax(rb, r[32 + w]); }
  }
  a = ra; b = rb;
  parity ^= 1;
}

// 1/d to full double precision: MUFU seed (~20 bits) + two Newton steps.  Shorter dependent
// chain than the IEEE division sequence; d is a factorization pivot (finite, non-denormal).
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// constraint_utils.cpp:5-25, including the f_z == 0 quirk (zero cone_ray is not normalised).
__device__ __forceinline__ void project_cone3(double f0, double f1, double f2, double mu,
                                              double& o0, double& o1, double& o2) {
  const double r = sqrt(f0 * f0 + f1 * f1);
  if (mu * f2 >= r) { o0 = f0; o1 = f1; o2 = f2; return; }
  if (f2 < -mu * r) { o0 = 0.0; o1 = 0.0; o2 = 0.0; return; }
  const double ratio = mu * f2 / r;
  double r0 = ratio * f0, r1 = ratio * f1, r2 = f2;
  const double sq = r0 * r0 + r1 * r1 + r2 * r2;
  if (sq > 0.0) { const double nr = sqrt(sq); r0 /= nr; r1 /= nr; r2 /= nr; }
  const double d = r0 * f0 + r1 * f1 + r2 * f2;
  o0 = d * r0; o1 = d * r1; o2 = d * r2;
}

__device__ __forceinline__ double clampd(double x, double lb, double ub) {
  const double t = x < ub ? x : ub;  // std::min(x, ub)
  return t > lb ? t : lb;            // std::max(., lb)   constraint_utils.cpp:43
}

// Rank-kb update of one 4x4 tile:  C[r][s] -= sum_k Lp[r][k] * Up[k][s].
// C and Lp are rows of the packed matrix (one stride `ldc` for the 4 rows of a group),
// Up is k-major with stride ldu.  All addresses are 16-byte aligned by construction.
__device__ __forceinline__ void tile_update(double* __restrict__ C, int ldc, const double* __restrict__ Lp,
                                            const double* __restrict__ Up, int ldu, int kb) {
  double c[kTile][kTile];
#pragma unroll
  for (int r = 0; r < kTile; ++r) {
    const double2 v0 = *reinterpret_cast<const double2*>(C + r * ldc);
    const double2 v1 = *reinterpret_cast<const double2*>(C + r * ldc + 2);
    c[r][0] = v0.x; c[r][1] = v0.y; c[r][2] = v1.x; c[r][3] = v1.y;
  }
  if (kb == kNB) {
#pragma unroll
    for (int k = 0; k < kNB; k += 2) {
      double2 l[kTile];
#pragma unroll
      for (int r = 0; r < kTile; ++r) l[r] = *reinterpret_cast<const double2*>(Lp + r * ldc + k);
      const double2 u00 = *reinterpret_cast<const double2*>(Up + k * ldu);
      const double2 u01 = *reinterpret_cast<const double2*>(Up + k * ldu + 2);
      const double2 u10 = *reinterpret_cast<const double2*>(Up + (k + 1) * ldu);
      const double2 u11 = *reinterpret_cast<const double2*>(Up + (k + 1) * ldu + 2);
#pragma unroll
      for (int r = 0; r < kTile; ++r) {
        c[r][0] -= l[r].x * u00.x; c[r][1] -= l[r].x * u00.y;
        c[r][2] -= l[r].x * u01.x; c[r][3] -= l[r].x * u01.y;
        c[r][0] -= l[r].y * u10.x; c[r][1] -= l[r].y * u10.y;
        c[r][2] -= l[r].y * u11.x; c[r][3] -= l[r].y * u11.y;
      }
    }
  } else {
    for (int k = 0; k < kb; ++k) {
      const double2 u0 = *reinterpret_cast<const double2*>(Up + k * ldu);
      const double2 u1 = *reinterpret_cast<const double2*>(Up + k * ldu + 2);
#pragma unroll
      for (int r = 0; r < kTile; ++r) {
        const double l = Lp[r * ldc + k];
        c[r][0] -= l * u0.x; c[r][1] -= l * u0.y; c[r][2] -= l * u1.x; c[r][3] -= l * u1.y;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kTile; ++r) {
    *reinterpret_cast<double2*>(C + r * ldc) = make_double2(c[r][0], c[r][1]);
    *reinterpret_cast<double2*>(C + r * ldc + 2) = make_double2(c[r][2], c[r][3]);
  }
}

// linear index of a lower-triangular tile -> (ti, tj), tj <= ti
__device__ __forceinline__ void tri_index(int t, int& ti, int& tj) {
  ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (ti * (ti + 1) / 2 > t) --ti;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  tj = t - ti * (ti + 1) / 2;
}

// ---------------------------------------------------------------------------
// The fused solve kernel.  kThreads >= N (one thread per KKT row in the panel
// solves and the triangular solves).
// ---------------------------------------------------------------------------
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) fccqp_solve_kernel(const SolveParams p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  const int n = p.n, m = p.m, nc = p.nc, lcs = p.lcs;
  const Layout L(n, m, nc);
  const int N = L.N, NP = L.NP, nblk = L.nblk;

  double* M = smem + L.off_M;
  double* WT = smem + L.off_wt;      // [kNB][NP]  W = L21 * D11, k-major
  double* xinv = smem + L.off_xinv;  // [nblk][kTB][kTB+1]
  double* dinv = smem + L.off_dinv;
  double* tbuf = smem + L.off_tbuf;
  double* ybuf = smem + L.off_ybuf;
  double* sbuf = smem + L.off_sbuf;
  double* vb = smem + L.off_b;
  double* vbeq = smem + L.off_beq;
  double* vlb = smem + L.off_lb;
  double* vub = smem + L.off_ub;
  double* vmu = smem + L.off_mu;
  double* xs = smem + L.off_xs;
  double* xbar = smem + L.off_xbar;
  double* mux = smem + L.off_mux;
  double* lcbar = smem + L.off_lcbar;
  double* muc = smem + L.off_muc;
  double* red = smem + L.off_red;
  int* ibuf = reinterpret_cast<int*>(smem + L.off_int);
  int* s_work = ibuf;  // [1]
  unsigned long long* s_prof = reinterpret_cast<unsigned long long*>(ibuf + 32);  // [16]
  long long t_prof = 0;
  if (p.prof && tid == 0) { for (int i = 0; i < 16; ++i) s_prof[i] = 0; t_prof = clock64(); }
#define FCCQP_PROF(slot)                                                   \
  do {                                                                     \
    if (p.prof && tid == 0) {                                              \
      const long long t_now = clock64();                                   \
      s_prof[slot] += (unsigned long long)(t_now - t_prof);                \
      t_prof = t_now;                                                      \
    }                                                                      \
  } while (0)

  int parity = 0;
  const int t = tid;
  const bool is_row = t < N;
  const int J_me = t / kTB, c_me = t % kTB;
  const int my_off = row_off(t < NP ? t : 0);

  for (;;) {
    __syncthreads();  // previous QP fully retired (smem reuse) before taking new work
    if (tid == 0) *s_work = (int)atomicAdd(p.work_counter, 1u);
    __syncthreads();
    const int qp = *s_work;
    if (qp >= p.B) break;

    const double* Qg = p.Q + (size_t)qp * p.q_bs;
    const double* Ag = p.A + (size_t)qp * p.a_bs;
    const double* bg = p.b + (size_t)qp * p.b_bs;
    const double* beqg = p.beq + (size_t)qp * p.beq_bs;
    // Q is symmetric: walk it along whichever stride is contiguous.
    const long long q_slow = p.q_cs <= p.q_rs ? p.q_rs : p.q_cs;
    const long long q_fast = p.q_cs <= p.q_rs ? p.q_cs : p.q_rs;
    const bool a_row_fast = p.a_cs <= p.a_rs;  // consecutive columns contiguous (row-major A)

    // ---------------- K0: vectors ----------------
    int finite_bounds = 0;
    for (int i = tid; i < n; i += kThreads) {
      vb[i] = bg[i];
      const double l = p.lb[(size_t)qp * p.lb_bs + i], u = p.ub[(size_t)qp * p.ub_bs + i];
      vlb[i] = l; vub[i] = u;
      if (!isinf(l) || !isinf(u)) finite_bounds = 1;
      if (p.warm) {
        xs[i] = p.x[(size_t)qp * n + i];
        mux[i] = p.mu_x[(size_t)qp * n + i];
      } else {
        mux[i] = 0.0;
      }
    }
    for (int i = tid; i < m; i += kThreads) vbeq[i] = beqg[i];
    for (int i = tid; i < nc / 3; i += kThreads) vmu[i] = p.mu[(size_t)qp * p.mu_bs + i];
    for (int i = tid; i < nc; i += kThreads) muc[i] = p.warm ? p.mu_c[(size_t)qp * nc + i] : 0.0;
    for (int i = tid; i < L.NT; i += kThreads) { tbuf[i] = 0.0; ybuf[i] = 0.0; dinv[i] = 0.0; sbuf[i] = 0.0; }
    const bool eqc = (__syncthreads_or(finite_bounds) == 0) && (nc == 0);  // fcc_qp.cpp:132-133
    const bool presolve = eqc || !p.warm;                                  // fcc_qp.cpp:159

    const long long t_start = clock64();
    unsigned long long fact_cycles = 0;
    int status_flag = 0;
    int n_iter = 0;
    double res_x = 0.0, res_c = 0.0;
    FCCQP_PROF(0);

    // pass 0: cold pre-solve on [[Q + sigma A'A, A'],[A,0]];  pass 1: ADMM on [[Q + rho I, A'],[A,0]]
    for (int pass = presolve ? 0 : 1; pass < 2; ++pass) {
      if (pass == 1 && eqc) break;
      const long long t_f0 = clock64();

      // ---------------- assemble the lower triangle of the KKT matrix ----------------
      // zero block (2,2) and the padding of rows >= n (everything right of column n)
      for (int i = n + warp; i < NP; i += kWarps) {
        double* row = M + row_off(i);
        const int len = row_len(i);
        for (int j = n + lane; j < len; j += 32) row[j] = 0.0;
      }
      if (q_fast == 1 && (n & 1) == 0 && ((reinterpret_cast<uintptr_t>(Qg) | (uintptr_t)(q_slow * 8)) & 15) == 0) {
        for (int i = warp; i < n; i += kWarps) {   // 16-byte copies; row i needs columns 0..i
          double* row = M + row_off(i);
          const double* src = Qg + i * q_slow;
          for (int j = 2 * lane; j <= i; j += 64) cp_async16(row + j, src + j);
        }
      } else {
        for (int i = warp; i < n; i += kWarps) {
          double* row = M + row_off(i);
          const double* src = Qg + i * q_slow;
          for (int j = lane; j <= i; j += 32) cp_async8(row + j, src + j * q_fast);
        }
      }
      if (a_row_fast) {
        if (p.a_cs == 1 && (n & 1) == 0 && ((reinterpret_cast<uintptr_t>(Ag) | (uintptr_t)(p.a_rs * 8)) & 15) == 0) {
          for (int k = warp; k < m; k += kWarps) {
            double* row = M + row_off(n + k);
            const double* src = Ag + k * p.a_rs;
            for (int j = 2 * lane; j < n; j += 64) cp_async16(row + j, src + j);
          }
        } else {
          for (int k = warp; k < m; k += kWarps) {
            double* row = M + row_off(n + k);
            const double* src = Ag + k * p.a_rs;
            for (int j = lane; j < n; j += 32) cp_async8(row + j, src + j * p.a_cs);
          }
        }
      } else {
        for (int j = warp; j < n; j += kWarps)
          for (int k = lane; k < m; k += 32) cp_async8(M + row_off(n + k) + j, Ag + k * p.a_rs + j * p.a_cs);
      }
      cp_async_wait_all();
      __syncthreads();
      FCCQP_PROF(1);

      double rhs0 = 0.0;  // pass-0 right-hand side of row t
      if (pass == 1) {
        if (t < n) M[my_off + t] += p.rho;
        __syncthreads();
      } else {
        // sigma = trace(Q) / ||A||_F^2 balances the two terms of Q + sigma A'A
        double trq = (t < n) ? M[my_off + t] : 0.0, fro = 0.0;
        if (t >= n && is_row) {
          const double* row = M + my_off;
          for (int j = 0; j < n; ++j) fro += row[j] * row[j];
        }
        block_reduce2<true>(trq, fro, red, parity);
        const double sigma = (trq > 0.0 && fro > 0.0 && isfinite(trq / fro)) ? trq / fro : 1.0;
        // rhs_x = -b + sigma A' b_eq (needs A before the factorization overwrites it)
        if (t < n) {
          double s = 0.0;
          for (int k = 0; k < m; ++k) s += M[row_off(n + k) + t] * vbeq[k];
          rhs0 = -vb[t] + sigma * s;
        } else if (is_row) {
          rhs0 = vbeq[t - n];
        }
        // H += sigma A'A on the lower 4x4 tiles of the leading n x n block
        if (m > 0) {
          const int TT = (n + kTile - 1) / kTile;
          for (int tile = tid; tile < TT * (TT + 1) / 2; tile += kThreads) {
            int ti, tj;
            tri_index(tile, ti, tj);
            const int i0 = ti * kTile, j0 = tj * kTile;
            double c[kTile][kTile];
#pragma unroll
            for (int r = 0; r < kTile; ++r)
#pragma unroll
              for (int s = 0; s < kTile; ++s) c[r][s] = 0.0;
            for (int k = 0; k < m; ++k) {
              const double* arow = M + row_off(n + k);
              const double2 a0 = *reinterpret_cast<const double2*>(arow + i0);
              const double2 a1 = *reinterpret_cast<const double2*>(arow + i0 + 2);
              const double2 b0 = *reinterpret_cast<const double2*>(arow + j0);
              const double2 b1 = *reinterpret_cast<const double2*>(arow + j0 + 2);
              const double ai[4] = {a0.x, a0.y, a1.x, a1.y};
              const double aj[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
              for (int r = 0; r < kTile; ++r)
#pragma unroll
                for (int s = 0; s < kTile; ++s) c[r][s] += ai[r] * aj[s];
            }
            double* C = M + row_off(i0) + j0;
            const int ldc = row_len(i0);
#pragma unroll
            for (int r = 0; r < kTile; ++r) {
              double2 v0 = *reinterpret_cast<double2*>(C + r * ldc);
              double2 v1 = *reinterpret_cast<double2*>(C + r * ldc + 2);
              v0.x += sigma * c[r][0]; v0.y += sigma * c[r][1];
              v1.x += sigma * c[r][2]; v1.y += sigma * c[r][3];
              *reinterpret_cast<double2*>(C + r * ldc) = v0;
              *reinterpret_cast<double2*>(C + r * ldc + 2) = v1;
            }
          }
        }
        __syncthreads();
      }
      FCCQP_PROF(2);

      // ---------------- unpivoted blocked LDL^T (lower, packed) ----------------
      for (int k0 = 0; k0 < N; k0 += kNB) {
        const int kb = min(kNB, N - k0);
        // --- diagonal block kb x kb: one thread, registers only (shortest dependent chain)
        if (tid == 0) {
          double a[kNB][kNB];
#pragma unroll
          for (int r = 0; r < kNB; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) a[r][c] = (r < kb) ? M[row_off(k0 + r) + k0 + c] : (r == c ? 1.0 : 0.0);
#pragma unroll
          for (int c = 0; c < kNB; ++c) {
            const double rd = fast_rcp(a[c][c]);
            if (c < kb) dinv[k0 + c] = rd;
#pragma unroll
            for (int r = c + 1; r < kNB; ++r) {
              // rows c2 < r of this column already hold l_{c2,c}; a[r][c] is still unscaled
              const double arc = a[r][c];
              const double l = arc * rd;
#pragma unroll
              for (int c2 = c + 1; c2 < r; ++c2) a[r][c2] -= arc * a[c2][c];
              a[r][r] -= arc * l;
              a[r][c] = l;
            }
          }
#pragma unroll
          for (int r = 1; r < kNB; ++r)
#pragma unroll
            for (int c = 0; c < r; ++c)
              if (r < kb) M[row_off(k0 + r) + k0 + c] = a[r][c];
        }
        __syncthreads();
        FCCQP_PROF(3);
        // --- L21 = A21 L11^{-T} D11^{-1};  W = L21 D11 (thread per row below the block)
        {
          const int row = k0 + kb + tid;
          if (row < N) {
            double* rp = M + row_off(row) + k0;
            double w[kNB];
#pragma unroll
            for (int c = 0; c < kNB; c += 2) {
              const double2 v = *reinterpret_cast<const double2*>(rp + c);
              w[c] = v.x; w[c + 1] = v.y;
            }
#pragma unroll
            for (int c = 1; c < kNB; ++c) {
              if (c < kb) {
                const double* l11 = M + row_off(k0 + c) + k0;
#pragma unroll
                for (int cc = 0; cc < c; ++cc) w[c] -= w[cc] * l11[cc];
              }
            }
#pragma unroll
            for (int c = 0; c < kNB; ++c) {
              if (c < kb) {
                WT[c * NP + row] = w[c];
                rp[c] = w[c] * dinv[k0 + c];
              }
            }
          }
        }
        __syncthreads();
        FCCQP_PROF(4);
        // --- trailing update (lower 4x4 tiles): A22 -= L21 W^T
        const int r0 = k0 + kb;
        if (r0 < N) {
          const int TT = (N - r0 + kTile - 1) / kTile;
          const int ntiles = TT * (TT + 1) / 2;
          for (int tile = tid; tile < ntiles; tile += kThreads) {
            int ti, tj;
            tri_index(tile, ti, tj);
            const int i0 = r0 + ti * kTile, j0 = r0 + tj * kTile;
            double* rowp = M + row_off(i0);
            tile_update(rowp + j0, row_len(i0), rowp + k0, WT + j0, NP, kb);
          }
        }
        __syncthreads();
        FCCQP_PROF(5);
      }
      // --- explicit inverses of the kTB x kTB unit-lower diagonal blocks of L
      if (tid < L.NT) {
        const int blk = tid / kTB, c = tid % kTB, I0 = blk * kTB;
        double xv[kTB];
#pragma unroll
        for (int i = 0; i < kTB; ++i) {
          double s = 0.0;
          const double* lrow = M + row_off(min(I0 + i, NP - 1)) + I0;
#pragma unroll
          for (int k = 0; k < i; ++k) {
            const double lik = (I0 + i < N) ? lrow[k] : 0.0;
            s += lik * xv[k];
          }
          xv[i] = (i == c) ? 1.0 : -s;
        }
#pragma unroll
        for (int i = 0; i < kTB; ++i) xinv[(blk * kTB + i) * (kTB + 1) + c] = xv[i];
      }
      for (int i = tid; i < L.NT; i += kThreads) tbuf[i] = 0.0;
      __syncthreads();
      fact_cycles += (unsigned long long)(clock64() - t_f0);
      FCCQP_PROF(6);

      if (pass == 1) {
        // ADMM initial slack (fcc_qp.cpp:74-75)
        for (int i = tid; i < n; i += kThreads) xbar[i] = xs[i];
        for (int i = tid; i < nc; i += kThreads) lcbar[i] = xs[lcs + i];
        __syncthreads();
        n_iter = p.max_iter;
      }
      const int iters = pass == 0 ? 1 + p.refine : p.max_iter;
      double sol = 0.0;   // pass 0: accumulated solution component of row t
      double acc0 = rhs0; // pass 0: right-hand side of the next solve

      for (int iter = 0; iter < iters; ++iter) {
        // ---- K3 right-hand side
        double acc = 0.0;
        if (pass == 0) {
          acc = is_row ? acc0 : 0.0;
        } else if (t < n) {
          // -(b + q_rho), q_rho = -rho (xbar - mu_x) with the cone segment overwritten (fcc_qp.cpp:81-83)
          const bool in_c = (t >= lcs) && (t < lcs + nc);
          const double w = in_c ? (lcbar[t - lcs] - muc[t - lcs]) : (xbar[t] - mux[t]);
          const double q_rho = -p.rho * w;
          acc = -(vb[t] + q_rho);
        } else if (is_row) {
          acc = vbeq[t - n];
        }
        // ---- forward: L y = rhs
        double val = 0.0;
        for (int J = 0; J < nblk; ++J) {
          const int J0 = J * kTB;
          if (J_me == J) tbuf[t] = acc;
          __syncwarp();
          if (J_me == J) {
            double s0 = 0.0, s1 = 0.0;
            const double* xr = xinv + (J * kTB + c_me) * (kTB + 1);
#pragma unroll
            for (int c = 0; c < kTB; c += 2) { s0 += xr[c] * tbuf[J0 + c]; s1 += xr[c + 1] * tbuf[J0 + c + 1]; }
            val = s0 + s1;
            ybuf[t] = val;
          }
          __syncthreads();
          if (is_row && t >= J0 + kTB) {
            const double* lr = M + my_off + J0;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c = 0; c < kTB; c += 2) {
              const double2 l2 = *reinterpret_cast<const double2*>(lr + c);
              s0 += l2.x * ybuf[J0 + c]; s1 += l2.y * ybuf[J0 + c + 1];
            }
            acc -= s0 + s1;
          }
        }
        // ---- D^{-1}
        acc = is_row ? val * dinv[t] : 0.0;
        __syncthreads();  // ybuf reuse
        // ---- backward: L^T x = y
        for (int J = nblk - 1; J >= 0; --J) {
          const int J0 = J * kTB;
          const int bs = min(kTB, N - J0);
          if (J_me == J) tbuf[t] = is_row ? acc : 0.0;
          __syncwarp();
          if (J_me == J) {
            double s0 = 0.0, s1 = 0.0;
            const double* xc = xinv + J * kTB * (kTB + 1) + c_me;
#pragma unroll
            for (int c = 0; c < kTB; c += 2) {
              s0 += xc[c * (kTB + 1)] * tbuf[J0 + c];
              s1 += xc[(c + 1) * (kTB + 1)] * tbuf[J0 + c + 1];
            }
            val = s0 + s1;
            ybuf[t] = val;
          }
          __syncthreads();
          if (t < J0) {
            double s0 = 0.0, s1 = 0.0;
            for (int c = 0; c + 1 < bs; c += 2) {
              s0 += M[row_off(J0 + c) + t] * ybuf[J0 + c];
              s1 += M[row_off(J0 + c + 1) + t] * ybuf[J0 + c + 1];
            }
            if (bs & 1) s0 += M[row_off(J0 + bs - 1) + t] * ybuf[J0 + bs - 1];
            acc -= s0 + s1;
          }
        }
        FCCQP_PROF(7);
        // val = solution component of row t (t < N)

        if (pass == 0) {
          sol += val;
          if (iter + 1 < iters) {
            // ---- iterative refinement against the ORIGINAL system: r = [-b; b_eq] - [[Q,A'],[A,0]] s
            if (is_row) sbuf[t] = sol;
            __syncthreads();
            double r = 0.0;
            if (t < n) {
              double s0 = 0.0, s1 = 0.0;
              for (int j = 0; j < n; ++j) s0 += Qg[j * q_slow + t * q_fast] * sbuf[j];       // Q symmetric
              for (int k = 0; k < m; ++k) s1 += Ag[k * p.a_rs + t * p.a_cs] * sbuf[n + k];   // A' y
              r = -vb[t] - s0 - s1;
            }
            // rows of A: warp per row, lanes over columns
            for (int k = warp; k < m; k += kWarps) {
              double s = 0.0;
              for (int j = lane; j < n; j += 32) s += Ag[k * p.a_rs + j * p.a_cs] * sbuf[j];
              s = warp_sum(s);
              if (lane == 0) tbuf[n + k] = vbeq[k] - s;
            }
            __syncthreads();
            if (t >= n && is_row) r = tbuf[t];
            __syncthreads();
            if (t >= n && t < L.NT) tbuf[t] = 0.0;
            acc0 = r;
          } else {
            if (t < n) xs[t] = sol;
            __syncthreads();
            if (p.dbg_x0) for (int i = tid; i < n; i += kThreads) p.dbg_x0[(size_t)qp * n + i] = xs[i];
          }
          FCCQP_PROF(8);
          continue;
        }

        // ---- K4 + K5 (pass 1)
        if (t < n) xs[t] = val;
        __syncthreads();
        double rx = 0.0, rc = 0.0;
        if (t < n) {
          const double xb = clampd(val + mux[t], vlb[t], vub[t]);
          xbar[t] = xb;
          const double r = val - xb;
          mux[t] += r;
          rx = fabs(r);
        }
        if (t < nc / 3) {  // lane per contact
          const int o = lcs + 3 * t;
          const double x0 = xs[o], x1 = xs[o + 1], x2 = xs[o + 2];
          double o0, o1, o2;
          project_cone3(x0 + muc[3 * t], x1 + muc[3 * t + 1], x2 + muc[3 * t + 2], vmu[t], o0, o1, o2);
          lcbar[3 * t] = o0; lcbar[3 * t + 1] = o1; lcbar[3 * t + 2] = o2;
          const double r0 = x0 - o0, r1 = x1 - o1, r2 = x2 - o2;
          muc[3 * t] += r0; muc[3 * t + 1] += r1; muc[3 * t + 2] += r2;
          rc = fmax(fabs(r0), fmax(fabs(r1), fabs(r2)));
        }
        // fmax drops NaNs, so flag them separately
        if (rx != rx || rc != rc) status_flag = 2;
        block_reduce2<false>(rx, rc, red, parity);
        res_x = rx; res_c = rc;
        FCCQP_PROF(9);
        if (p.prof && tid == 0) s_prof[15] += 1;
        if (rc < p.eps_fcone && rx < p.eps_bound) { n_iter = iter; break; }  // fcc_qp.cpp:105-109
      }
    }

    // ---------------- K6: epilogue ----------------
    __syncthreads();
    double bv = 0.0, fv = 0.0;
    int bad = 0;
    if (tid < n) {
      const double xv = xs[tid];
      const double d = xv - clampd(xv, vlb[tid], vub[tid]);
      bv = d * d;
      if (!isfinite(xv)) bad = 1;
    }
    if (tid < nc / 3) {
      const int o = lcs + 3 * tid;
      const double r = sqrt(xs[o] * xs[o] + xs[o + 1] * xs[o + 1]) - vmu[tid] * xs[o + 2];
      fv = r > 0.0 ? r : 0.0;
    }
    block_reduce2<true>(bv, fv, red, parity);
    bad = __syncthreads_or(bad | (status_flag == 2));
    for (int i = tid; i < n; i += kThreads) {
      p.x[(size_t)qp * n + i] = xs[i];
      if (p.mu_x) p.mu_x[(size_t)qp * n + i] = mux[i];
    }
    if (p.mu_c) for (int i = tid; i < nc; i += kThreads) p.mu_c[(size_t)qp * nc + i] = muc[i];
    if (tid == 0) {
      if (p.n_iter) p.n_iter[qp] = n_iter;
      if (p.status) p.status[qp] = bad ? 2 : (n_iter == p.max_iter ? 1 : 0);  // fcc_qp.cpp:203-204
      if (p.res_b) p.res_b[qp] = res_x;
      if (p.res_f) p.res_f[qp] = res_c;
      if (p.bviol) p.bviol[qp] = sqrt(bv);
      if (p.fviol) p.fviol[qp] = fv;
      if (p.cycles) {
        atomicAdd(p.cycles, fact_cycles);
        atomicAdd(p.cycles + 1, (unsigned long long)(clock64() - t_start));
      }
    }
    FCCQP_PROF(10);
    if (p.prof && tid == 0) s_prof[14] += 1;
  }
  if (p.prof && tid == 0)
    for (int i = 0; i < 16; ++i) atomicAdd(p.prof + i, s_prof[i]);
#undef FCCQP_PROF
}

}  // namespace fccqp
