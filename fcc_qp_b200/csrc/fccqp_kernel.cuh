// fccqp_kernel.cuh -- sm_100a device code of the batched FCCQP solve.
//
// One CTA solves one QP at a time (persistent CTAs pull QP indices from a
// global work counter, so the 1-2 % of QPs that run to max_iter do not stall
// the rest of the batch).  Everything between stage-in and the final store
// lives in shared memory / registers:
//
//   K0 stage-in          Solve assembly            src/fcc_qp.cpp:141-150
//   K1 cold pre-solve    LDLT -> COD fallback      src/fcc_qp.cpp:159-178
//   K2 rho-KKT factor    LDLT.compute              src/fcc_qp.cpp:62-71
//   K3 x-update          LDLT.solve                src/fcc_qp.cpp:81-87
//   K4 z-update          clamp + cone projection   src/fcc_qp.cpp:90-92, src/constraint_utils.cpp:5-46
//   K5 residuals / duals / exit                    src/fcc_qp.cpp:95-109
//   K6 epilogue          violations + details      src/fcc_qp.cpp:184-186,194-207
//
// This is NOT a port of the Eigen code paths:
//   * K1 solves the indefinite, singular-(1,1)-block KKT system by blocked
//     Gaussian elimination with partial pivoting on the augmented matrix
//     [K | rhs] (the reference gets there through a failed LDLT and a
//     complete orthogonal decomposition); same unique solution when K is
//     nonsingular.
//   * K2 is an unpivoted blocked LDL^T of the quasi-definite rho-KKT matrix
//     (positive pivots for Q + rho I, negative for the Schur complement), which
//     exists for every symmetric permutation, so no pivot search is needed.
//   * K3 is a blocked triangular solve with explicitly inverted 16x16 diagonal
//     blocks (two short GEMV chains instead of 2N dependent steps).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fccqp {

constexpr int kNB = 8;    // panel width of the blocked factorizations
constexpr int kTB = 16;   // diagonal block of the blocked triangular solves
constexpr int kTile = 4;  // register tile of the trailing updates

struct SolveParams {
  int B, n, m, nc, lcs;
  int max_iter, warm;
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
  double* gscratch;        // per-CTA KKT matrix slab when it does not fit in shared memory
  long long gscratch_stride;
  double* dbg_x0;          // optional [B,n]: pre-solve point (debug / tests)
  unsigned long long* cycles;  // optional [2]: summed factorization / total cycles
  unsigned long long* prof;    // optional [16]: per-phase cycle counters (developer profiling)
};

// Shared-memory carve-up, identical on host (sizing) and device (pointers).
struct Layout {
  int N, NP, LD, nblk, NT;  // NT = nblk * kTB (padded vector length)
  size_t off_M, off_pb, off_xinv, off_dinv, off_tbuf, off_ybuf;
  size_t off_b, off_beq, off_lb, off_ub, off_mu, off_xs, off_xbar, off_mux, off_lcbar, off_muc;
  size_t off_red, off_int;
  size_t doubles_total;   // excluding M when M lives in global memory
  size_t m_doubles;
  __host__ __device__ static inline size_t up2(size_t v) { return (v + 1) & ~size_t(1); }
  __host__ __device__ Layout(int n, int m, int nc, bool m_in_smem) {
    N = n + m;
    NP = (N + kTile - 1) / kTile * kTile;
    LD = (N + 1 + 3) / 4 * 4 + 2;  // even (16B rows for vector access), LD/2 odd
    nblk = (N + kTB - 1) / kTB;
    NT = nblk * kTB;
    m_doubles = (size_t)NP * LD;
    size_t o = 0;
    off_M = o;     if (m_in_smem) o += up2(m_doubles);
    // pb (LU panel exchange / LDL^T W panel) is dead once the factorization is done, which is
    // when xinv (inverted diagonal blocks) is built: they share one region.
    off_pb = o;    off_xinv = o;
    {
      const size_t a = up2((size_t)2 * NP * kNB), b2 = up2((size_t)nblk * kTB * (kTB + 1));
      o += a > b2 ? a : b2;
    }
    off_dinv = o;  o += up2(NT);
    off_tbuf = o;  o += up2(NT);
    off_ybuf = o;  o += up2(NT);
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
    off_red = o;   o += 4 * 32;   // reduction scratch (2 buffers x 2 values x 32 warps... see block_max2)
    off_int = o;   o += 64;       // ints: argmax indices, pivots, work index
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

// Rank-kb update of one 4x4 tile:  C[i0..i0+3][j0..j0+3] -= sum_k Lrow[i][k] * Urow[k][j].
// Lp points at M[i0][k0] (row stride LD), Up at the k-th row of the "U" operand at column j0
// (row stride ldu).  All addresses are 16-byte aligned by construction.
template <typename T>
__device__ __forceinline__ void tile_update(T* __restrict__ C, int LD, const T* __restrict__ Lp,
                                            const T* __restrict__ Up, int ldu, int kb) {
  T c[kTile][kTile];
#pragma unroll
  for (int r = 0; r < kTile; ++r) {
    const double2 v0 = *reinterpret_cast<const double2*>(C + (size_t)r * LD);
    const double2 v1 = *reinterpret_cast<const double2*>(C + (size_t)r * LD + 2);
    c[r][0] = v0.x; c[r][1] = v0.y; c[r][2] = v1.x; c[r][3] = v1.y;
  }
  if (kb == kNB) {
#pragma unroll
    for (int k = 0; k < kNB; k += 2) {
      double2 l[kTile];
#pragma unroll
      for (int r = 0; r < kTile; ++r) l[r] = *reinterpret_cast<const double2*>(Lp + (size_t)r * LD + k);
      const double2 u00 = *reinterpret_cast<const double2*>(Up + (size_t)k * ldu);
      const double2 u01 = *reinterpret_cast<const double2*>(Up + (size_t)k * ldu + 2);
      const double2 u10 = *reinterpret_cast<const double2*>(Up + (size_t)(k + 1) * ldu);
      const double2 u11 = *reinterpret_cast<const double2*>(Up + (size_t)(k + 1) * ldu + 2);
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
      const double2 u0 = *reinterpret_cast<const double2*>(Up + (size_t)k * ldu);
      const double2 u1 = *reinterpret_cast<const double2*>(Up + (size_t)k * ldu + 2);
#pragma unroll
      for (int r = 0; r < kTile; ++r) {
        const T l = Lp[(size_t)r * LD + k];
        c[r][0] -= l * u0.x; c[r][1] -= l * u0.y; c[r][2] -= l * u1.x; c[r][3] -= l * u1.y;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kTile; ++r) {
    *reinterpret_cast<double2*>(C + (size_t)r * LD) = make_double2(c[r][0], c[r][1]);
    *reinterpret_cast<double2*>(C + (size_t)r * LD + 2) = make_double2(c[r][2], c[r][3]);
  }
}

// ---------------------------------------------------------------------------
// The fused solve kernel.  kThreads >= N (one thread per KKT row in the panel
// factorizations and the triangular solves).
// ---------------------------------------------------------------------------
template <int kThreads, bool kGlobalM>
__global__ void __launch_bounds__(kThreads) fccqp_solve_kernel(const SolveParams p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  constexpr int kRowsPerLane = kThreads / 32;  // rows per lane in the warp-0 back substitution
  const int n = p.n, m = p.m, nc = p.nc, lcs = p.lcs;
  const Layout L(n, m, nc, !kGlobalM);
  const int N = L.N, NP = L.NP, LD = L.LD, nblk = L.nblk;

  double* M = kGlobalM ? (p.gscratch + (size_t)blockIdx.x * p.gscratch_stride) : (smem + L.off_M);
  double* pb = smem + L.off_pb;
  double* xinv = smem + L.off_xinv;
  double* dinv = smem + L.off_dinv;
  double* tbuf = smem + L.off_tbuf;
  double* ybuf = smem + L.off_ybuf;
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
  int* widx = ibuf;           // [2][32] argmax row per warp
  int* piv = ibuf + 64;       // [kNB]
  int* s_work = ibuf + 80;    // [1]
  unsigned long long* s_prof = reinterpret_cast<unsigned long long*>(ibuf + 96);  // [16]
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
  // `red` is 4*32 doubles: block_reduce2 uses it as 2 x 64.  The LU argmax never overlaps a
  // block_reduce2 call in time (barriers in between), so it reuses red[0..64) as wmax[2][32].
  double* wmax = red;

  int parity = 0;
  unsigned long long fact_cycles = 0;

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
    for (int i = tid; i < L.NT; i += kThreads) { tbuf[i] = 0.0; ybuf[i] = 0.0; dinv[i] = 0.0; }
    const bool eqc = (__syncthreads_or(finite_bounds) == 0) && (nc == 0);  // fcc_qp.cpp:132-133
    const bool presolve = eqc || !p.warm;                                  // fcc_qp.cpp:159

    const long long t_start = clock64();
    int status_flag = 0;
    FCCQP_PROF(0);

    // ---------------- K1: cold pre-solve (blocked LU, partial pivoting, augmented rhs) ----------
    if (presolve) {
      for (int i = tid; i < NP * LD / 2; i += kThreads)
        reinterpret_cast<double2*>(M)[i] = make_double2(0.0, 0.0);
      __syncthreads();
      for (int i = warp; i < n; i += kWarps)
        for (int j = lane; j < n; j += 32) M[(size_t)i * LD + j] = Qg[i * q_slow + j * q_fast];
      if (a_row_fast) {
        for (int i = warp; i < m; i += kWarps)
          for (int j = lane; j < n; j += 32) {
            const double v = Ag[i * p.a_rs + j * p.a_cs];
            M[(size_t)(n + i) * LD + j] = v;
            M[(size_t)j * LD + n + i] = v;
          }
      } else {
        for (int j = warp; j < n; j += kWarps)
          for (int i = lane; i < m; i += 32) {
            const double v = Ag[i * p.a_rs + j * p.a_cs];
            M[(size_t)(n + i) * LD + j] = v;
            M[(size_t)j * LD + n + i] = v;
          }
      }
      for (int i = tid; i < N; i += kThreads) M[(size_t)i * LD + N] = i < n ? -vb[i] : vbeq[i - n];
      __syncthreads();
      FCCQP_PROF(1);

      for (int k0 = 0; k0 < N; k0 += kNB) {
        const int kb = min(kNB, N - k0);
        // --- panel: thread t owns physical row k0 + t
        const int t = tid;
        const bool have_row = (k0 + t) < N;
        double a[kNB];
#pragma unroll
        for (int c = 0; c < kNB; ++c)
          a[c] = (have_row && c < kb) ? M[(size_t)(k0 + t) * LD + k0 + c] : 0.0;
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
          if (c < kb) {  // uniform
            const int buf = c & 1;
            double* pbb = pb + (size_t)buf * NP * kNB;
            if (have_row) {
#pragma unroll
              for (int cc = 0; cc < kNB; cc += 2)
                *reinterpret_cast<double2*>(pbb + (size_t)t * kNB + cc) = make_double2(a[cc], a[cc + 1]);
            }
            double v = (have_row && t >= c) ? fabs(a[c]) : -1.0;
            int vi = t;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const double ov = __shfl_xor_sync(0xffffffffu, v, o);
              const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
              if (ov > v || (ov == v && oi < vi)) { v = ov; vi = oi; }
            }
            if (lane == 0) { wmax[buf * 32 + warp] = v; widx[buf * 32 + warp] = vi; }
            __syncthreads();
            double best = wmax[buf * 32];
            int pt = widx[buf * 32];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) {
              const double ov = wmax[buf * 32 + w];
              const int oi = widx[buf * 32 + w];
              if (ov > best || (ov == best && oi < pt)) { best = ov; pt = oi; }
            }
            if (!(best > 0.0)) { status_flag = 2; }  // singular (or NaN) column
            if (tid == 0) piv[c] = pt;
            double prow[kNB];
#pragma unroll
            for (int cc = 0; cc < kNB; cc += 2) {
              const double2 v2 = *reinterpret_cast<const double2*>(pbb + (size_t)pt * kNB + cc);
              prow[cc] = v2.x; prow[cc + 1] = v2.y;
            }
            if (t == pt && pt != c) {
#pragma unroll
              for (int cc = 0; cc < kNB; cc += 2) {
                const double2 v2 = *reinterpret_cast<const double2*>(pbb + (size_t)c * kNB + cc);
                a[cc] = v2.x; a[cc + 1] = v2.y;
              }
            }
            if (t == c) {
#pragma unroll
              for (int cc = 0; cc < kNB; ++cc) a[cc] = prow[cc];
            }
            if (have_row && t > c) {
              const double l = a[c] / prow[c];
              a[c] = l;
#pragma unroll
              for (int cc = c + 1; cc < kNB; ++cc) a[cc] -= l * prow[cc];
            }
          }
        }
        if (have_row) {
#pragma unroll
          for (int c = 0; c < kNB; ++c)
            if (c < kb) M[(size_t)(k0 + t) * LD + k0 + c] = a[c];
        }
        __syncthreads();
        FCCQP_PROF(2);
        // --- row swaps + U12 = L11^{-1} A12 on the columns right of the panel (incl. rhs column N)
        for (int j = k0 + kb + tid; j <= N; j += kThreads) {
          for (int c = 0; c < kb; ++c) {
            const int pt = piv[c];
            if (pt != c) {
              const double t0 = M[(size_t)(k0 + c) * LD + j];
              M[(size_t)(k0 + c) * LD + j] = M[(size_t)(k0 + pt) * LD + j];
              M[(size_t)(k0 + pt) * LD + j] = t0;
            }
          }
          double v[kNB];
#pragma unroll
          for (int c = 0; c < kNB; ++c) v[c] = c < kb ? M[(size_t)(k0 + c) * LD + j] : 0.0;
#pragma unroll
          for (int c = 1; c < kNB; ++c) {
            if (c < kb) {
#pragma unroll
              for (int cc = 0; cc < c; ++cc) v[c] -= M[(size_t)(k0 + c) * LD + k0 + cc] * v[cc];
            }
          }
#pragma unroll
          for (int c = 1; c < kNB; ++c)
            if (c < kb) M[(size_t)(k0 + c) * LD + j] = v[c];
        }
        __syncthreads();
        FCCQP_PROF(3);
        // --- trailing update A22 -= L21 U12 (4x4 register tiles)
        const int r0 = k0 + kb;
        if (r0 < N) {
          const int TR = (N - r0 + kTile - 1) / kTile;
          const int TC = (N + 1 - r0 + kTile - 1) / kTile;
          for (int tile = tid; tile < TR * TC; tile += kThreads) {
            const int ti = tile / TC, tj = tile - ti * TC;
            const int i0 = r0 + ti * kTile, j0 = r0 + tj * kTile;
            tile_update<double>(M + (size_t)i0 * LD + j0, LD, M + (size_t)i0 * LD + k0,
                                M + (size_t)k0 * LD + j0, LD, kb);
          }
        }
        __syncthreads();
        FCCQP_PROF(4);
      }
      // --- back substitution U x = y (warp 0; lane holds rows lane, lane+32, ...)
      for (int i = tid; i < N; i += kThreads) dinv[i] = 1.0 / M[(size_t)i * LD + i];
      __syncthreads();
      if (warp == 0) {
        double yl[kRowsPerLane];
#pragma unroll
        for (int q = 0; q < kRowsPerLane; ++q) {
          const int r = lane + 32 * q;
          yl[q] = r < N ? M[(size_t)r * LD + N] : 0.0;
        }
        for (int i = N - 1; i >= 0; --i) {
          const int qi = i >> 5;
          double yi = 0.0;
#pragma unroll
          for (int q = 0; q < kRowsPerLane; ++q) yi = (q == qi) ? yl[q] : yi;
          const double xi = __shfl_sync(0xffffffffu, yi * dinv[i], i & 31);
#pragma unroll
          for (int q = 0; q < kRowsPerLane; ++q) {
            const int r = lane + 32 * q;
            if (r < i) yl[q] -= M[(size_t)r * LD + i] * xi;
            else if (r == i) yl[q] = xi;
          }
        }
#pragma unroll
        for (int q = 0; q < kRowsPerLane; ++q) {
          const int r = lane + 32 * q;
          if (r < n) xs[r] = yl[q];
        }
      }
      __syncthreads();
      if (p.dbg_x0) for (int i = tid; i < n; i += kThreads) p.dbg_x0[(size_t)qp * n + i] = xs[i];
      FCCQP_PROF(5);
    }

    int n_iter = 0;
    double res_x = 0.0, res_c = 0.0;

    if (!eqc) {
      // ---------------- K2: rho-KKT, unpivoted blocked LDL^T (lower) ----------------
      const long long t_f0 = clock64();
      for (int i = tid; i < NP * LD / 2; i += kThreads)
        reinterpret_cast<double2*>(M)[i] = make_double2(0.0, 0.0);
      __syncthreads();
      for (int i = warp; i < n; i += kWarps)
        for (int j = lane; j <= i; j += 32)
          M[(size_t)i * LD + j] = Qg[i * q_slow + j * q_fast] + (i == j ? p.rho : 0.0);
      if (a_row_fast) {
        for (int i = warp; i < m; i += kWarps)
          for (int j = lane; j < n; j += 32) M[(size_t)(n + i) * LD + j] = Ag[i * p.a_rs + j * p.a_cs];
      } else {
        for (int j = warp; j < n; j += kWarps)
          for (int i = lane; i < m; i += 32) M[(size_t)(n + i) * LD + j] = Ag[i * p.a_rs + j * p.a_cs];
      }
      __syncthreads();
      FCCQP_PROF(6);

      double* WT = pb;  // [kNB][NP]  W = L21 * D11, k-major
      for (int k0 = 0; k0 < N; k0 += kNB) {
        const int kb = min(kNB, N - k0);
        // --- diagonal block kb x kb by warp 0 (lane t holds row k0+t), shuffles only
        if (warp == 0) {
          double a[kNB];
#pragma unroll
          for (int c = 0; c < kNB; ++c)
            a[c] = (lane < kb && c <= lane) ? M[(size_t)(k0 + lane) * LD + k0 + c] : 0.0;
#pragma unroll
          for (int c = 0; c < kNB; ++c) {
            const double dc = __shfl_sync(0xffffffffu, a[c], c);
            const double colv = a[c];
            const double l = colv / dc;
#pragma unroll
            for (int c2 = c + 1; c2 < kNB; ++c2) {
              const double v = __shfl_sync(0xffffffffu, colv, c2);
              if (lane >= c2) a[c2] -= l * v;
            }
            if (lane > c) a[c] = l;
          }
          if (lane < kb) {
#pragma unroll
            for (int c = 0; c < kNB; ++c) {
              if (c < lane) M[(size_t)(k0 + lane) * LD + k0 + c] = a[c];
              if (c == lane) dinv[k0 + lane] = 1.0 / a[c];
            }
          }
        }
        __syncthreads();
        FCCQP_PROF(7);
        // --- L21 = A21 L11^{-T} D11^{-1};  W = L21 D11 (thread per row below the block)
        {
          const int row = k0 + kb + tid;
          if (row < N) {
            double w[kNB];
#pragma unroll
            for (int c = 0; c < kNB; ++c) w[c] = c < kb ? M[(size_t)row * LD + k0 + c] : 0.0;
#pragma unroll
            for (int c = 1; c < kNB; ++c) {
              if (c < kb) {
#pragma unroll
                for (int cc = 0; cc < c; ++cc) w[c] -= w[cc] * M[(size_t)(k0 + c) * LD + k0 + cc];
              }
            }
#pragma unroll
            for (int c = 0; c < kNB; ++c) {
              if (c < kb) {
                WT[(size_t)c * NP + row] = w[c];
                M[(size_t)row * LD + k0 + c] = w[c] * dinv[k0 + c];
              }
            }
          }
          // rows NP-padding of WT columns must not hold NaNs that reach real data: they only
          // feed padded tile rows/cols, which are never read back.
        }
        __syncthreads();
        FCCQP_PROF(8);
        // --- trailing update (lower tiles): A22 -= L21 W^T
        const int r0 = k0 + kb;
        if (r0 < N) {
          const int TT = (N - r0 + kTile - 1) / kTile;
          const int ntiles = TT * (TT + 1) / 2;
          for (int tile = tid; tile < ntiles; tile += kThreads) {
            int ti = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
            while (ti * (ti + 1) / 2 > tile) --ti;
            while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
            const int tj = tile - ti * (ti + 1) / 2;
            const int i0 = r0 + ti * kTile, j0 = r0 + tj * kTile;
            tile_update<double>(M + (size_t)i0 * LD + j0, LD, M + (size_t)i0 * LD + k0,
                                WT + j0, NP, kb);
          }
        }
        __syncthreads();
        FCCQP_PROF(9);
      }
      // --- explicit inverses of the kTB x kTB unit-lower diagonal blocks of L
      if (tid < L.NT) {
        const int blk = tid / kTB, c = tid % kTB, I0 = blk * kTB;
        double xv[kTB];
#pragma unroll
        for (int i = 0; i < kTB; ++i) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < i; ++k) {
            const double lik = (I0 + i < N) ? M[(size_t)(I0 + i) * LD + I0 + k] : 0.0;
            s += lik * xv[k];
          }
          xv[i] = (i == c) ? 1.0 : -s;
        }
#pragma unroll
        for (int i = 0; i < kTB; ++i) xinv[((size_t)blk * kTB + i) * (kTB + 1) + c] = xv[i];
      }
      for (int i = tid; i < L.NT; i += kThreads) tbuf[i] = 0.0;
      __syncthreads();
      fact_cycles += (unsigned long long)(clock64() - t_f0);
      FCCQP_PROF(10);

      // ---------------- ADMM loop (fcc_qp.cpp:74-110) ----------------
      for (int i = tid; i < n; i += kThreads) xbar[i] = xs[i];
      for (int i = tid; i < nc; i += kThreads) lcbar[i] = xs[lcs + i];
      __syncthreads();

      n_iter = p.max_iter;
      const int t = tid;
      const bool is_row = t < N;
      const int J_me = t / kTB, c_me = t % kTB;
      for (int iter = 0; iter < p.max_iter; ++iter) {
        // K3 rhs: -(b + q_rho), q_rho = -rho (xbar - mu_x) with the cone segment overwritten
        double acc = 0.0;
        if (t < n) {
          const bool in_c = (t >= lcs) && (t < lcs + nc);
          const double w = in_c ? (lcbar[t - lcs] - muc[t - lcs]) : (xbar[t] - mux[t]);
          const double q_rho = -p.rho * w;
          acc = -(vb[t] + q_rho);
        } else if (is_row) {
          acc = vbeq[t - n];
        }
        // forward: L y = rhs
        double val = 0.0;
        for (int J = 0; J < nblk; ++J) {
          const int J0 = J * kTB;
          if (J_me == J) tbuf[t] = acc;
          __syncwarp();
          if (J_me == J) {
            double s0 = 0.0, s1 = 0.0;
            const double* xr = xinv + ((size_t)J * kTB + c_me) * (kTB + 1);
#pragma unroll
            for (int c = 0; c < kTB; c += 2) { s0 += xr[c] * tbuf[J0 + c]; s1 += xr[c + 1] * tbuf[J0 + c + 1]; }
            val = s0 + s1;
            ybuf[t] = val;
          }
          __syncthreads();
          if (is_row && t >= J0 + kTB) {
            const double* lr = M + (size_t)t * LD + J0;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c = 0; c < kTB; c += 2) {
              const double2 l2 = *reinterpret_cast<const double2*>(lr + c);
              s0 += l2.x * ybuf[J0 + c]; s1 += l2.y * ybuf[J0 + c + 1];
            }
            acc -= s0 + s1;
          }
        }
        // D^{-1}
        acc = is_row ? val * dinv[t] : 0.0;
        __syncthreads();  // ybuf reuse
        // backward: L^T x = y
        for (int J = nblk - 1; J >= 0; --J) {
          const int J0 = J * kTB;
          const int bs = min(kTB, N - J0);
          if (J_me == J) tbuf[t] = is_row ? acc : 0.0;
          __syncwarp();
          if (J_me == J) {
            double s0 = 0.0, s1 = 0.0;
            const double* xc = xinv + (size_t)J * kTB * (kTB + 1) + c_me;
#pragma unroll
            for (int c = 0; c < kTB; c += 2) {
              s0 += xc[(size_t)c * (kTB + 1)] * tbuf[J0 + c];
              s1 += xc[(size_t)(c + 1) * (kTB + 1)] * tbuf[J0 + c + 1];
            }
            val = s0 + s1;
            ybuf[t] = val;
          }
          __syncthreads();
          if (t < J0) {
            double s0 = 0.0, s1 = 0.0;
            for (int c = 0; c + 1 < bs; c += 2) {
              s0 += M[(size_t)(J0 + c) * LD + t] * ybuf[J0 + c];
              s1 += M[(size_t)(J0 + c + 1) * LD + t] * ybuf[J0 + c + 1];
            }
            if (bs & 1) s0 += M[(size_t)(J0 + bs - 1) * LD + t] * ybuf[J0 + bs - 1];
            acc -= s0 + s1;
          }
        }
        // val = x_t for t < N
        if (t < n) xs[t] = val;
        __syncthreads();
        FCCQP_PROF(11);
        // K4 + K5
        double rx = 0.0, rc = 0.0;
        if (t < n) {
          const double xv = val;
          const double xb = clampd(xv + mux[t], vlb[t], vub[t]);
          xbar[t] = xb;
          const double r = xv - xb;
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
        // NaN-propagating max: fmax drops NaNs, so flag them separately
        if (rx != rx || rc != rc) status_flag = 2;
        block_reduce2<false>(rx, rc, red, parity);
        res_x = rx; res_c = rc;
        FCCQP_PROF(12);
        if (p.prof && tid == 0) s_prof[15] += 1;
        if (rc < p.eps_fcone && rx < p.eps_bound) { n_iter = iter; break; }
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
    fact_cycles = 0;
    FCCQP_PROF(13);
    if (p.prof && tid == 0) s_prof[14] += 1;
  }
  if (p.prof && tid == 0)
    for (int i = 0; i < 16; ++i) atomicAdd(p.prof + i, s_prof[i]);
#undef FCCQP_PROF
}

}  // namespace fccqp
