// fccqp_kernel.cuh -- sm_100a device code of the batched FCCQP solve (v3: FP64 tensor-core tiles).
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
// This is NOT a port of the Eigen code paths.  The linear algebra is re-derived so that it is
// pivot-free, symmetric, and made of 8x8 tiles that map onto the FP64 tensor-core instruction
// (DMMA, mma.sync.m8n8k4.f64):
//
//   * K1.  The reference solves the indefinite system [[Q,A'],[A,0]] s = [-b; b_eq] whose (1,1)
//     block is singular (zero-cost force variables) through a failed LDLT and a complete
//     orthogonal decomposition.  Here the SAME solution comes from the augmented-Lagrangian
//     form [[Q + sigma A'A, A'],[A,0]] with right-hand side [-b + sigma A' b_eq; b_eq]: on
//     {Ax = b_eq} the added term is constant, so x is unchanged, while Q + sigma A'A is positive
//     definite exactly when the KKT matrix is nonsingular.  That matrix is quasi-definite, so an
//     unpivoted LDL^T exists; one step of iterative refinement against the ORIGINAL system
//     removes the sigma-dependent rounding.
//   * K2 is the same unpivoted blocked LDL^T on [[Q + rho I, A'],[A,0]].
//   * Blocked right-looking LDL^T, panel width 8.  Per panel: (P1) one thread factors the 8x8
//     diagonal tile in registers and inverts its unit-lower factor; (P2) every sub-diagonal tile
//     becomes L = A * inv(L11)' * inv(D) with two DMMAs; (P3) the trailing tiles get
//     C -= (L D) L' with two DMMAs each, operands read straight from the tile storage.
//   * K3.  The 8x8 inverses are composed (again with DMMAs) into explicit inverses of the 32x32
//     diagonal blocks of L, stored in place.  A triangular solve is then ceil(N/32) steps of
//     "one warp applies a 32x32 inverse, the others subtract a 32-column slab" instead of N
//     dependent scalar steps.
//
// Storage: the lower triangle of the padded KKT matrix as dense 8x8 tiles (512 B each),
// tile (I,J) at index I(I+1)/2+J.  Variables are padded to a multiple of 8 (n8) before the
// constraint rows start (pads are decoupled unit pivots).  Inside a tile element (r,c) lives at
// r*8 + (((c/2) ^ (r/2)) & 3)*2 + (c&1): the 16-byte chunks of a row are XOR-swizzled by r/2
// (the TMA SWIZZLE_64B pattern), which makes the tensor-core fragment access (row = lane/4,
// column pair = lane%4), the row-per-thread access of the forward solve and the
// column-per-thread access of the backward solve all bank-conflict-free within a tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fccqp {

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

// Shared-memory carve-up, identical on host (sizing) and device (pointers).
struct Layout {
  int n, m, nc;
  int n8, m8, N8;   // padded sizes (multiples of 8)
  int NB, NBx;      // number of 8-tiles per side; tiles covering the variable rows
  int NB32, NT;     // 32-row solve blocks; padded vector length (NB32 * 32)
  size_t off_M, off_dinv, off_dneg, off_tbuf, off_ybuf, off_sbuf, off_xs, off_lcbar, off_muc, off_mu;
  size_t off_red, off_int;
  size_t doubles_total;
  __host__ __device__ static inline size_t up2(size_t v) { return (v + 1) & ~size_t(1); }
  __host__ __device__ Layout(int n_, int m_, int nc_) {
    n = n_; m = m_; nc = nc_;
    n8 = (n + 7) & ~7; m8 = (m + 7) & ~7; N8 = n8 + m8;
    NB = N8 >> 3; NBx = n8 >> 3;
    NB32 = (N8 + 31) >> 5; NT = NB32 * 32;
    size_t o = 0;
    off_M = o;     o += (size_t)(NB * (NB + 1) / 2) * 64;
    off_dinv = o;  o += NT;
    off_dneg = o;  o += NT;
    off_tbuf = o;  o += NT;
    off_ybuf = o;  o += NT;
    off_sbuf = o;  o += NT;
    off_xs = o;    o += n8 + 8;
    off_lcbar = o; o += up2(nc + 2);
    off_muc = o;   o += up2(nc + 2);
    off_mu = o;    o += up2(nc / 3 + 2);
    off_red = o;   o += 4 * 32;   // block_reduce2 scratch: 2 buffers x 2 values x 32 warps
    off_int = o;   o += 32;       // ints: work index, profiling slots
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

// ---------------------------------------------------------------------------
// Tile storage
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ int tile_off(int I, int J) { return (((I * (I + 1)) >> 1) + J) << 6; }
// element (r,c) inside a tile (16-byte chunks XOR-swizzled by r/2)
__host__ __device__ __forceinline__ int el_off(int r, int c) {
  return (r << 3) + ((((c >> 1) ^ (r >> 1)) & 3) << 1) + (c & 1);
}
// element (i,j), i >= j in tile terms, of the packed matrix
__host__ __device__ __forceinline__ int mat_off(int i, int j) { return tile_off(i >> 3, j >> 3) + el_off(i & 7, j & 7); }

// FP64 tensor-core step: C(8x8) += A(8x4) B(4x8).  Fragments (PTX ISA, m8n8k4 .f64): lane l holds
// A[l/4][l%4], B[l%4][l/4], C[l/4][2(l%4)], C[l/4][2(l%4)+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// C(8x8) += Aop(8x8) Bop(8x8) with the contraction index split as k = 2(l%4)+s over two DMMAs:
//   a = (Aop[l/4][2q], Aop[l/4][2q+1]),  b = (Bop[2q][l/4], Bop[2q+1][l/4]),  q = l%4.
// With "fragC(X)" = (X[l/4][2q], X[l/4][2q+1]) (one 16-byte load) and
//      "fragT(X)" = (X[2q][l/4], X[2q+1][l/4]) (two 8-byte loads):
//   X Y'  : a = fragC(X), b = fragC(Y)        X Y  : a = fragC(X), b = fragT(Y)
//   X' Y  : a = fragT(X), b = fragT(Y)        X' Y': a = fragT(X), b = fragC(Y)
// and a product held in registers as fragC(P) is fragT(P') for the next product.
__device__ __forceinline__ void mma8(double2& c, const double2 a, const double2 b) {
  dmma(c.x, c.y, a.x, b.x);
  dmma(c.x, c.y, a.y, b.y);
}

// LDL^T of one 8x8 tile by ONE thread, all in registers (shortest dependent chain), followed by
// the in-place inverse of its unit-lower factor.  Writes inv(L11) (zeros above, ones on the
// diagonal) back into the tile, 1/d and -d into dinv / dneg.
__device__ __forceinline__ void factor_diag_tile(double* __restrict__ tile, double* __restrict__ dinv,
                                                 double* __restrict__ dneg) {
  double a[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) a[r][c] = tile[el_off(r, c)];
  double rd[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    rd[c] = fast_rcp(a[c][c]);
#pragma unroll
    for (int r = c + 1; r < 8; ++r) {
      // rows c2 < r of this column already hold l_{c2,c}; a[r][c] is still unscaled (= l_rc d_c)
      const double arc = a[r][c];
      const double l = arc * rd[c];
#pragma unroll
      for (int c2 = c + 1; c2 < r; ++c2) a[r][c2] -= arc * a[c2][c];
      a[r][r] -= arc * l;
      a[r][c] = l;
    }
  }
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    *reinterpret_cast<double2*>(dinv + c) = make_double2(rd[c], rd[c + 1]);
    *reinterpret_cast<double2*>(dneg + c) = make_double2(-a[c][c], -a[c + 1][c + 1]);
  }
  // X = inv(L), row by row: X[i][j] = -(L[i][j] + sum_{j<k<i} L[i][k] X[k][j])
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    double x[8];
#pragma unroll
    for (int j = 0; j < i; ++j) {
      double s = a[i][j];
#pragma unroll
      for (int k = j + 1; k < i; ++k) s += a[i][k] * a[k][j];
      x[j] = -s;
    }
#pragma unroll
    for (int j = 0; j < i; ++j) a[i][j] = x[j];
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c0 = 2 * cc, c1 = 2 * cc + 1;
      const double v0 = c0 < r ? a[r][c0] : (c0 == r ? 1.0 : 0.0);
      const double v1 = c1 < r ? a[r][c1] : (c1 == r ? 1.0 : 0.0);
      *reinterpret_cast<double2*>(tile + el_off(r, c0)) = make_double2(v0, v1);
    }
  }
}

// ---------------------------------------------------------------------------
// The fused solve kernel.  kThreads >= padded KKT size N8 (one thread per KKT row in the
// triangular solves and all vector work).
// ---------------------------------------------------------------------------
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) fccqp_solve_kernel(const SolveParams p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  const int n = p.n, m = p.m, nc = p.nc, lcs = p.lcs;
  const Layout L(n, m, nc);
  const int n8 = L.n8, N8 = L.N8, NB = L.NB, NBx = L.NBx, NB32 = L.NB32;
  const int NBT = NB * (NB + 1) / 2;

  double* M = smem + L.off_M;
  double* dinv = smem + L.off_dinv;
  double* dneg = smem + L.off_dneg;
  double* tbuf = smem + L.off_tbuf;
  double* ybuf = smem + L.off_ybuf;
  double* sbuf = smem + L.off_sbuf;
  double* xs = smem + L.off_xs;
  double* lcbar = smem + L.off_lcbar;
  double* muc = smem + L.off_muc;
  double* vmu = smem + L.off_mu;
  double* red = smem + L.off_red;
  int* ibuf = reinterpret_cast<int*>(smem + L.off_int);
  int* s_work = ibuf;  // [1]
  unsigned long long* s_prof = reinterpret_cast<unsigned long long*>(ibuf + 2);  // [16]
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
  // thread-per-row identity: rows [0,n) variables, [n,n8) pads, [n8,n8+m) constraints, rest pads
  const int t = tid;
  const bool is_x = t < n;
  const bool is_c = t >= n8 && t < n8 + m;
  const bool is_row = t < N8;
  const bool in_cone = is_x && t >= lcs && t < lcs + nc;
  const int tb = t >> 3, tr = t & 7, tf = tr >> 1;  // tile row, row in tile, row swizzle
  // tensor-core fragment coordinates of this lane
  const int fr = lane >> 2, fq = lane & 3;
  const int fragC = (fr << 3) + (((fq ^ (fr >> 1)) & 3) << 1);                       // (fr, 2fq..2fq+1)
  const int fragT = ((2 * fq) << 3) + ((((fr >> 1) ^ fq) & 3) << 1) + (fr & 1);      // (2fq, fr); +8 for (2fq+1, fr)

  for (;;) {
    __syncthreads();  // previous QP fully retired (smem reuse) before taking new work
    if (tid == 0) *s_work = (int)atomicAdd(p.work_counter, 1u);
    __syncthreads();
    const int qp = *s_work;
    if (qp >= p.B) break;

    const double* Qg = p.Q + (size_t)qp * p.q_bs;
    const double* Ag = p.A + (size_t)qp * p.a_bs;
    // Q is symmetric: walk it along whichever stride is contiguous.
    const long long q_slow = p.q_cs <= p.q_rs ? p.q_rs : p.q_cs;
    const long long q_fast = p.q_cs <= p.q_rs ? p.q_cs : p.q_rs;
    const bool q_vec = q_fast == 1 && ((reinterpret_cast<uintptr_t>(Qg) | (uintptr_t)(q_slow * 8)) & 15) == 0;
    const bool a_vec = p.a_cs == 1 && ((reinterpret_cast<uintptr_t>(Ag) | (uintptr_t)(p.a_rs * 8)) & 15) == 0;

    // ---------------- K0: vectors (one register per row and vector) ----------------
    double v_b = 0.0;       // b (variable rows) or b_eq (constraint rows)
    double v_lb = 0.0, v_ub = 0.0, v_mux = 0.0, v_xbar = 0.0, v_x = 0.0;
    int finite_bounds = 0;
    if (is_x) {
      v_b = p.b[(size_t)qp * p.b_bs + t];
      v_lb = p.lb[(size_t)qp * p.lb_bs + t];
      v_ub = p.ub[(size_t)qp * p.ub_bs + t];
      if (!isinf(v_lb) || !isinf(v_ub)) finite_bounds = 1;
      if (p.warm) {
        v_x = p.x[(size_t)qp * n + t];
        v_mux = p.mu_x[(size_t)qp * n + t];
      }
    } else if (is_c) {
      v_b = p.beq[(size_t)qp * p.beq_bs + (t - n8)];
    }
    if (t < n8) xs[t] = v_x;
    if (t < nc / 3) vmu[t] = p.mu[(size_t)qp * p.mu_bs + t];
    if (t < nc) muc[t] = p.warm ? p.mu_c[(size_t)qp * nc + t] : 0.0;
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

      // ---------------- assemble the lower tiles of the padded KKT matrix ----------------
      // one warp per tile, lane = (row fr, column pair 2fq): 8 x 64-byte row segments from HBM/L2
      {
        int I = 0, J = warp;
        while (J > I) { J -= I + 1; ++I; }
        for (int idx = warp; idx < NBT; idx += kWarps) {
          double* dst = M + tile_off(I, J) + fragC;
          const int gi = 8 * I + fr, gc = 8 * J + 2 * fq;
          if (J >= NBx) {
            // (2,2) block: zero; decoupled unit pivots on the constraint pads
            const int k = gi - n8;
            *reinterpret_cast<double2*>(dst) =
                make_double2((gi == gc && k >= m) ? 1.0 : 0.0, (gi == gc + 1 && k >= m) ? 1.0 : 0.0);
          } else if (I < NBx) {
            // Q (symmetric; row gi read along the contiguous direction), unit pivots on the pads
            const bool e0 = gi < n && gc < n, e1 = gi < n && gc + 1 < n;
            const double* src = Qg + gi * q_slow + gc * q_fast;
            if (e1 && q_vec) {
              cp_async16(dst, src);
            } else {
              if (e0) cp_async8(dst, src); else dst[0] = (gi == gc) ? 1.0 : 0.0;
              if (e1) cp_async8(dst + 1, src + q_fast); else dst[1] = (gi == gc + 1) ? 1.0 : 0.0;
            }
          } else {
            // A rows
            const int k = gi - n8;
            const bool e0 = k < m && gc < n, e1 = k < m && gc + 1 < n;
            const double* src = Ag + k * p.a_rs + gc * p.a_cs;
            if (e1 && a_vec) {
              cp_async16(dst, src);
            } else {
              if (e0) cp_async8(dst, src); else dst[0] = 0.0;
              if (e1) cp_async8(dst + 1, src + p.a_cs); else dst[1] = 0.0;
            }
          }
          J += kWarps;
          while (J > I) { J -= I + 1; ++I; }
        }
      }
      cp_async_wait_all();
      __syncthreads();
      FCCQP_PROF(1);

      double rhs0 = 0.0;  // pass-0 right-hand side of row t
      if (pass == 1) {
        if (is_x) M[mat_off(t, t)] += p.rho;
        __syncthreads();
      } else {
        // sigma = trace(Q) / ||A||_F^2 balances the two terms of Q + sigma A'A
        double trq = is_x ? M[mat_off(t, t)] : 0.0, fro = 0.0;
        if (is_c) {
          tbuf[t] = v_b;
          for (int jb = 0; jb < NBx; ++jb) {
            const double2* row = reinterpret_cast<const double2*>(M + tile_off(tb, jb) + tr * 8);
#pragma unroll
            for (int c = 0; c < 4; ++c) { const double2 v = row[c ^ tf]; fro += v.x * v.x + v.y * v.y; }
          }
        }
        block_reduce2<true>(trq, fro, red, parity);
        const double sigma = (trq > 0.0 && fro > 0.0 && isfinite(trq / fro)) ? trq / fro : 1.0;
        // rhs_x = -b + sigma A' b_eq (needs A before the factorization overwrites it)
        if (is_x) {
          double s0 = 0.0, s1 = 0.0;
          const int colo = tr & 1, colc = tr >> 1;
          for (int ib = NBx; ib < NB; ++ib) {
            const double* tl = M + tile_off(ib, tb) + colo;
            const double* bq = tbuf + ib * 8;
#pragma unroll
            for (int r = 0; r < 8; r += 2) {
              s0 += tl[r * 8 + (((colc ^ (r >> 1)) & 3) << 1)] * bq[r];
              s1 += tl[(r + 1) * 8 + (((colc ^ (r >> 1)) & 3) << 1)] * bq[r + 1];
            }
          }
          rhs0 = -v_b + sigma * (s0 + s1);
        } else if (is_c) {
          rhs0 = v_b;
        }
        // H += sigma A'A, tile by tile on the tensor cores (lower tiles of the variable block)
        if (NB > NBx) {
          const int TT = NBx * (NBx + 1) / 2;
          const int lo = TT * warp / kWarps, hi = TT * (warp + 1) / kWarps;
          int I = 0, J = lo;
          while (J > I) { J -= I + 1; ++I; }
          for (int idx = lo; idx < hi; ++idx) {
            double2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
            int kb = NBx;
            for (; kb + 1 < NB; kb += 2) {
              const double* ai = M + tile_off(kb, I) + fragT;
              const double* aj = M + tile_off(kb, J) + fragT;
              const double* ai2 = M + tile_off(kb + 1, I) + fragT;
              const double* aj2 = M + tile_off(kb + 1, J) + fragT;
              const double2 a0 = make_double2(ai[0], ai[8]), b0 = make_double2(aj[0], aj[8]);
              const double2 a1 = make_double2(ai2[0], ai2[8]), b1 = make_double2(aj2[0], aj2[8]);
              mma8(acc0, a0, b0);
              mma8(acc1, a1, b1);
            }
            if (kb < NB) {
              const double* ai = M + tile_off(kb, I) + fragT;
              const double* aj = M + tile_off(kb, J) + fragT;
              mma8(acc0, make_double2(ai[0], ai[8]), make_double2(aj[0], aj[8]));
            }
            double2* cp = reinterpret_cast<double2*>(M + tile_off(I, J) + fragC);
            double2 c = *cp;
            c.x += sigma * (acc0.x + acc1.x);
            c.y += sigma * (acc0.y + acc1.y);
            *cp = c;
            if (++J > I) { J = 0; ++I; }
          }
        }
        __syncthreads();
      }
      FCCQP_PROF(2);

      // ---------------- unpivoted blocked LDL^T on 8x8 tiles ----------------
      for (int k = 0; k < NB; ++k) {
        const int k0 = k * 8;
        double* dtile = M + tile_off(k, k);
        // --- P1: diagonal tile, one thread (rotating over the warps), registers only
        if (tid == ((k % kWarps) << 5)) factor_diag_tile(dtile, dinv + k0, dneg + k0);
        __syncthreads();
        FCCQP_PROF(3);
        if (k + 1 < NB) {
          // --- P2: L_ik = A_ik inv(L11)' inv(D11) for the tiles below the diagonal one
          {
            const double2 li = *reinterpret_cast<const double2*>(dtile + fragC);
            const double2 di = *reinterpret_cast<const double2*>(dinv + k0 + 2 * fq);
            for (int i = k + 1 + warp; i < NB; i += kWarps) {
              double2* ap = reinterpret_cast<double2*>(M + tile_off(i, k) + fragC);
              double2 w = make_double2(0.0, 0.0);
              mma8(w, *ap, li);
              *ap = make_double2(w.x * di.x, w.y * di.y);
            }
          }
          __syncthreads();
          FCCQP_PROF(4);
          // --- P3: trailing tiles C_ij -= (L_ik D) L_jk', k < j <= i, split evenly over the warps
          {
            const int R = NB - 1 - k, T = R * (R + 1) / 2;
            const int lo = T * warp / kWarps, hi = T * (warp + 1) / kWarps;
            const double2 dn = *reinterpret_cast<const double2*>(dneg + k0 + 2 * fq);
            int ri = 0, rj = lo;
            while (rj > ri) { rj -= ri + 1; ++ri; }
            int idx = lo;
            while (idx < hi) {
              const int cnt = min(ri - rj + 1, hi - idx);
              const int i = k + 1 + ri;
              double2 a = *reinterpret_cast<const double2*>(M + tile_off(i, k) + fragC);
              a.x *= dn.x; a.y *= dn.y;   // -(L_ik D)
              double* cp = M + tile_off(i, k + 1 + rj) + fragC;   // consecutive j: +64 doubles
              int j = k + 1 + rj;
              int c = 0;
              for (; c + 1 < cnt; c += 2, j += 2, cp += 128) {
                const double2 b0 = *reinterpret_cast<const double2*>(M + tile_off(j, k) + fragC);
                const double2 b1 = *reinterpret_cast<const double2*>(M + tile_off(j + 1, k) + fragC);
                double2 c0 = *reinterpret_cast<const double2*>(cp);
                double2 c1 = *reinterpret_cast<const double2*>(cp + 64);
                mma8(c0, a, b0);
                mma8(c1, a, b1);
                *reinterpret_cast<double2*>(cp) = c0;
                *reinterpret_cast<double2*>(cp + 64) = c1;
              }
              if (c < cnt) {
                const double2 b0 = *reinterpret_cast<const double2*>(M + tile_off(j, k) + fragC);
                double2 c0 = *reinterpret_cast<const double2*>(cp);
                mma8(c0, a, b0);
                *reinterpret_cast<double2*>(cp) = c0;
              }
              idx += cnt;
              rj += cnt;
              if (rj > ri) { rj = 0; ++ri; }
            }
          }
          __syncthreads();
          FCCQP_PROF(5);
        }
      }
      // ---------------- explicit inverses of the 32x32 diagonal blocks of L, in place ----------------
      // level 1: 16x16 = [[X1,0],[-X2 L21 X1, X2]] from the 8x8 inverses left by P1
      for (int a = warp; 2 * a + 1 < NB; a += kWarps) {
        const double* x1 = M + tile_off(2 * a, 2 * a) + fragT;
        double2* l21 = reinterpret_cast<double2*>(M + tile_off(2 * a + 1, 2 * a) + fragC);
        const double2 x2 = *reinterpret_cast<const double2*>(M + tile_off(2 * a + 1, 2 * a + 1) + fragC);
        double2 tt = make_double2(0.0, 0.0);                 // fragC((L21 X1)') = fragT(L21 X1)
        mma8(tt, make_double2(x1[0], x1[8]), *l21);          // X1' L21'
        double2 r = make_double2(0.0, 0.0);
        mma8(r, x2, tt);                                     // X2 (L21 X1)
        *l21 = make_double2(-r.x, -r.y);
      }
      __syncthreads();
      // level 2: 32x32 = [[A,0],[-B L A, B]] with 16x16 A, B; one warp per (block, tile column).
      // Both columns of a block read tiles the other one overwrites: all products first, one
      // barrier, then the stores (the two columns of a block always share a round).
      for (int base = 0; base < 2 * NB32; base += kWarps) {
        const int w = base + warp;
        const int q4 = (w >> 1) * 4, col = w & 1;
        const bool act = w < 2 * NB32 && q4 + 2 < NB;
        const bool two = q4 + 3 < NB;  // second tile row of the lower-left 16x16 exists
        double2 r0 = make_double2(0.0, 0.0), r1 = make_double2(0.0, 0.0);
        if (act) {
          // T(:,col) = L A(:,col), kept transposed in registers
          double2 t0 = make_double2(0.0, 0.0), t1 = make_double2(0.0, 0.0);
          const double2 l01 = *reinterpret_cast<const double2*>(M + tile_off(q4 + 2, q4 + 1) + fragC);
          double2 l11 = make_double2(0.0, 0.0);
          if (two) l11 = *reinterpret_cast<const double2*>(M + tile_off(q4 + 3, q4 + 1) + fragC);
          if (col == 0) {
            const double2 l00 = *reinterpret_cast<const double2*>(M + tile_off(q4 + 2, q4) + fragC);
            const double* a00 = M + tile_off(q4, q4) + fragT;
            const double* a10 = M + tile_off(q4 + 1, q4) + fragT;
            const double2 f00 = make_double2(a00[0], a00[8]), f10 = make_double2(a10[0], a10[8]);
            mma8(t0, f00, l00); mma8(t0, f10, l01);          // T00' = A00' L00' + A10' L01'
            if (two) {
              const double2 l10 = *reinterpret_cast<const double2*>(M + tile_off(q4 + 3, q4) + fragC);
              mma8(t1, f00, l10); mma8(t1, f10, l11);        // T10'
            }
          } else {
            const double* a11 = M + tile_off(q4 + 1, q4 + 1) + fragT;
            const double2 f11 = make_double2(a11[0], a11[8]);
            mma8(t0, f11, l01);                              // T01' = A11' L01'
            if (two) mma8(t1, f11, l11);                     // T11'
          }
          // R(:,col) = B T(:,col)
          const double2 b00 = *reinterpret_cast<const double2*>(M + tile_off(q4 + 2, q4 + 2) + fragC);
          mma8(r0, b00, t0);
          if (two) {
            const double2 b10 = *reinterpret_cast<const double2*>(M + tile_off(q4 + 3, q4 + 2) + fragC);
            const double2 b11 = *reinterpret_cast<const double2*>(M + tile_off(q4 + 3, q4 + 3) + fragC);
            mma8(r1, b10, t0); mma8(r1, b11, t1);
          }
        }
        __syncthreads();
        if (act) {
          *reinterpret_cast<double2*>(M + tile_off(q4 + 2, q4 + col) + fragC) = make_double2(-r0.x, -r0.y);
          if (two)
            *reinterpret_cast<double2*>(M + tile_off(q4 + 3, q4 + col) + fragC) = make_double2(-r1.x, -r1.y);
        }
      }
      __syncthreads();
      fact_cycles += (unsigned long long)(clock64() - t_f0);
      FCCQP_PROF(6);

      if (pass == 1) {
        // ADMM initial slack (fcc_qp.cpp:74-75): x_bar = x, lambda_c_bar = x[lambda_c segment]
        v_xbar = v_x;
        if (t < nc) lcbar[t] = xs[lcs + t];
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
          acc = acc0;
        } else if (is_x) {
          // -(b + q_rho), q_rho = -rho (xbar - mu_x) with the cone segment overwritten (fcc_qp.cpp:81-83)
          const double w = in_cone ? (lcbar[t - lcs] - muc[t - lcs]) : (v_xbar - v_mux);
          const double q_rho = -p.rho * w;
          acc = -(v_b + q_rho);
        } else if (is_c) {
          acc = v_b;
        }
        // ---- forward: L y = rhs, 32 rows per step (warp J applies inv(L_JJ), later warps subtract)
        double val = 0.0;
        for (int J = 0; J < NB32; ++J) {
          const int Jb0 = J * 4;
          if (warp == J) {
            tbuf[t] = acc;
            __syncwarp();
            double s0 = 0.0, s1 = 0.0;
            if (is_row) {
              for (int jb = Jb0; jb <= tb; ++jb) {
                const double2* xr = reinterpret_cast<const double2*>(M + tile_off(tb, jb) + tr * 8);
                const double2* tv = reinterpret_cast<const double2*>(tbuf + jb * 8);
#pragma unroll
                for (int c = 0; c < 4; ++c) {   // logical chunk c lives at physical chunk c ^ tf
                  const double2 x = xr[c ^ tf], v = tv[c];
                  s0 += x.x * v.x; s1 += x.y * v.y;
                }
              }
            }
            val = s0 + s1;
            ybuf[t] = val;
          }
          __syncthreads();
          if (warp > J && is_row) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const double2* lr = reinterpret_cast<const double2*>(M + tile_off(tb, Jb0 + jj) + tr * 8);
              const double2* yv = reinterpret_cast<const double2*>(ybuf + (Jb0 + jj) * 8);
#pragma unroll
              for (int c = 0; c < 4; c += 2) {
                const double2 l0 = lr[c ^ tf], y0 = yv[c];
                const double2 l1 = lr[(c + 1) ^ tf], y1 = yv[c + 1];
                s0 += l0.x * y0.x; s1 += l0.y * y0.y;
                s2 += l1.x * y1.x; s3 += l1.y * y1.y;
              }
            }
            acc -= (s0 + s1) + (s2 + s3);
          }
        }
        // ---- D^{-1}
        acc = is_row ? val * dinv[t] : 0.0;
        // ---- backward: L' x = y
        for (int J = NB32 - 1; J >= 0; --J) {
          const int Jb0 = J * 4, Jb1 = min(Jb0 + 4, NB);
          const int colo = tr & 1, colc = tr >> 1;
          if (warp == J) {
            tbuf[t] = acc;
            __syncwarp();
            double s0 = 0.0, s1 = 0.0;
            if (is_row) {
              for (int ib = tb; ib < Jb1; ++ib) {
                const double* xc = M + tile_off(ib, tb) + colo;
                const double* tv = tbuf + ib * 8;
#pragma unroll
                for (int r = 0; r < 8; r += 2) {
                  s0 += xc[r * 8 + (((colc ^ (r >> 1)) & 3) << 1)] * tv[r];
                  s1 += xc[(r + 1) * 8 + (((colc ^ (r >> 1)) & 3) << 1)] * tv[r + 1];
                }
              }
            }
            val = s0 + s1;
            ybuf[t] = val;
          }
          __syncthreads();
          if (warp < J) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int ib = Jb0; ib < Jb1; ++ib) {
              const double* lc = M + tile_off(ib, tb) + colo;
              const double* yv = ybuf + ib * 8;
#pragma unroll
              for (int r = 0; r < 8; r += 4) {
                s0 += lc[r * 8 + (((colc ^ (r >> 1)) & 3) << 1)] * yv[r];
                s1 += lc[(r + 1) * 8 + (((colc ^ (r >> 1)) & 3) << 1)] * yv[r + 1];
                s2 += lc[(r + 2) * 8 + (((colc ^ ((r + 2) >> 1)) & 3) << 1)] * yv[r + 2];
                s3 += lc[(r + 3) * 8 + (((colc ^ ((r + 2) >> 1)) & 3) << 1)] * yv[r + 3];
              }
            }
            acc -= (s0 + s1) + (s2 + s3);
          }
        }
        FCCQP_PROF(7);
        // val = solution component of row t (t < N8)

        if (pass == 0) {
          sol += val;
          if (iter + 1 < iters) {
            // ---- iterative refinement against the ORIGINAL system: r = [-b; b_eq] - [[Q,A'],[A,0]] s
            if (is_row) sbuf[t] = sol;
            __syncthreads();
            double r = 0.0;
            if (is_x) {
              double s0 = 0.0, s1 = 0.0;
#pragma unroll 10
              for (int j = 0; j < n; ++j) s0 += Qg[j * q_slow + t * q_fast] * sbuf[j];           // Q symmetric
#pragma unroll 10
              for (int k = 0; k < m; ++k) s1 += Ag[k * p.a_rs + t * p.a_cs] * sbuf[n8 + k];      // A' y
              r = -v_b - s0 - s1;
            }
            // rows of A: warp per row pair, lanes over columns
            for (int k = 2 * warp; k < m; k += 2 * kWarps) {
              double s0 = 0.0, s1 = 0.0;
              const bool two = k + 1 < m;
              for (int j = lane; j < n; j += 32) {
                const double xj = sbuf[j];
                s0 += Ag[k * p.a_rs + j * p.a_cs] * xj;
                if (two) s1 += Ag[(k + 1) * p.a_rs + j * p.a_cs] * xj;
              }
              s0 = warp_sum(s0); s1 = warp_sum(s1);
              if (lane == 0) { ybuf[n8 + k] = s0; if (two) ybuf[n8 + k + 1] = s1; }
            }
            __syncthreads();
            if (is_c) r = v_b - ybuf[t];
            acc0 = r;
          } else {
            v_x = sol;
            if (t < n8) xs[t] = is_x ? sol : 0.0;
            if (p.dbg_x0 && is_x) p.dbg_x0[(size_t)qp * n + t] = sol;
            __syncthreads();
          }
          FCCQP_PROF(8);
          continue;
        }

        // ---- K4 + K5 (pass 1)
        if (is_x) { xs[t] = val; v_x = val; }
        __syncthreads();
        double rx = 0.0, rc = 0.0;
        if (is_x) {
          const double xb = clampd(val + v_mux, v_lb, v_ub);
          v_xbar = xb;
          const double r = val - xb;
          v_mux += r;
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
    if (is_x) {
      const double d = v_x - clampd(v_x, v_lb, v_ub);
      bv = d * d;
      if (!isfinite(v_x)) bad = 1;
    }
    if (tid < nc / 3) {
      const int o = lcs + 3 * tid;
      const double r = sqrt(xs[o] * xs[o] + xs[o + 1] * xs[o + 1]) - vmu[tid] * xs[o + 2];
      fv = r > 0.0 ? r : 0.0;
    }
    block_reduce2<true>(bv, fv, red, parity);
    bad = __syncthreads_or(bad | (status_flag == 2));
    if (is_x) {
      p.x[(size_t)qp * n + t] = v_x;
      if (p.mu_x) p.mu_x[(size_t)qp * n + t] = v_mux;
    }
    if (p.mu_c && t < nc) p.mu_c[(size_t)qp * nc + t] = muc[t];
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
