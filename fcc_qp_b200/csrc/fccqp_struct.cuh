// fccqp_struct.cuh -- structure-exploiting variant of the fused FCCQP solve.
//
// Same contract as fccqp_solve_kernel (fccqp_kernel.cuh): one CTA solves one QP at a time,
// FCCQP::Solve + GetSolution of the reference (src/fcc_qp.cpp:114-207), FP64 throughout.  What
// changes is the KKT system that gets factored.  Whole-body-control QPs are mostly made of
// variables that appear in the cost only through their own square -- torques, constraint forces,
// slacks: the rows of Q that belong to them are diagonal (fccqp.pdf eq. 10; 38-40 of the 60
// variables of the walking log) -- and the reference assembles and factors the full
// (n + m) x (n + m) matrix regardless (src/fcc_qp.cpp:141-150, 62-71).  Here every QP is classified ON
// THE DEVICE while it is staged (nothing is assumed about the layout of x, nothing is declared by
// the caller):
//
//   R   variables whose Q row has an off-diagonal entry                      -> stay in the system
//   D+  separable (diagonal-only Q row), cost q_j > 0, A_eq column with >= 2 entries
//                                                                            -> eliminated, column kept
//   D1  separable, q_j > 0, A_eq column with 0 or 1 entries (actuator selection, slack identity,
//       forces of contacts that are off)                                    -> eliminated, O(1) each
//   D0  separable, q_j = 0 (zero-cost constraint forces)                     -> trailing block
//
// With h_j = q_j (+ rho in the ADMM pass) the separable variables satisfy
//   x_j = (r_j - a_j' y) / h_j,
// and the remaining unknowns (x_R, y, x_D0) solve the REDUCED symmetric system
//
//   [ Q_RR (+rho I)   A_R'            0        ] [x_R ]   [ r_R                          ]
//   [ A_R            -C               A_D0     ] [ y  ] = [ r_y - sum_j a_j r_j / h_j    ]
//   [ 0               A_D0'           (rho I)  ] [x_D0]   [ r_D0                         ]
//
//   C = sum_{j in D+ u D1} a_j a_j' / h_j   (D1 columns only touch one diagonal entry).
//
// In this order -- positive block, negative block, zero-cost block last -- the matrix has an
// unpivoted LDL^T whenever Qbar = blkdiag(Q_RR, diag q_D) > 0 and A_eq keeps full row rank without
// the D0 columns; the pivots then come out (+, -, +).  The inertia is checked after every
// factorization and a QP that fails it (or whose structure exceeds the caps the shared-memory
// layout was sized for) is handed to the general kernel through a device-side list, never to the
// CPU.  Cassie log: 72 padded KKT rows instead of 104 (45 tiles instead of 91, 9 tile columns
// instead of 13, no sigma A'A product), humanoid 88 instead of 144, quadruped 56 instead of 88,
// multi-contact humanoid 120 instead of 192.
//
// Accuracy.  Eliminating variables whose cost is 1e-6 puts 1e6-sized terms into C; measured against
// the compiled reference on the walking log (tools/proto/struct_full_check.py, the numpy model of
// this kernel) the unpivoted reduced solve alone is within 2e-7 relative with identical iteration
// counts on all 2019 QPs, and ONE step of iterative refinement of the cold pre-solve against the
// ORIGINAL Q and A_eq (SolveParams::struct_refine, default on) brings it back to 7e-11 -- the level of
// the general kernel.  The ADMM pass needs none (h_j >= rho).
#pragma once
#include "fccqp_kernel.cuh"

namespace fccqp {

enum : int { VT_R = 0, VT_DP = 1, VT_D1 = 2, VT_D0 = 3, VT_NONE = 4 };

// Per-thread view of "its" variable (thread j < n owns variable j in all vector work).
struct VarClass {
  int type, pos, krow, nnz;
  double aval, qd;
};

// Pointers into the int scratch region of StructLayout.
struct StructInts {
  int *vtype, *vpos, *rlist, *dplist, *d0list, *d1var, *rowcnt, *sepf, *wtot;
};

// Classification of one QP (all threads of the CTA).  Reads ALL of Q and A_eq once (the algorithmic
// read of the path: they are inputs, every entry matters); what the assembly reads again afterwards
// comes from L1/L2.  Returns nonzero when the QP does not fit (caps) or is structurally singular
// (a zero-cost variable that no constraint touches).
template <int kThreads>
__device__ __forceinline__ int struct_classify(const int n, const int m, const int m8, const double* __restrict__ Qg,
                                               const long long q_slow, const long long q_fast,
                                               const double* __restrict__ Ag, const long long a_rs, const long long a_cs,
                                               double* __restrict__ qd_s, const StructInts& I, const int capR,
                                               const int capP, const int cap0, VarClass& vc, int& nr, int& ndp, int& nd0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  constexpr int kCol = kThreads / 32;   // n <= kThreads: at most kCol columns per lane
  // ---- Q: one warp per row (read along the contiguous direction of the symmetric matrix), two rows in flight
#pragma unroll 1
  for (int i0 = warp; i0 < n; i0 += 2 * kWarps) {
    const int i1 = i0 + kWarps;
    const bool has1 = i1 < n;
    const double* r0 = Qg + (long long)i0 * q_slow;
    const double* r1 = Qg + (long long)(has1 ? i1 : i0) * q_slow;
    double va[kCol], vb[kCol];
#pragma unroll
    for (int u = 0; u < kCol; ++u) {
      const int c = lane + 32 * u;
      va[u] = c < n ? r0[(long long)c * q_fast] : 0.0;
      vb[u] = c < n ? r1[(long long)c * q_fast] : 0.0;
    }
    bool nza = false, nzb = false;
#pragma unroll
    for (int u = 0; u < kCol; ++u) {
      const int c = lane + 32 * u;
      if (c < n) {
        if (c == i0) qd_s[i0] = va[u]; else nza = nza || (va[u] != 0.0);
        if (has1) { if (c == i1) qd_s[i1] = vb[u]; else nzb = nzb || (vb[u] != 0.0); }
      }
    }
    nza = __any_sync(0xffffffffu, nza);
    nzb = __any_sync(0xffffffffu, nzb);
    if (lane == 0) { I.sepf[i0] = nza ? 0 : 1; if (has1) I.sepf[i1] = nzb ? 0 : 1; }
  }
  // ---- A_eq: one thread per column (coalesced for row-major A): entry count, last entry
  int nnz = 0, kr = 0;
  double av = 0.0;
  if (tid < n) {
    const double* col = Ag + (long long)tid * a_cs;
#pragma unroll 8
    for (int k = 0; k < m; ++k) {
      const double v = col[(long long)k * a_rs];
      if (v != 0.0) { ++nnz; kr = k; av = v; }
    }
  }
  for (int e = tid; e < m8 + 8; e += kThreads) { I.rowcnt[e] = 0; I.d1var[e] = -1; }
  __syncthreads();
  int type = VT_NONE;
  double qd = 0.0;
  if (tid < n) {
    qd = qd_s[tid];
    if (!I.sepf[tid] || !(qd >= 0.0)) type = VT_R;          // (NaN or negative curvature: leave it to the dense block)
    else if (qd < 1e-200) type = VT_D0;
    else type = nnz <= 1 ? VT_D1 : VT_DP;
    if (type == VT_D1 && nnz == 1) atomicAdd(&I.rowcnt[kr], 1);
  }
  __syncthreads();
  // two one-entry columns on the same constraint row: keep them as general columns
  if (type == VT_D1 && nnz == 1 && I.rowcnt[kr] > 1) type = VT_DP;
  const unsigned bR = __ballot_sync(0xffffffffu, type == VT_R);
  const unsigned bP = __ballot_sync(0xffffffffu, type == VT_DP);
  const unsigned b0 = __ballot_sync(0xffffffffu, type == VT_D0);
  if (lane == 0) { I.wtot[3 * warp] = __popc(bR); I.wtot[3 * warp + 1] = __popc(bP); I.wtot[3 * warp + 2] = __popc(b0); }
  __syncthreads();
  int oR = 0, oP = 0, o0 = 0;
  nr = 0; ndp = 0; nd0 = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const int a = I.wtot[3 * w], b = I.wtot[3 * w + 1], c = I.wtot[3 * w + 2];
    if (w < warp) { oR += a; oP += b; o0 += c; }
    nr += a; ndp += b; nd0 += c;
  }
  const unsigned lt = (1u << lane) - 1u;
  const int nr8 = (nr + 7) & ~7;
  int bad = (nr > capR || ndp > capP || nd0 > cap0) ? 1 : 0;   // block-uniform
  int pos = 0;
  if (!bad && tid < n) {
    if (type == VT_R) { pos = oR + __popc(bR & lt); I.rlist[pos] = tid; }
    else if (type == VT_DP) { pos = oP + __popc(bP & lt); I.dplist[pos] = tid; }
    else if (type == VT_D0) { const int q = o0 + __popc(b0 & lt); I.d0list[q] = tid; pos = nr8 + m8 + q; }
    else { pos = kr; if (nnz == 1) I.d1var[kr] = tid; }
    I.vtype[tid] = type;
    I.vpos[tid] = pos;
  }
  bad |= __syncthreads_or(type == VT_D0 && nnz == 0);
  vc.type = type; vc.pos = pos; vc.krow = kr; vc.nnz = nnz; vc.aval = av; vc.qd = qd;
  return bad;
}

// Probe: classifies `ns` QPs spread evenly over the batch and reports the largest structure seen
// (out[0..2] = max nr, ndp, nd0; out[3] = QPs that are structurally singular) -- the host sizes the
// shared-memory layout of the solve kernel from it.
template <int kThreads>
__global__ void __launch_bounds__(kThreads) fccqp_struct_probe_kernel(const SolveParams p, const int ns, int* __restrict__ out) {
  __shared__ double qd_s[kThreads];
  __shared__ int ints[5 * kThreads + 2 * (kThreads + 16) + 96];
  StructInts I;
  int* q = ints;
  I.vtype = q; q += kThreads;
  I.vpos = q; q += kThreads;
  I.rlist = q; q += kThreads;
  I.dplist = q; q += kThreads;
  I.d0list = q; q += kThreads;
  I.d1var = q; q += kThreads + 16;
  I.rowcnt = q; q += kThreads + 16;
  I.wtot = q;
  __shared__ int sepf[kThreads];
  I.sepf = sepf;
  const int s = blockIdx.x;
  const long long qp = ns > 1 ? (long long)s * (p.B - 1) / (ns - 1) : 0;
  const long long q_slow = p.q_cs <= p.q_rs ? p.q_rs : p.q_cs;
  const long long q_fast = p.q_cs <= p.q_rs ? p.q_cs : p.q_rs;
  const int m8 = (p.m + 7) & ~7;
  VarClass vc;
  int nr, ndp, nd0;
  const int bad = struct_classify<kThreads>(p.n, p.m, m8, p.Q + (size_t)qp * p.q_bs, q_slow, q_fast,
                                            p.A + (size_t)qp * p.a_bs, p.a_rs, p.a_cs, qd_s, I, 1 << 30, 1 << 30, 1 << 30,
                                            vc, nr, ndp, nd0);
  if (threadIdx.x == 0) {
    atomicMax(out, nr); atomicMax(out + 1, ndp); atomicMax(out + 2, nd0);
    if (bad) atomicAdd(out + 3, 1);
  }
}

// x-update through the reduced system.  r = this thread's entry of the full-space right-hand side
// (variable rows; the constraint rows' b_eq is in shared memory).  Returns x_j for variable threads.
// use_op: the factors have been replaced by the explicit inverse of the reduced matrix (long-running
// QPs).  refine: one step of iterative refinement against the original Q and A_eq (cold pre-solve).
struct StructQP {
  int nr, nr8, ndp, dpt, nd0, NB, NB32, N8;
};

template <int kThreads>
__device__ __forceinline__ double struct_xsolve(const SolveParams& p, const StructLayout& L, double* __restrict__ smem,
                                                const StructInts& I, const StructQP& S, const VarClass& vc,
                                                const double* __restrict__ Qg, const double* __restrict__ Ag,
                                                const long long q_slow, const long long q_fast, const double r,
                                                const double hi, const double shift, const bool use_op, const bool refine) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr int kWarps = kThreads / 32;
  const int n = p.n, m = p.m;
  double* const M = smem + L.off_M;
  const double* const AP = smem + L.off_AP;
  double* const dinv = smem + L.off_dinv;
  double* const tbuf = smem + L.off_tbuf;
  double* const ybuf = smem + L.off_ybuf;
  double* const sred = smem + L.off_sred;
  double* const rf = smem + L.off_rf;
  double* const vd = smem + L.off_vd;
  double* const d1c = smem + L.off_d1c;
  const double* const beqs = smem + L.off_beq;
  double* const xs = smem + L.off_xs;
  const bool is_x = t < n;
  const int m8 = L.m8, dptc = L.dptc;
  const int yrow = t - S.nr8;                       // constraint index of this thread's reduced row
  const bool row_R = t < S.nr, row_y = yrow >= 0 && yrow < m;
  const int zrow = t - S.nr8 - m8;
  const bool row_0 = zrow >= 0 && zrow < S.nd0;
  // ---- reduced right-hand side
  if (is_x) {
    rf[t] = r;
    if (vc.type == VT_DP) vd[vc.pos] = r * hi;
    else if (vc.type == VT_D1 && vc.nnz == 1) d1c[vc.krow] = vc.aval * (r * hi);
  }
  __syncthreads();
  double acc = 0.0;
  if (row_R) acc = rf[I.rlist[t]];
  else if (row_y) {
    const double* arow = AP + (size_t)((yrow >> 3) * dptc) * 64 + (yrow & 7) * 8;
    double s = 0.0;
#pragma unroll 1
    for (int kt = 0; kt < S.dpt; ++kt) s += row_dot8(arow + 64 * kt, (yrow & 7) >> 1, vd + 8 * kt);
    acc = beqs[yrow] - s;
    if (I.d1var[yrow] >= 0) acc -= d1c[yrow];
  } else if (row_0) acc = rf[I.d0list[zrow]];
  double val;
#ifdef FCCQP_DEV
  unsigned long long* trbuf = nullptr; int trn = 0;
  if (use_op) val = g_apply(M, tbuf, acc, S.NB, S.N8);
  else val = kkt_solve(M, dinv, tbuf, ybuf, acc, S.NB, S.NB32, S.N8, trbuf, trn);
#else
  if (use_op) val = g_apply(M, tbuf, acc, S.NB, S.N8);
  else val = kkt_solve(M, dinv, tbuf, ybuf, acc, S.NB, S.NB32, S.N8);
#endif
  if (t < S.N8) sred[t] = val;
  __syncthreads();
  const double* ys = sred + S.nr8;
  auto recover = [&]() -> double {
    double x = 0.0;
    if (is_x) {
      if (vc.type == VT_R || vc.type == VT_D0) x = sred[vc.pos];
      else if (vc.type == VT_DP) {
        const double* tcol = AP + (size_t)(vc.pos >> 3) * 64 + (vc.pos & 1);
        const int colc = (vc.pos & 7) >> 1;
        double s = 0.0;
#pragma unroll 1
        for (int It = 0; It < L.mt; ++It) s += col_dot8(tcol + (size_t)It * dptc * 64, colc, 0, ys + 8 * It);
        x = (r - s) * hi;
      } else {
        x = vc.nnz == 1 ? (r - vc.aval * ys[vc.krow]) * hi : r * hi;
      }
    }
    return x;
  };
  if (refine) {
    // residual of the ORIGINAL KKT system at (x, y); it vanishes on the eliminated variables by
    // construction, so its R / constraint / D0 entries are the residual of the reduced system
    const double x = recover();
    if (is_x) xs[t] = x;
    __syncthreads();
#pragma unroll 1
    for (int k = warp; k < m; k += kWarps) {
      const double* arow = Ag + (long long)k * p.a_rs;
      double s = 0.0;
      for (int j = lane; j < n; j += 32) s = fma(arow[(long long)j * p.a_cs], xs[j], s);
      s = warp_sum(s);
      if (lane == 0) d1c[k] = beqs[k] - s;
    }
    if (is_x && (vc.type == VT_R || vc.type == VT_D0)) {
      double s0 = (vc.qd + shift) * x, s1 = 0.0;
      if (vc.type == VT_R) {
        s0 = shift * x;
        const double* qcol = Qg + (long long)t * q_fast;
#pragma unroll 4
        for (int a = 0; a < S.nr; ++a) s0 = fma(qcol[(long long)I.rlist[a] * q_slow], sred[a], s0);
      }
      const double* acol = Ag + (long long)t * p.a_cs;
#pragma unroll 4
      for (int k = 0; k < m; ++k) s1 = fma(acol[(long long)k * p.a_rs], ys[k], s1);
      rf[t] = r - (s0 + s1);
    }
    __syncthreads();
    double acc2 = 0.0;
    if (row_R) acc2 = rf[I.rlist[t]];
    else if (row_y) acc2 = d1c[yrow];
    else if (row_0) acc2 = rf[I.d0list[zrow]];
#ifdef FCCQP_DEV
    const double dv = kkt_solve(M, dinv, tbuf, ybuf, acc2, S.NB, S.NB32, S.N8, trbuf, trn);
#else
    const double dv = kkt_solve(M, dinv, tbuf, ybuf, acc2, S.NB, S.NB32, S.N8);
#endif
    if (t < S.N8) sred[t] = val + dv;
    __syncthreads();
  }
  return recover();
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) fccqp_struct_kernel(const SolveParams p) {
  extern __shared__ __align__(16) double smem[];
  const StructLayout& L = p.slay;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  const int n = p.n, m = p.m, nc = p.nc, lcs = p.lcs;
  const int n8 = L.n8, m8 = L.m8, mt = L.mt, dptc = L.dptc;

  double* const M = smem + L.off_M;
  double* const AP = smem + L.off_AP;
  double* const dinv = smem + L.off_dinv;
  double* const dneg = smem + L.off_dneg;
  double* const vd = smem + L.off_vd;
  double* const hinv = smem + L.off_hinv;
  double* const beqs = smem + L.off_beq;
  double* const qd_s = smem + L.off_qd;
  double* const xs = smem + L.off_xs;
  double* const lcbar = smem + L.off_lcbar;
  double* const muc = smem + L.off_muc;
  double* const vmu = smem + L.off_mu;
  double* const red = smem + L.off_red;
  int* const ibuf = reinterpret_cast<int*>(smem + L.off_int);
  int* const s_work = ibuf;
  StructInts I;
  I.vtype = ibuf + L.io_vtype; I.vpos = ibuf + L.io_vpos; I.rlist = ibuf + L.io_rlist; I.dplist = ibuf + L.io_dplist;
  I.d0list = ibuf + L.io_d0list; I.d1var = ibuf + L.io_d1var; I.rowcnt = ibuf + L.io_rowcnt; I.sepf = ibuf + L.io_sepf;
  I.wtot = ibuf + L.io_wtot;

  int parity = 0;
  const int t = tid;
  const bool is_x = t < n;
  const bool in_cone = is_x && t >= lcs && t < lcs + nc;
  const int fr = lane >> 2, fq = lane & 3;
  const int fragC = (fr << 3) + (((fq ^ (fr >> 1)) & 3) << 1);
  const long long q_slow = p.q_cs <= p.q_rs ? p.q_rs : p.q_cs;
  const long long q_fast = p.q_cs <= p.q_rs ? p.q_cs : p.q_rs;

  for (;;) {
    __syncthreads();  // previous QP fully retired (smem reuse) before taking new work
    if (tid == 0) *s_work = (int)atomicAdd(p.work_counter, 1u);
    __syncthreads();
    const int qp = *s_work;
    if (qp >= p.B) break;
    const double* Qg = p.Q + (size_t)qp * p.q_bs;
    const double* Ag = p.A + (size_t)qp * p.a_bs;

    // ---------------- K0: vectors ----------------
    double v_b = 0.0, v_lb = 0.0, v_ub = 0.0, v_mux = 0.0, v_xbar = 0.0, v_x = 0.0;
    int finite_bounds = 0;
    if (is_x) {
      v_b = p.b[(size_t)qp * p.b_bs + t];
      v_lb = p.lb[(size_t)qp * p.lb_bs + t];
      v_ub = p.ub[(size_t)qp * p.ub_bs + t];
      if (!isinf(v_lb) || !isinf(v_ub)) finite_bounds = 1;
      if (p.warm) { v_x = p.x[(size_t)qp * n + t]; v_mux = p.mu_x[(size_t)qp * n + t]; }
    }
    if (t < m) beqs[t] = p.beq[(size_t)qp * p.beq_bs + t];
    if (t < n8) xs[t] = v_x;
    if (t < nc / 3) vmu[t] = p.mu[(size_t)qp * p.mu_bs + t];
    if (t < nc) muc[t] = p.warm ? p.mu_c[(size_t)qp * nc + t] : 0.0;
    for (int e = t; e < L.ndp8c + 8; e += kThreads) { vd[e] = 0.0; hinv[e] = 0.0; }
    const bool eqc = (nc == 0) && (__syncthreads_or(finite_bounds) == 0);   // fcc_qp.cpp:132-133
    const bool presolve = eqc || !p.warm;                                      // fcc_qp.cpp:159

    const long long t_start = clock64();
    unsigned long long fact_cycles = 0;
    int status_flag = 0;
    int n_iter = 0;
    double res_x = 0.0, res_c = 0.0;

    // ---------------- classification (reads all of Q and A_eq) ----------------
    VarClass vc;
    StructQP S;
    bool defer = struct_classify<kThreads>(n, m, m8, Qg, q_slow, q_fast, Ag, p.a_rs, p.a_cs, qd_s, I, L.nr8c, L.ndp8c,
                                           L.nd08c, vc, S.nr, S.ndp, S.nd0) != 0;
    S.nr8 = (S.nr + 7) & ~7;
    S.dpt = (S.ndp + 7) >> 3;
    const int nd08 = (S.nd0 + 7) & ~7;
    S.N8 = S.nr8 + m8 + nd08;
    S.NB = S.N8 >> 3;
    S.NB32 = (S.N8 + 31) >> 5;
    const int NBr = S.nr8 >> 3;
    const int yrow = t - S.nr8;

    // pass 0: cold pre-solve (shift 0);  pass 1: ADMM (shift rho).  Same lazy scheme as the general kernel:
    // x-update 0 of a cold solve is the identity, the rho-KKT system is only factored for QPs that iterate.
    for (int pass = presolve ? 0 : 1; pass < 2 && !defer; ++pass) {
      if (pass == 1 && eqc) break;
      if (pass == 1) {
        v_xbar = v_x;                             // fcc_qp.cpp:74-75
        if (in_cone) lcbar[t - lcs] = v_x;
        __syncthreads();
        n_iter = p.max_iter;
      }
      const int iters = pass == 0 ? 1 : p.max_iter;
      const double shift = pass == 0 ? 0.0 : p.rho;
      const double hi = (vc.type == VT_DP || vc.type == VT_D1) ? 1.0 / (vc.qd + shift) : 0.0;
      bool factored = false;
      bool full_inverse = false;

#pragma unroll 1
      for (int iter = 0; iter < iters; ++iter) {
        double val = v_x;
        if (!(pass == 1 && iter == 0 && presolve && p.first_update_identity)) {
          if (!factored) {
            factored = true;
            const long long t_f0 = clock64();
            // ---------------- assemble the reduced KKT matrix (lower tiles) and the D+ columns ----------------
            {
              const int nbt = (S.NB * (S.NB + 1)) >> 1;
              const double2 z2 = make_double2(0.0, 0.0);
              for (int e = t; e < nbt * 32; e += kThreads) st2(M + 2 * e, z2);
              for (int e = t; e < mt * dptc * 32; e += kThreads) st2(AP + 2 * e, z2);
            }
            __syncthreads();
            // Q_RR: one warp per reduced row
#pragma unroll 1
            for (int a = warp; a < S.nr; a += kWarps) {
              const double* qrow = Qg + (long long)I.rlist[a] * q_slow;
              for (int bb = lane; bb <= a; bb += 32) cp_async8(M + mat_off(a, bb), qrow + (long long)I.rlist[bb] * q_fast);
            }
            // A_eq: R columns -> constraint rows of the matrix, D0 columns -> transposed into the trailing rows,
            // D+ columns -> AP tiles (row tile I, column tile kt at (I * dptc + kt) * 64)
            if (p.a_cs == 1 || p.a_rs != 1) {
#pragma unroll 1
              for (int k = warp; k < m; k += kWarps) {
                const double* arow = Ag + (long long)k * p.a_rs;
                const int yr = S.nr8 + k;
                for (int j = lane; j < n; j += 32) {
                  const int ty = I.vtype[j], ps = I.vpos[j];
                  const double* src = arow + (long long)j * p.a_cs;
                  if (ty == VT_R) cp_async8(M + mat_off(yr, ps), src);
                  else if (ty == VT_D0) cp_async8(M + mat_off(ps, yr), src);
                  else if (ty == VT_DP) cp_async8(AP + (size_t)((k >> 3) * dptc + (ps >> 3)) * 64 + el_off(k & 7, ps & 7), src);
                }
              }
            } else {
              // column-major A_eq (Eigen callers): one warp per column, lanes along it
#pragma unroll 1
              for (int j = warp; j < n; j += kWarps) {
                const int ty = I.vtype[j], ps = I.vpos[j];
                if (ty == VT_D1) continue;
                const double* acol = Ag + (long long)j * p.a_cs;
                for (int k = lane; k < m; k += 32) {
                  const int yr = S.nr8 + k;
                  const double* src = acol + k;
                  if (ty == VT_R) cp_async8(M + mat_off(yr, ps), src);
                  else if (ty == VT_D0) cp_async8(M + mat_off(ps, yr), src);
                  else cp_async8(AP + (size_t)((k >> 3) * dptc + (ps >> 3)) * 64 + el_off(k & 7, ps & 7), src);
                }
              }
            }
            // decoupled unit pivots on the pads, cost of the zero-cost block, 1/h of the D+ columns
            if (t < S.N8) {
              const bool pad = (t >= S.nr && t < S.nr8) || (yrow >= m && yrow < m8) || (t >= S.nr8 + m8 + S.nd0);
              if (pad) M[mat_off(t, t)] = 1.0;
            }
            if (vc.type == VT_D0) M[mat_off(vc.pos, vc.pos)] = vc.qd + shift;
            if (vc.type == VT_DP) hinv[vc.pos] = hi;
            cp_async_wait_all();
            __syncthreads();
            if (t < S.nr && shift != 0.0) M[mat_off(t, t)] += shift;
            // -C on the constraint block: C_IJ = sum_kt AP_I,kt diag(1/h) AP_J,kt'  (one tile per warp and round)
            {
              const int ntile = (mt * (mt + 1)) >> 1;
#pragma unroll 1
              for (int w = warp; w < ntile; w += kWarps) {
                // (Ic, Jc) of the w-th lower tile
                int Ic = 0;
                while (((Ic + 1) * (Ic + 2)) >> 1 <= w) ++Ic;
                const int Jc = w - ((Ic * (Ic + 1)) >> 1);
                double2 ca = make_double2(0.0, 0.0), cb = make_double2(0.0, 0.0);
                const double* ai = AP + (size_t)(Ic * dptc) * 64 + fragC;
                const double* aj = AP + (size_t)(Jc * dptc) * 64 + fragC;
#pragma unroll 1
                for (int kt = 0; kt < S.dpt; ++kt) {
                  const double2 a = ld2(ai + 64 * kt), b = ld2(aj + 64 * kt), h = ld2(hinv + 8 * kt + 2 * fq);
                  dmma(ca.x, ca.y, a.x * h.x, b.x);
                  dmma(cb.x, cb.y, a.y * h.y, b.y);
                }
                st2(M + tile_off(NBr + Ic, NBr + Jc) + fragC, make_double2(-(ca.x + cb.x), -(ca.y + cb.y)));
              }
            }
            __syncthreads();
            // one-entry columns: a^2 / h on one diagonal entry each (at most one such column per row); the pad
            // rows of the constraint block get their unit pivot back (the tile store above overwrote it)
            if (vc.type == VT_D1 && vc.nnz == 1) M[mat_off(S.nr8 + vc.krow, S.nr8 + vc.krow)] -= vc.aval * vc.aval * hi;
            if (yrow >= m && yrow < m8) M[mat_off(t, t)] = 1.0;
            __syncthreads();
#ifdef FCCQP_DEV
            { unsigned long long* trbuf = nullptr; int trn = 0; factor_tiles<kThreads>(M, dinv, dneg, S.NB, S.NB32, trbuf, trn); }
#else
            factor_tiles<kThreads>(M, dinv, dneg, S.NB, S.NB32);
#endif
            fact_cycles += (unsigned long long)(clock64() - t_f0);
            // inertia (+ on R and D0 rows and all pads, - on the constraint rows): anything else goes to the general kernel
            {
              bool badp = false;
              if (t < S.N8) {
                const double dn = dneg[t];
                const bool neg = yrow >= 0 && yrow < m;
                badp = !isfinite(dn) || (neg ? !(dn > 0.0) : !(dn < 0.0));
              }
              if (__syncthreads_or(badp)) { defer = true; break; }
            }
          }  // lazy factorization

          // ---- right-hand side of this thread's variable (constraint rows: b_eq, in shared memory)
          double r = 0.0;
          if (is_x) {
            if (pass == 0) r = -v_b;
            else {
              // -(b + q_rho), q_rho = -rho (xbar - mu_x) with the cone segment overwritten (fcc_qp.cpp:81-83)
              const double w = in_cone ? (lcbar[t - lcs] - muc[t - lcs]) : (v_xbar - v_mux);
              r = -(v_b - p.rho * w);
            }
          }
          if (!full_inverse && pass == 1 && iter >= p.full_inverse_at) {
            // long-running QP: explicit inverse of the reduced matrix, every later x-update is one symmetric product
            __syncthreads();
            complete_inverse<kThreads>(M, S.NB);
            form_g<kThreads>(M, dinv, S.NB, S.NB, S.NB);
            full_inverse = true;
          }
          val = struct_xsolve<kThreads>(p, L, smem, I, S, vc, Qg, Ag, q_slow, q_fast, r, hi, shift, full_inverse,
                                        pass == 0 && p.struct_refine != 0);
        }  // x-update solve

        if (pass == 0) {
          v_x = val;
          if (t < n8) xs[t] = is_x ? val : 0.0;
          if (p.dbg_x0 && is_x) p.dbg_x0[(size_t)qp * n + t] = val;
          __syncthreads();
          continue;
        }

        // ---- K4 + K5 (identical to the general kernel)
        if (is_x) { xs[t] = val; v_x = val; }
        __syncthreads();
        double rx = 0.0, rc = 0.0;
        const bool relax = p.alpha != 1.0;
        if (is_x) {
          const double xh = relax ? fma(p.alpha, val, (1.0 - p.alpha) * v_xbar) : val;
          const double xb = clampd(xh + v_mux, v_lb, v_ub);
          v_xbar = xb;
          const double rr = xh - xb;
          v_mux += rr;
          rx = fabs(rr);
        }
        if (t < nc / 3) {  // lane per contact
          const int o = lcs + 3 * t;
          double x0 = xs[o], x1 = xs[o + 1], x2 = xs[o + 2];
          if (relax) {
            x0 = fma(p.alpha, x0, (1.0 - p.alpha) * lcbar[3 * t]);
            x1 = fma(p.alpha, x1, (1.0 - p.alpha) * lcbar[3 * t + 1]);
            x2 = fma(p.alpha, x2, (1.0 - p.alpha) * lcbar[3 * t + 2]);
          }
          double o0, o1, o2;
          project_cone3(x0 + muc[3 * t], x1 + muc[3 * t + 1], x2 + muc[3 * t + 2], vmu[t], o0, o1, o2);
          lcbar[3 * t] = o0; lcbar[3 * t + 1] = o1; lcbar[3 * t + 2] = o2;
          const double r0 = x0 - o0, r1 = x1 - o1, r2 = x2 - o2;
          muc[3 * t] += r0; muc[3 * t + 1] += r1; muc[3 * t + 2] += r2;
          rc = fmax(fabs(r0), fmax(fabs(r1), fabs(r2)));
        }
        if (rx != rx || rc != rc) status_flag = 2;
        const int conv = __syncthreads_and((rc < p.eps_fcone) && (rx < p.eps_bound));   // fcc_qp.cpp:105-109
        if (conv || iter + 1 == iters) {
          block_reduce2<false>(rx, rc, red, parity);
          res_x = rx; res_c = rc;
          if (conv) { n_iter = iter; break; }
        }
      }
    }

    if (defer) {
      // not reducible within the caps / wrong inertia: the general kernel takes this QP (device-side list)
      if (tid == 0) p.pending_list[atomicAdd(p.pending_count, 1u)] = qp;
      continue;
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
      const double rr = sqrt(xs[o] * xs[o] + xs[o + 1] * xs[o + 1]) - vmu[tid] * xs[o + 2];
      fv = rr > 0.0 ? rr : 0.0;
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
  }
}

}  // namespace fccqp
