/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference solver.
 * See fccqp_oracle.h for the usage rules and the parity-pinning statement.
 *
 * What is restated (reference file:line, relative to /root/reference):
 *   FCCQP::FCCQP          src/fcc_qp.cpp:24-55     -> fccqp_oracle_create
 *   FCCQP::Solve          src/fcc_qp.cpp:114-191   -> fccqp_oracle_solve
 *   FCCQP::DoADMM         src/fcc_qp.cpp:57-112    -> do_admm
 *   FCCQP::GetSolution    src/fcc_qp.cpp:194-207   -> fccqp_oracle_get_solution
 *   project_to_friction_cone / project_to_bounds / calc_*_violation
 *                         src/constraint_utils.cpp:5-65
 * and, because the arithmetic lives in the vendored Eigen 3.3.90
 * (libigl/eigen @1f05f51, pinned EXACT at CMakeLists.txt:10):
 *   LDLT::compute / ldlt_inplace<Lower>::unblocked   eigen/Eigen/src/Cholesky/LDLT.h:293-393,489-520
 *   LDLT::_solve_impl                                eigen/Eigen/src/Cholesky/LDLT.h:557-591
 *   ColPivHouseholderQR::computeInPlace / rank       eigen/Eigen/src/QR/ColPivHouseholderQR.h:479-581,255-263
 *   CompleteOrthogonalDecomposition::computeInPlace  eigen/Eigen/src/QR/CompleteOrthogonalDecomposition.h:410-465
 *   COD::_solve_impl / applyZAdjointOnTheLeftInPlace eigen/.../CompleteOrthogonalDecomposition.h:467-528
 *   makeHouseholder / applyHouseholderOnTheLeft/Right eigen/Eigen/src/Householder/Householder.h:63-164
 * The restatement follows the same algorithms step by step but is written in
 * plain scalar C; floating-point summation order differs from Eigen's
 * vectorised kernels, so agreement with the compiled reference is to rounding
 * (checked in tests/test_oracle.py), not bit-for-bit.
 */
#define _POSIX_C_SOURCE 200809L
#include "fccqp_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

struct fccqp_oracle {
  int n, m, nc, lcs, N;
  int max_iter;
  double rho, eps_fcone, eps_bound;
  int warm_start;
  /* workspace (fcc_qp.hpp:139-171) */
  double *M_kkt, *M_kkt_pre, *b_kkt, *kkt_sol;
  double *x, *x_bar, *lambda_c_bar, *mu_x, *mu_lambda_c, *q_rho, *x_res, *lambda_c_res;
  /* factorizations */
  double *ldlt, *ldlt_pre, *tmpN;
  int *tr, *tr_pre;
  int ldlt_pre_ok;
  double *qr, *hcoef, *zcoef, *cnu, *cnd, *tmpN2;
  int *ctr, *cperm;
  int qr_nonzero_pivots, qr_rank;
  double qr_maxpivot;
  /* results */
  double x_res_norm, lambda_c_res_norm, bounds_viol, friction_cone_viol;
  int n_iter;
  double solve_time, factorization_time;
  int presolve_path;
  /* test hook (not in the reference): residual history of the last DoADMM, [2 * iter] = bound residual,
   * [2 * iter + 1] = friction-cone residual -- what the exit test of fcc_qp.cpp:105 compared with eps */
  double* trace;
  int trace_cap;
};

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define AT(a, ld, i, j) ((a)[(size_t)(i) + (size_t)(j) * (size_t)(ld)])

/* ---------------- constraint_utils.cpp ---------------- */

/* constraint_utils.cpp:5-25 */
void fccqp_oracle_project_cone3(const double f[3], double mu, double out[3]) {
  double norm_fxy = sqrt(f[0] * f[0] + f[1] * f[1]);
  if (mu * f[2] >= norm_fxy) { /* inside the cone */
    out[0] = f[0]; out[1] = f[1]; out[2] = f[2];
    return;
  }
  if (f[2] < -mu * norm_fxy) { /* polar cone: origin */
    out[0] = out[1] = out[2] = 0.0;
    return;
  }
  double xy_ratio = mu * f[2] / norm_fxy;
  double ray[3] = {xy_ratio * f[0], xy_ratio * f[1], f[2]};
  /* Eigen normalize(): only divides when the squared norm is > 0 (Dot.h:142-148) */
  double sq = ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2];
  if (sq > 0.0) {
    double nr = sqrt(sq);
    ray[0] /= nr; ray[1] /= nr; ray[2] /= nr;
  }
  double d = ray[0] * f[0] + ray[1] * f[1] + ray[2] * f[2];
  out[0] = d * ray[0]; out[1] = d * ray[1]; out[2] = d * ray[2];
}

/* constraint_utils.cpp:48-59 */
double fccqp_oracle_cone_violation(const double* f, int nc, const double* mu) {
  double v = 0.0;
  for (int i = 0; i < nc / 3; ++i) {
    const double* g = f + 3 * i;
    double r = sqrt(g[0] * g[0] + g[1] * g[1]) - mu[i] * g[2];
    v += r > 0.0 ? r : 0.0;
  }
  return v;
}

/* constraint_utils.cpp:37-46 */
static double clampd(double x, double lb, double ub) {
  double t = x < ub ? x : ub; /* std::min(x, ub) */
  return t > lb ? t : lb;     /* std::max(., lb) */
}

/* constraint_utils.cpp:61-65 */
double fccqp_oracle_bound_violation(const double* x, const double* lb, const double* ub, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) {
    double d = x[i] - clampd(x[i], lb[i], ub[i]);
    s += d * d;
  }
  return sqrt(s);
}

/* ---------------- Eigen LDLT (lower, diagonal pivoting) ---------------- */

/* LDLT.h:293-393. Returns 1 for Success, 0 for NumericalIssue. */
static int ldlt_compute(int N, double* a, int* tr, double* temp) {
  if (N <= 1) {
    if (N == 1) tr[0] = 0;
    return 1;
  }
  int found_zero_pivot = 0, ret = 1;
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(AT(a, N, k, k));
    for (int i = k + 1; i < N; ++i) {
      double v = fabs(AT(a, N, i, i));
      if (v > best) { best = v; p = i; }
    }
    tr[k] = p;
    if (p != k) {
      for (int j = 0; j < k; ++j) {
        double t = AT(a, N, k, j); AT(a, N, k, j) = AT(a, N, p, j); AT(a, N, p, j) = t;
      }
      for (int i = p + 1; i < N; ++i) {
        double t = AT(a, N, i, k); AT(a, N, i, k) = AT(a, N, i, p); AT(a, N, i, p) = t;
      }
      double t = AT(a, N, k, k); AT(a, N, k, k) = AT(a, N, p, p); AT(a, N, p, p) = t;
      for (int i = k + 1; i < p; ++i) {
        double u = AT(a, N, i, k); AT(a, N, i, k) = AT(a, N, p, i); AT(a, N, p, i) = u;
      }
    }
    int rs = N - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; ++j) temp[j] = AT(a, N, j, j) * AT(a, N, k, j);
      double s = 0.0;
      for (int j = 0; j < k; ++j) s += AT(a, N, k, j) * temp[j];
      AT(a, N, k, k) -= s;
      for (int i = k + 1; i < N; ++i) {
        double acc = 0.0;
        for (int j = 0; j < k; ++j) acc += AT(a, N, i, j) * temp[j];
        AT(a, N, i, k) -= acc;
      }
    }
    double akk = AT(a, N, k, k);
    int pivot_is_valid = fabs(akk) > 0.0;
    if (k == 0 && !pivot_is_valid) {
      for (int j = 0; j < N; ++j) {
        tr[j] = j;
        for (int i = j + 1; i < N; ++i) ret = ret && (AT(a, N, i, j) == 0.0);
      }
      return ret;
    }
    if (rs > 0 && pivot_is_valid)
      for (int i = k + 1; i < N; ++i) AT(a, N, i, k) /= akk;
    if (found_zero_pivot && pivot_is_valid) ret = 0;
    else if (!pivot_is_valid) found_zero_pivot = 1;
  }
  return ret;
}

/* LDLT.h:557-591; rhs -> dst (may alias) */
static void ldlt_solve(int N, const double* a, const int* tr, const double* rhs, double* dst) {
  if (dst != rhs) memcpy(dst, rhs, sizeof(double) * (size_t)N);
  for (int k = 0; k < N; ++k)
    if (tr[k] != k) { double t = dst[k]; dst[k] = dst[tr[k]]; dst[tr[k]] = t; }
  for (int j = 0; j < N; ++j) { /* unit lower forward substitution */
    double xj = dst[j];
    if (xj != 0.0)
      for (int i = j + 1; i < N; ++i) dst[i] -= AT(a, N, i, j) * xj;
  }
  const double tol = 1.0 / DBL_MAX;
  for (int i = 0; i < N; ++i) {
    double d = AT(a, N, i, i);
    if (fabs(d) > tol) dst[i] /= d; else dst[i] = 0.0;
  }
  for (int i = N - 1; i >= 0; --i) { /* unit upper (L^T) back substitution */
    double s = dst[i];
    for (int j = i + 1; j < N; ++j) s -= AT(a, N, j, i) * dst[j];
    dst[i] = s;
  }
  for (int k = N - 1; k >= 0; --k)
    if (tr[k] != k) { double t = dst[k]; dst[k] = dst[tr[k]]; dst[tr[k]] = t; }
}

/* ---------------- Eigen Householder + ColPivQR + COD (square N x N) ---------------- */

/* Householder.h:63-94 on a strided vector v[0..len): in place, returns tau, beta. */
static void make_householder_inplace(double* v, size_t stride, int len, double* tau, double* beta) {
  double tail_sq = 0.0;
  for (int i = 1; i < len; ++i) tail_sq += v[i * stride] * v[i * stride];
  double c0 = v[0];
  if (len == 1 || tail_sq <= DBL_MIN) {
    *tau = 0.0; *beta = c0;
    for (int i = 1; i < len; ++i) v[i * stride] = 0.0;
  } else {
    double bt = sqrt(c0 * c0 + tail_sq);
    if (c0 >= 0.0) bt = -bt;
    for (int i = 1; i < len; ++i) v[i * stride] /= (c0 - bt);
    *tau = (bt - c0) / bt;
    *beta = bt;
  }
}

/* ColPivHouseholderQR.h:479-581 */
static void colpivqr_compute(fccqp_oracle* o) {
  const int N = o->N;
  double* qr = o->qr;
  for (int k = 0; k < N; ++k) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += AT(qr, N, i, k) * AT(qr, N, i, k);
    o->cnd[k] = o->cnu[k] = sqrt(s);
  }
  double mx = 0.0;
  for (int k = 0; k < N; ++k) if (o->cnu[k] > mx) mx = o->cnu[k];
  double th = mx * DBL_EPSILON;
  const double threshold_helper = th * th / (double)N;
  const double norm_downdate_threshold = sqrt(DBL_EPSILON);
  o->qr_nonzero_pivots = N;
  o->qr_maxpivot = 0.0;
  for (int k = 0; k < N; ++k) {
    int big = k;
    double bn = o->cnu[k];
    for (int j = k + 1; j < N; ++j) if (o->cnu[j] > bn) { bn = o->cnu[j]; big = j; }
    double big_sq = bn * bn;
    if (o->qr_nonzero_pivots == N && big_sq < threshold_helper * (double)(N - k))
      o->qr_nonzero_pivots = k;
    o->ctr[k] = big;
    if (big != k) {
      for (int i = 0; i < N; ++i) {
        double t = AT(qr, N, i, k); AT(qr, N, i, k) = AT(qr, N, i, big); AT(qr, N, i, big) = t;
      }
      double t = o->cnu[k]; o->cnu[k] = o->cnu[big]; o->cnu[big] = t;
      t = o->cnd[k]; o->cnd[k] = o->cnd[big]; o->cnd[big] = t;
    }
    double tau, beta;
    make_householder_inplace(&AT(qr, N, k, k), 1, N - k, &tau, &beta);
    o->hcoef[k] = tau;
    AT(qr, N, k, k) = beta;
    if (fabs(beta) > o->qr_maxpivot) o->qr_maxpivot = fabs(beta);
    /* applyHouseholderOnTheLeft to bottomRightCorner(N-k, N-k-1) */
    int rows = N - k, cols = N - k - 1;
    if (cols > 0) {
      if (rows == 1) {
        for (int j = k + 1; j < N; ++j) AT(qr, N, k, j) *= (1.0 - tau);
      } else if (tau != 0.0) {
        for (int j = k + 1; j < N; ++j) {
          double t = 0.0;
          for (int i = k + 1; i < N; ++i) t += AT(qr, N, i, k) * AT(qr, N, i, j);
          t += AT(qr, N, k, j);
          AT(qr, N, k, j) -= tau * t;
          for (int i = k + 1; i < N; ++i) AT(qr, N, i, j) -= tau * AT(qr, N, i, k) * t;
        }
      }
    }
    for (int j = k + 1; j < N; ++j) {
      if (o->cnu[j] != 0.0) {
        double temp = fabs(AT(qr, N, k, j)) / o->cnu[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0.0 ? 0.0 : temp;
        double r = o->cnu[j] / o->cnd[j];
        double temp2 = temp * r * r;
        if (temp2 <= norm_downdate_threshold) {
          double s = 0.0;
          for (int i = k + 1; i < N; ++i) s += AT(qr, N, i, j) * AT(qr, N, i, j);
          o->cnd[j] = o->cnu[j] = sqrt(s);
        } else {
          o->cnu[j] *= sqrt(temp);
        }
      }
    }
  }
  /* colsPermutation: identity with transpositions applied on the right */
  for (int j = 0; j < N; ++j) o->cperm[j] = j;
  for (int k = 0; k < N; ++k) {
    int t = o->cperm[k]; o->cperm[k] = o->cperm[o->ctr[k]]; o->cperm[o->ctr[k]] = t;
  }
  /* rank(): ColPivHouseholderQR.h:255-263 with threshold eps*diagonalSize (:378-384) */
  double pre = fabs(o->qr_maxpivot) * (DBL_EPSILON * (double)N);
  int rank = 0;
  for (int i = 0; i < o->qr_nonzero_pivots; ++i) rank += fabs(AT(qr, N, i, i)) > pre;
  o->qr_rank = rank;
}

/* CompleteOrthogonalDecomposition.h:410-465 */
static void cod_compute(fccqp_oracle* o) {
  const int N = o->N;
  double* qr = o->qr;
  colpivqr_compute(o);
  const int rank = o->qr_rank, cols = N;
  if (rank < cols) {
    for (int k = rank - 1; k >= 0; --k) {
      if (k != rank - 1)
        for (int i = 0; i <= k; ++i) {
          double t = AT(qr, N, i, k); AT(qr, N, i, k) = AT(qr, N, i, rank - 1); AT(qr, N, i, rank - 1) = t;
        }
      double beta, tau;
      /* row(k).tail(cols-rank+1) is strided by N in column-major storage */
      make_householder_inplace(&AT(qr, N, k, rank - 1), (size_t)N, cols - rank + 1, &tau, &beta);
      o->zcoef[k] = tau;
      AT(qr, N, k, rank - 1) = beta;
      if (k > 0) {
        /* topRightCorner(k, cols-rank+1).applyHouseholderOnTheRight(essential = row(k).tail(cols-rank)) */
        int bc = cols - rank + 1;
        if (bc == 1) {
          for (int i = 0; i < k; ++i) AT(qr, N, i, rank - 1) *= (1.0 - tau);
        } else if (tau != 0.0) {
          for (int i = 0; i < k; ++i) {
            double t = 0.0;
            for (int j = rank; j < cols; ++j) t += AT(qr, N, i, j) * AT(qr, N, k, j);
            t += AT(qr, N, i, rank - 1);
            AT(qr, N, i, rank - 1) -= tau * t;
            for (int j = rank; j < cols; ++j) AT(qr, N, i, j) -= tau * t * AT(qr, N, k, j);
          }
        }
      }
      if (k != rank - 1)
        for (int i = 0; i <= k; ++i) {
          double t = AT(qr, N, i, k); AT(qr, N, i, k) = AT(qr, N, i, rank - 1); AT(qr, N, i, rank - 1) = t;
        }
    }
  }
}

/* CompleteOrthogonalDecomposition.h:492-528 (+ :467-486) */
static void cod_solve(fccqp_oracle* o, const double* rhs, double* dst) {
  const int N = o->N, rank = o->qr_rank;
  const double* qr = o->qr;
  if (rank == 0) { memset(dst, 0, sizeof(double) * (size_t)N); return; }
  double* c = o->tmpN2;
  memcpy(c, rhs, sizeof(double) * (size_t)N);
  /* c = Q^T rhs: apply H_0, H_1, ..., H_{rank-1} in this order */
  for (int k = 0; k < rank; ++k) {
    double tau = o->hcoef[k];
    int rows = N - k;
    if (rows == 1) { c[k] *= (1.0 - tau); }
    else if (tau != 0.0) {
      double t = c[k];
      for (int i = k + 1; i < N; ++i) t += AT(qr, N, i, k) * c[i];
      c[k] -= tau * t;
      for (int i = k + 1; i < N; ++i) c[i] -= tau * AT(qr, N, i, k) * t;
    }
  }
  /* solve T z = c(0:rank) with T upper triangular rank x rank */
  for (int i = rank - 1; i >= 0; --i) {
    double s = c[i];
    for (int j = i + 1; j < rank; ++j) s -= AT(qr, N, i, j) * dst[j];
    dst[i] = s / AT(qr, N, i, i);
  }
  if (rank < N) {
    for (int i = rank; i < N; ++i) dst[i] = 0.0;
    /* applyZAdjointOnTheLeftInPlace */
    for (int k = 0; k < rank; ++k) {
      if (k != rank - 1) { double t = dst[k]; dst[k] = dst[rank - 1]; dst[rank - 1] = t; }
      double tau = o->zcoef[k];
      int rows = N - rank + 1;
      if (rows == 1) { dst[rank - 1] *= (1.0 - tau); }
      else if (tau != 0.0) {
        double t = dst[rank - 1];
        for (int j = rank; j < N; ++j) t += AT(qr, N, k, j) * dst[j];
        dst[rank - 1] -= tau * t;
        for (int j = rank; j < N; ++j) dst[j] -= tau * AT(qr, N, k, j) * t;
      }
      if (k != rank - 1) { double t = dst[k]; dst[k] = dst[rank - 1]; dst[rank - 1] = t; }
    }
  }
  /* dst = colsPermutation * dst : out[perm[i]] = y[i] */
  memcpy(c, dst, sizeof(double) * (size_t)N);
  for (int i = 0; i < N; ++i) dst[o->cperm[i]] = c[i];
}

/* ---------------- FCCQP ---------------- */

static double* dalloc(size_t k) { return (double*)calloc(k ? k : 1, sizeof(double)); }

/* fcc_qp.cpp:24-55 */
fccqp_oracle* fccqp_oracle_create(int n, int m, int nc, int lcs) {
  if (n < 0 || m < 0 || nc < 0 || nc % 3 != 0 || lcs < 0 || lcs + nc > n) return NULL;
  fccqp_oracle* o = (fccqp_oracle*)calloc(1, sizeof(*o));
  o->n = n; o->m = m; o->nc = nc; o->lcs = lcs; o->N = n + m;
  o->max_iter = 1000; o->rho = 1e-6; o->eps_fcone = 1e-3; o->eps_bound = 1e-6; /* hpp:30-35 */
  size_t N = (size_t)o->N;
  o->M_kkt = dalloc(N * N); o->M_kkt_pre = dalloc(N * N);
  o->ldlt = dalloc(N * N); o->ldlt_pre = dalloc(N * N); o->qr = dalloc(N * N);
  o->b_kkt = dalloc(N); o->kkt_sol = dalloc(N); o->tmpN = dalloc(N); o->tmpN2 = dalloc(N);
  o->hcoef = dalloc(N); o->zcoef = dalloc(N); o->cnu = dalloc(N); o->cnd = dalloc(N);
  o->tr = (int*)calloc(N ? N : 1, sizeof(int)); o->tr_pre = (int*)calloc(N ? N : 1, sizeof(int));
  o->ctr = (int*)calloc(N ? N : 1, sizeof(int)); o->cperm = (int*)calloc(N ? N : 1, sizeof(int));
  o->x = dalloc(n); o->x_bar = dalloc(n); o->mu_x = dalloc(n); o->q_rho = dalloc(n);
  o->x_res = dalloc(n);
  o->lambda_c_bar = dalloc(nc); o->mu_lambda_c = dalloc(nc); o->lambda_c_res = dalloc(nc);
  return o;
}

void fccqp_oracle_destroy(fccqp_oracle* o) {
  if (!o) return;
  free(o->M_kkt); free(o->M_kkt_pre); free(o->ldlt); free(o->ldlt_pre); free(o->qr);
  free(o->b_kkt); free(o->kkt_sol); free(o->tmpN); free(o->tmpN2); free(o->hcoef);
  free(o->zcoef); free(o->cnu); free(o->cnd); free(o->tr); free(o->tr_pre); free(o->ctr);
  free(o->cperm); free(o->x); free(o->x_bar); free(o->mu_x); free(o->q_rho); free(o->x_res);
  free(o->lambda_c_bar); free(o->mu_lambda_c); free(o->lambda_c_res);
  free(o);
}

void fccqp_oracle_set_options(fccqp_oracle* o, int max_iter, double rho, double eps_fcone,
                              double eps_bound) {
  o->max_iter = max_iter; o->rho = rho; o->eps_fcone = eps_fcone; o->eps_bound = eps_bound;
}
void fccqp_oracle_set_warm_start(fccqp_oracle* o, int warm) { o->warm_start = warm != 0; }

static double inf_norm_like_reference(const double* v, int len) {
  /* abs(max(maxCoeff, -minCoeff)); the reference crashes on len == 0
   * (fcc_qp.cpp:98) -- defined here as 0. */
  if (len == 0) return 0.0;
  double mx = v[0], mn = v[0];
  for (int i = 1; i < len; ++i) { if (v[i] > mx) mx = v[i]; if (v[i] < mn) mn = v[i]; }
  double r = mx > -mn ? mx : -mn;
  return fabs(r);
}

/* fcc_qp.cpp:57-112 */
/* process-wide over-relaxation of the restatement (tests of the product's extension only; default 1 = reference) */
static double g_relaxation = 1.0;
void fccqp_oracle_set_relaxation(double alpha) { g_relaxation = alpha; }
/* Extension of the product (fccqp_options::adapt_rho_interval), NOT in the reference: every `interval` ADMM iterations
 * rho is rebalanced from the primal / dual residual ratio (0 = off = the reference's fixed rho).  Process-wide. */
static int g_adapt_interval = 0;
void fccqp_oracle_set_adaptive_rho(int interval) { g_adapt_interval = interval; }
/* residual history of the following solves goes to buf[2 * cap] (NULL: off); caller-owned */
void fccqp_oracle_set_trace(fccqp_oracle* o, double* buf, int cap) { o->trace = buf; o->trace_cap = cap; }

static void do_admm(fccqp_oracle* o, const double* b, const double* mu, const double* lb,
                    const double* ub) {
  const int n = o->n, N = o->N, nc = o->nc, lcs = o->lcs;
  memcpy(o->M_kkt, o->M_kkt_pre, sizeof(double) * (size_t)N * N);
  for (int i = 0; i < n; ++i) AT(o->M_kkt, N, i, i) += o->rho;

  double t0 = now_s();
  memcpy(o->ldlt, o->M_kkt, sizeof(double) * (size_t)N * N);
  ldlt_compute(N, o->ldlt, o->tr, o->tmpN);
  o->factorization_time += now_s() - t0;

  memcpy(o->x_bar, o->x, sizeof(double) * (size_t)n);
  memcpy(o->lambda_c_bar, o->x + lcs, sizeof(double) * (size_t)nc);

  o->n_iter = o->max_iter;
  double rho = o->rho;                       /* (changes only with the adaptive-rho extension) */
  double* const xbar_prev = o->tmpN2;        /* [N >= n]: x_bar of the previous iteration */
  double* const lcbar_prev = (double*)malloc(sizeof(double) * (size_t)(nc ? nc : 1));
  for (int iter = 0; iter < o->max_iter; ++iter) {
    for (int i = 0; i < n; ++i) o->q_rho[i] = -rho * (o->x_bar[i] - o->mu_x[i]);
    for (int i = 0; i < nc; ++i)
      o->q_rho[lcs + i] = -rho * (o->lambda_c_bar[i] - o->mu_lambda_c[i]);
    if (g_adapt_interval > 0) {
      memcpy(xbar_prev, o->x_bar, sizeof(double) * (size_t)n);
      memcpy(lcbar_prev, o->lambda_c_bar, sizeof(double) * (size_t)nc);
    }
    for (int i = 0; i < n; ++i) o->b_kkt[i] = -(b[i] + o->q_rho[i]);

    ldlt_solve(N, o->ldlt, o->tr, o->b_kkt, o->kkt_sol);
    memcpy(o->x, o->kkt_sol, sizeof(double) * (size_t)n);

    /* Extension of the product (include/fccqp.h, fccqp_options::relaxation), NOT in the reference:
     * x_hat = alpha x + (1 - alpha) x_bar_prev takes the place of x in the z-update, the residuals and the
     * dual update.  alpha == 1 (the default) is the reference's code path, statement for statement. */
    const double alpha = g_relaxation;
    if (alpha == 1.0) {
      for (int i = 0; i < n; ++i) o->x_bar[i] = clampd(o->x[i] + o->mu_x[i], lb[i], ub[i]);
      for (int c = 0; c < nc / 3; ++c) {
        double f[3];
        for (int k = 0; k < 3; ++k) f[k] = o->x[lcs + 3 * c + k] + o->mu_lambda_c[3 * c + k];
        fccqp_oracle_project_cone3(f, mu[c], o->lambda_c_bar + 3 * c);
      }
      for (int i = 0; i < n; ++i) o->x_res[i] = o->x[i] - o->x_bar[i];
      for (int i = 0; i < nc; ++i) o->lambda_c_res[i] = o->x[lcs + i] - o->lambda_c_bar[i];
    } else {
      for (int i = 0; i < n; ++i) {
        const double xh = fma(alpha, o->x[i], (1.0 - alpha) * o->x_bar[i]);
        o->x_bar[i] = clampd(xh + o->mu_x[i], lb[i], ub[i]);
        o->x_res[i] = xh - o->x_bar[i];
      }
      for (int c = 0; c < nc / 3; ++c) {
        double f[3], lh[3];
        for (int k = 0; k < 3; ++k) {
          lh[k] = fma(alpha, o->x[lcs + 3 * c + k], (1.0 - alpha) * o->lambda_c_bar[3 * c + k]);
          f[k] = lh[k] + o->mu_lambda_c[3 * c + k];
        }
        fccqp_oracle_project_cone3(f, mu[c], o->lambda_c_bar + 3 * c);
        for (int k = 0; k < 3; ++k) o->lambda_c_res[3 * c + k] = lh[k] - o->lambda_c_bar[3 * c + k];
      }
    }
    o->x_res_norm = inf_norm_like_reference(o->x_res, n);
    o->lambda_c_res_norm = inf_norm_like_reference(o->lambda_c_res, nc);
    if (o->trace && iter < o->trace_cap) {
      o->trace[2 * iter] = o->x_res_norm;
      o->trace[2 * iter + 1] = o->lambda_c_res_norm;
    }

    for (int i = 0; i < n; ++i) o->mu_x[i] += o->x_res[i];
    for (int i = 0; i < nc; ++i) o->mu_lambda_c[i] += o->lambda_c_res[i];

    if (o->lambda_c_res_norm < o->eps_fcone && o->x_res_norm < o->eps_bound) {
      o->n_iter = iter;
      break;
    }
    /* Adaptive rho (extension; include/fccqp.h, fccqp_options::adapt_rho_interval): with scaled duals mu = y / rho the
     * primal residual is r_p = max(|x_hat - x_bar|, |lambda_hat - lambda_bar|) and the dual one r_d = rho |z_k - z_{k-1}|;
     * when they are more than a factor 5 apart rho moves by sqrt(r_p / r_d) (at most 10x per update), the scaled duals
     * are rescaled so that y stays put, and the rho-KKT matrix is factored again. */
    if (g_adapt_interval > 0 && (iter + 1) % g_adapt_interval == 0 && iter + 1 < o->max_iter) {
      double dz = 0.0;
      for (int i = 0; i < n; ++i) { const double d = fabs(o->x_bar[i] - xbar_prev[i]); if (d > dz) dz = d; }
      for (int i = 0; i < nc; ++i) { const double d = fabs(o->lambda_c_bar[i] - lcbar_prev[i]); if (d > dz) dz = d; }
      const double rp = o->x_res_norm > o->lambda_c_res_norm ? o->x_res_norm : o->lambda_c_res_norm;
      const double rd = rho * dz;
      double ratio = sqrt(rp / (rd > 1e-300 ? rd : 1e-300));
      if (ratio > 10.0) ratio = 10.0;
      if (ratio < 0.1) ratio = 0.1;
      if (ratio > 5.0 || ratio < 0.2) {
        double rho_new = rho * ratio;
        if (rho_new > 1e9) rho_new = 1e9;
        if (rho_new < 1e-9) rho_new = 1e-9;
        const double sc = rho / rho_new;
        for (int i = 0; i < n; ++i) o->mu_x[i] *= sc;
        for (int i = 0; i < nc; ++i) o->mu_lambda_c[i] *= sc;
        rho = rho_new;
        memcpy(o->M_kkt, o->M_kkt_pre, sizeof(double) * (size_t)N * N);
        for (int i = 0; i < n; ++i) AT(o->M_kkt, N, i, i) += rho;
        memcpy(o->ldlt, o->M_kkt, sizeof(double) * (size_t)N * N);
        ldlt_compute(N, o->ldlt, o->tr, o->tmpN);
      }
    }
  }
  free(lcbar_prev);
}

/* fcc_qp.cpp:114-191 */
int fccqp_oracle_solve(fccqp_oracle* o, const double* Q, const double* b, const double* A,
                       const double* beq, const double* mu, int nmu, const double* lb,
                       const double* ub) {
  const int n = o->n, m = o->m, N = o->N, nc = o->nc, lcs = o->lcs;
  if (nmu < nc / 3) return -1; /* friction_coeffs.at(i) -> std::out_of_range */
  double start = now_s();

  int equality_constrained = (nc == 0);
  for (int i = 0; i < n && equality_constrained; ++i)
    if (!isinf(lb[i]) || !isinf(ub[i])) equality_constrained = 0;

  if (!o->warm_start) {
    memset(o->mu_x, 0, sizeof(double) * (size_t)n);
    memset(o->mu_lambda_c, 0, sizeof(double) * (size_t)nc);
  }
  memset(o->M_kkt_pre, 0, sizeof(double) * (size_t)N * N);
  memset(o->M_kkt, 0, sizeof(double) * (size_t)N * N);
  memset(o->b_kkt, 0, sizeof(double) * (size_t)N);

  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) AT(o->M_kkt_pre, N, i, j) = AT(Q, n, i, j);
    for (int i = 0; i < m; ++i) {
      AT(o->M_kkt_pre, N, n + i, j) = AT(A, m, i, j);
      AT(o->M_kkt_pre, N, j, n + i) = AT(A, m, i, j);
    }
  }
  for (int i = 0; i < n; ++i) o->b_kkt[i] = -b[i];
  for (int i = 0; i < m; ++i) o->b_kkt[n + i] = beq[i];

  o->factorization_time = 0; o->n_iter = 0; o->x_res_norm = 0; o->lambda_c_res_norm = 0;
  o->presolve_path = 0;

  if (equality_constrained || !o->warm_start) {
    double t0 = now_s();
    memcpy(o->ldlt_pre, o->M_kkt_pre, sizeof(double) * (size_t)N * N);
    o->ldlt_pre_ok = ldlt_compute(N, o->ldlt_pre, o->tr_pre, o->tmpN);
    if (!o->ldlt_pre_ok) {
      memcpy(o->qr, o->M_kkt_pre, sizeof(double) * (size_t)N * N);
      cod_compute(o);
    }
    o->factorization_time += now_s() - t0;
    if (o->ldlt_pre_ok) {
      ldlt_solve(N, o->ldlt_pre, o->tr_pre, o->b_kkt, o->kkt_sol);
      o->presolve_path = 1;
    } else {
      cod_solve(o, o->b_kkt, o->kkt_sol);
      o->presolve_path = 2;
    }
    memcpy(o->x, o->kkt_sol, sizeof(double) * (size_t)n);
  }

  if (!equality_constrained) do_admm(o, b, mu, lb, ub);

  o->bounds_viol = fccqp_oracle_bound_violation(o->x, lb, ub, n);
  o->friction_cone_viol = fccqp_oracle_cone_violation(o->x + lcs, nc, mu);
  o->solve_time = now_s() - start;
  return 0;
}

/* fcc_qp.cpp:194-207 */
void fccqp_oracle_get_solution(const fccqp_oracle* o, double* z, int* n_iter, int* status,
                               double* details6) {
  memcpy(z, o->x, sizeof(double) * (size_t)o->n);
  *n_iter = o->n_iter;
  *status = (o->n_iter == o->max_iter) ? 1 : 0;
  details6[0] = o->x_res_norm;
  details6[1] = o->lambda_c_res_norm;
  details6[2] = o->bounds_viol;
  details6[3] = o->friction_cone_viol;
  details6[4] = o->solve_time;
  details6[5] = o->factorization_time;
}

int fccqp_oracle_presolve_path(const fccqp_oracle* o) { return o->presolve_path; }

void fccqp_oracle_get_state(const fccqp_oracle* o, double* x, double* mu_x, double* mu_c) {
  memcpy(x, o->x, sizeof(double) * (size_t)o->n);
  memcpy(mu_x, o->mu_x, sizeof(double) * (size_t)o->n);
  memcpy(mu_c, o->mu_lambda_c, sizeof(double) * (size_t)o->nc);
}
void fccqp_oracle_set_state(fccqp_oracle* o, const double* x, const double* mu_x,
                            const double* mu_c) {
  memcpy(o->x, x, sizeof(double) * (size_t)o->n);
  memcpy(o->mu_x, mu_x, sizeof(double) * (size_t)o->n);
  memcpy(o->mu_lambda_c, mu_c, sizeof(double) * (size_t)o->nc);
}

/* ---------------- batch drivers (same contract as oracle/ref_shim.cpp) ---------------- */

typedef struct {
  int t, nthreads, B, n, m, nc, lcs, max_iter, warm_mode;
  double rho, eps_fcone, eps_bound;
  const double *Q, *b, *A, *beq, *mu, *lb, *ub;
  long mu_stride, bound_stride;
  double* z; int* n_iter; int* status; double* details6;
  double elapsed;
  pthread_barrier_t* bar;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  long lo = (long)j->B * j->t / j->nthreads, hi = (long)j->B * (j->t + 1) / j->nthreads;
  fccqp_oracle* o = fccqp_oracle_create(j->n, j->m, j->nc, j->lcs);
  fccqp_oracle_set_options(o, j->max_iter, j->rho, j->eps_fcone, j->eps_bound);
  pthread_barrier_wait(j->bar);
  double t0 = now_s();
  const long n = j->n, m = j->m;
  for (long i = lo; i < hi; ++i) {
    o->warm_start = (j->warm_mode == 1 && i > lo);
    fccqp_oracle_solve(o, j->Q + i * n * n, j->b + i * n, j->A + i * m * n, j->beq + i * m,
                       j->mu + i * j->mu_stride, j->nc / 3, j->lb + i * j->bound_stride,
                       j->ub + i * j->bound_stride);
    fccqp_oracle_get_solution(o, j->z + i * n, j->n_iter + i, j->status + i, j->details6 + i * 6);
  }
  j->elapsed = now_s() - t0;
  fccqp_oracle_destroy(o);
  return NULL;
}

double fccqp_oracle_solve_batch(int B, int n, int m, int nc, int lcs, int max_iter, double rho,
                                double eps_fcone, double eps_bound, int warm_mode, int nthreads,
                                const double* Qcm, const double* b, const double* Acm,
                                const double* beq, const double* mu, long mu_stride,
                                const double* lb, const double* ub, long bound_stride, double* z,
                                int* n_iter, int* status, double* details6) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = B > 0 ? B : 1;
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
  batch_job* jobs = (batch_job*)calloc((size_t)nthreads, sizeof(batch_job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) {
    batch_job j = {t, nthreads, B, n, m, nc, lcs, max_iter, warm_mode, rho, eps_fcone, eps_bound,
                   Qcm, b, Acm, beq, mu, lb, ub, mu_stride, bound_stride, z, n_iter, status,
                   details6, 0.0, &bar};
    jobs[t] = j;
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  double mx = 0.0;
  for (int t = 0; t < nthreads; ++t) {
    pthread_join(th[t], NULL);
    if (jobs[t].elapsed > mx) mx = jobs[t].elapsed;
  }
  pthread_barrier_destroy(&bar);
  free(jobs); free(th);
  return mx;
}

void fccqp_oracle_solve_lanes(void** handles, int B, int warm, const double* Qcm, const double* b,
                              const double* Acm, const double* beq, const double* mu,
                              long mu_stride, const double* lb, const double* ub,
                              long bound_stride, double* z, int* n_iter, int* status,
                              double* details6) {
  for (long i = 0; i < B; ++i) {
    fccqp_oracle* o = (fccqp_oracle*)handles[i];
    const long n = o->n, m = o->m;
    o->warm_start = warm != 0;
    fccqp_oracle_solve(o, Qcm + i * n * n, b + i * n, Acm + i * m * n, beq + i * m,
                       mu + i * mu_stride, o->nc / 3, lb + i * bound_stride, ub + i * bound_stride);
    fccqp_oracle_get_solution(o, z + i * n, n_iter + i, status + i, details6 + i * 6);
  }
}

int fccqp_oracle_hardware_threads(void) {
  long k = sysconf(_SC_NPROCESSORS_ONLN);
  return k > 0 ? (int)k : 1;
}
