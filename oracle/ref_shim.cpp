// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// C-ABI shim around the UNMODIFIED reference solver (fcc_qp::FCCQP,
// /root/reference/src/fcc_qp.hpp:54-171).  The reference sources are compiled
// where they lie (see oracle/Makefile); nothing from /root/reference is copied
// into this repository.  The resulting oracle/_ref/libfccqp_ref.so is the
// primary parity oracle and the "reference" CPU baseline of bench.py.
//
// Matrix arguments are COLUMN-major (Eigen's native layout) so that the
// reference sees zero-copy Ref<const MatrixXd> views, exactly like a C++
// caller of the reference would pass them.
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "fcc_qp.hpp"

using fcc_qp::FCCQP;
using fcc_qp::FCCQPOptions;
using fcc_qp::FCCQPSolution;
using Eigen::Map;
using Eigen::MatrixXd;
using Eigen::VectorXd;

namespace {
struct RefDims { int n, m, nc, lcs; };
struct RefHandle {
  RefDims d;
  FCCQP solver;
  RefHandle(int n, int m, int nc, int lcs) : d{n, m, nc, lcs}, solver(n, m, nc, lcs) {}
};

inline void solve_one(RefHandle* h, const double* Qcm, const double* b,
                      const double* Acm, const double* beq, const double* mu,
                      int nmu, const double* lb, const double* ub) {
  const RefDims& d = h->d;
  Map<const MatrixXd> Q(Qcm, d.n, d.n);
  Map<const MatrixXd> A(Acm, d.m, d.n);
  Map<const VectorXd> bv(b, d.n), beqv(beq, d.m), lbv(lb, d.n), ubv(ub, d.n);
  std::vector<double> fr(mu, mu + nmu);
  h->solver.Solve(Q, bv, A, beqv, fr, lbv, ubv);
}

inline void fetch(RefHandle* h, double* z, int* n_iter, int* status, double* details6) {
  FCCQPSolution s = h->solver.GetSolution();
  for (int i = 0; i < h->d.n; ++i) z[i] = s.z(i);
  *n_iter = s.details.n_iter;
  *status = static_cast<int>(s.details.solve_status);
  details6[0] = s.details.admm_residual_bounds;
  details6[1] = s.details.admm_residual_friction_cone;
  details6[2] = s.details.bounds_viol;
  details6[3] = s.details.friction_cone_viol;
  details6[4] = s.details.solve_time;
  details6[5] = s.details.factorization_time;
}
}  // namespace

extern "C" {

void* fccqp_ref_create(int n, int m, int nc, int lcs) { return new RefHandle(n, m, nc, lcs); }
void fccqp_ref_destroy(void* h) { delete static_cast<RefHandle*>(h); }

void fccqp_ref_set_options(void* h, int max_iter, double rho, double eps_fcone, double eps_bound) {
  FCCQPOptions o;
  o.max_iter = max_iter; o.rho = rho; o.eps_fcone = eps_fcone; o.eps_bound = eps_bound;
  static_cast<RefHandle*>(h)->solver.set_options(o);
}
void fccqp_ref_set_warm_start(void* h, int warm) {
  static_cast<RefHandle*>(h)->solver.set_warm_start(warm != 0);
}
void fccqp_ref_solve(void* h, const double* Qcm, const double* b, const double* Acm,
                     const double* beq, const double* mu, int nmu, const double* lb,
                     const double* ub) {
  solve_one(static_cast<RefHandle*>(h), Qcm, b, Acm, beq, mu, nmu, lb, ub);
}
void fccqp_ref_get_solution(void* h, double* z, int* n_iter, int* status, double* details6) {
  fetch(static_cast<RefHandle*>(h), z, n_iter, status, details6);
}

// Solve B stacked QPs with `nthreads` std::threads, one reference FCCQP object
// per thread, static contiguous partition.
//   warm_mode 0: every QP cold (set_warm_start(false))
//   warm_mode 1: warm-sequential inside each thread's chunk (first QP cold),
//                i.e. fcc_qp_test.py:86-89 when nthreads == 1.
// lb/ub/mu strides: 0 => shared by every QP, else elements per QP.
// Returns the wall time (seconds) of the slowest thread's Solve+GetSolution loop.
double fccqp_ref_solve_batch(int B, int n, int m, int nc, int lcs, int max_iter, double rho,
                             double eps_fcone, double eps_bound, int warm_mode, int nthreads,
                             const double* Qcm, const double* b, const double* Acm,
                             const double* beq, const double* mu, long mu_stride,
                             const double* lb, const double* ub, long bound_stride,
                             double* z, int* n_iter, int* status, double* details6) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = B > 0 ? B : 1;
  std::vector<double> elapsed(nthreads, 0.0);
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  auto work = [&](int t) {
    long lo = (long)B * t / nthreads, hi = (long)B * (t + 1) / nthreads;
    RefHandle h(n, m, nc, lcs);
    fccqp_ref_set_options(&h, max_iter, rho, eps_fcone, eps_bound);
    ready.fetch_add(1);
    while (!go.load()) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    for (long i = lo; i < hi; ++i) {
      h.solver.set_warm_start(warm_mode == 1 && i > lo);
      solve_one(&h, Qcm + i * n * n, b + i * n, Acm + i * (long)m * n, beq + i * m,
                mu + i * mu_stride, nc / 3, lb + i * bound_stride, ub + i * bound_stride);
      fetch(&h, z + i * n, n_iter + i, status + i, details6 + i * 6);
    }
    elapsed[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
  while (ready.load() < nthreads) std::this_thread::yield();
  go.store(true);
  for (auto& x : th) x.join();
  double mx = 0;
  for (double e : elapsed) mx = e > mx ? e : mx;
  return mx;
}

// One solve on each of B persistent handles (lane-wise warm start, config 5).
void fccqp_ref_solve_lanes(void** handles, int B, int warm, const double* Qcm, const double* b,
                           const double* Acm, const double* beq, const double* mu,
                           long mu_stride, const double* lb, const double* ub, long bound_stride,
                           double* z, int* n_iter, int* status, double* details6) {
  for (long i = 0; i < B; ++i) {
    RefHandle* h = static_cast<RefHandle*>(handles[i]);
    const int n = h->d.n, m = h->d.m;
    h->solver.set_warm_start(warm != 0);
    solve_one(h, Qcm + i * n * n, b + i * n, Acm + i * (long)m * n, beq + i * m,
              mu + i * mu_stride, h->d.nc / 3, lb + i * bound_stride, ub + i * bound_stride);
    fetch(h, z + i * n, n_iter + i, status + i, details6 + i * 6);
  }
}

int fccqp_ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
}
