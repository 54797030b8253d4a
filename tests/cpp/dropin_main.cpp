// C++ drop-in check: a caller written against the reference's src/fcc_qp.hpp (Eigen types,
// fcc_qp::FCCQP / FCCQPOptions / FCCQPSolution, src/fcc_qp.hpp:73-121) compiled against THIS
// repository's include/fcc_qp.hpp and linked with libfccqp_b200.so.  Exit code 0 = all checks pass.
//   g++ -std=c++17 -I<eigen> -I include tests/cpp/dropin_main.cpp -L fcc_qp_b200 -lfccqp_b200 -Wl,-rpath,...
#include <Eigen/Dense>
#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

#include "fcc_qp.hpp"

using Eigen::MatrixXd;
using Eigen::VectorXd;
using fcc_qp::FCCQP;
using fcc_qp::FCCQPOptions;
using fcc_qp::FCCQPSolution;

static int fails = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++fails; } \
  } while (0)

int main() {
  const double inf = std::numeric_limits<double>::infinity();
  FCCQPOptions options;
  CHECK(options.max_iter == 1000 && options.rho == 1e-6 && options.eps_fcone == 1e-3 && options.eps_bound == 1e-6);
  options.rho = 5e-5; options.eps_fcone = 1e-6; options.eps_bound = 1e-6; options.max_iter = 100;

  {  // equality-only problem: closed-form KKT solution, n_iter = 0
    FCCQP solver(3, 1, 0, 3);
    solver.set_options(options);
    MatrixXd Q = MatrixXd::Identity(3, 3);
    VectorXd b = VectorXd::Zero(3), b_eq = VectorXd::Ones(1);
    MatrixXd A = MatrixXd::Ones(1, 3);
    VectorXd lb = VectorXd::Constant(3, -inf), ub = VectorXd::Constant(3, inf);
    solver.Solve(Q, b, A, b_eq, std::vector<double>{}, lb, ub);
    FCCQPSolution sol = solver.GetSolution();
    CHECK(sol.details.n_iter == 0);
    for (int i = 0; i < 3; ++i) CHECK(std::abs(sol.z(i) - 1.0 / 3.0) < 1e-12);
  }
  {  // projection onto one friction cone: min 1/2 |x - f|^2, x in F(mu = 0.5), f = (3, 4, 1) -> (0.84, 1.12, 2.8)
    FCCQP solver(3, 0, 3, 0);
    solver.set_options(options);
    solver.set_rho(1.0);
    solver.set_max_iter(500);
    solver.set_warm_start(false);
    CHECK(solver.contact_vars_start() == 0);
    MatrixXd Q = MatrixXd::Identity(3, 3);
    VectorXd f(3); f << 3.0, 4.0, 1.0;
    VectorXd b = -f;
    MatrixXd A(0, 3); VectorXd b_eq(0);
    VectorXd lb = VectorXd::Constant(3, -inf), ub = VectorXd::Constant(3, inf);
    solver.Solve(Q, b, A, b_eq, std::vector<double>{0.5}, lb, ub);
    FCCQPSolution sol = solver.GetSolution();
    CHECK(std::abs(sol.z(0) - 0.84) < 1e-4 && std::abs(sol.z(1) - 1.12) < 1e-4 && std::abs(sol.z(2) - 2.8) < 1e-4);
    CHECK(sol.details.n_iter > 0 && sol.details.n_iter < 500);
    CHECK(sol.details.solve_time > 0.0);
    // warm restart from the converged state: immediate exit
    solver.set_warm_start(true);
    solver.Solve(Q, b, A, b_eq, std::vector<double>{0.5}, lb, ub);
    CHECK(solver.GetSolution().details.n_iter <= 1);
    // column-major blocks of a larger matrix (outer stride != rows), as Eigen::Ref allows
    MatrixXd big = MatrixXd::Zero(5, 5);
    big.topLeftCorner(3, 3) = Q;
    solver.set_warm_start(false);
    solver.Solve(big.topLeftCorner(3, 3), b, A, b_eq, std::vector<double>{0.5}, lb, ub);
    CHECK(std::abs(solver.GetSolution().z(2) - 2.8) < 1e-4);
  }
  {  // batched C++ entry point: B cone projections with different targets, against single solves
    const int B = 16;
    std::vector<double> Q(B * 9, 0.0), b(B * 3), mu(B), lbv(3, -inf), ubv(3, inf);
    for (int k = 0; k < B; ++k) {
      for (int i = 0; i < 3; ++i) Q[k * 9 + i * 3 + i] = 1.0;
      b[k * 3 + 0] = -(1.0 + 0.25 * k); b[k * 3 + 1] = -(2.0 - 0.1 * k); b[k * 3 + 2] = -(0.5 + 0.05 * k);
      mu[k] = 0.4 + 0.02 * k;
    }
    fcc_qp::FCCQPBatch batch(3, 0, 3, 0);
    options.rho = 1.0; options.max_iter = 500;
    batch.set_options(options);
    fcc_qp::FCCQPBatchProblem p;
    p.Q = Q.data(); p.b = b.data(); p.friction_coeffs = mu.data(); p.lb = lbv.data(); p.ub = ubv.data();
    p.shared_bounds = true;
    batch.Solve(B, p);
    const fcc_qp::FCCQPBatchSolution& bs = batch.GetSolution();
    FCCQP single(3, 0, 3, 0);
    single.set_options(options);
    for (int k = 0; k < B; ++k) {
      MatrixXd Qk = MatrixXd::Identity(3, 3);
      VectorXd bk(3); bk << b[k * 3], b[k * 3 + 1], b[k * 3 + 2];
      MatrixXd A(0, 3); VectorXd beq(0);
      VectorXd lb = VectorXd::Constant(3, -inf), ub = VectorXd::Constant(3, inf);
      single.Solve(Qk, bk, A, beq, std::vector<double>{mu[k]}, lb, ub);
      FCCQPSolution s1 = single.GetSolution();
      for (int i = 0; i < 3; ++i) CHECK(std::abs(bs.z[k * 3 + i] - s1.z(i)) < 1e-12);
      CHECK(bs.n_iter[k] == s1.details.n_iter);
      CHECK(bs.solve_status[k] == (int)s1.details.solve_status);
    }
  }
  std::printf(fails ? "dropin_main: %d failure(s)\n" : "dropin_main: ok\n", fails);
  return fails ? 1 : 0;
}
