/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference solver.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; it is the checker, never the product.
 * Parity pinned: yes -- tests/test_oracle.py checks this restatement against
 * (a) golden vectors produced by the UNMODIFIED reference compiled here
 * (oracle/_ref, generator tests/golden/make_goldens.py) and (b) the live
 * oracle/_ref library whenever it is present.
 *
 * All matrices are COLUMN-major, like the reference's Eigen types.
 */
#ifndef FCCQP_ORACLE_H
#define FCCQP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fccqp_oracle fccqp_oracle;

fccqp_oracle* fccqp_oracle_create(int n, int m, int nc, int lcs);
void fccqp_oracle_destroy(fccqp_oracle* o);
void fccqp_oracle_set_options(fccqp_oracle* o, int max_iter, double rho, double eps_fcone,
                              double eps_bound);
void fccqp_oracle_set_warm_start(fccqp_oracle* o, int warm);
/* returns 0, or -1 when nmu < nc/3 (the reference throws std::out_of_range). */
int fccqp_oracle_solve(fccqp_oracle* o, const double* Qcm, const double* b, const double* Acm,
                       const double* beq, const double* mu, int nmu, const double* lb,
                       const double* ub);
/* details6 = {admm_residual_bounds, admm_residual_friction_cone, bounds_viol,
 *             friction_cone_viol, solve_time, factorization_time} */
void fccqp_oracle_get_solution(const fccqp_oracle* o, double* z, int* n_iter, int* status,
                               double* details6);
/* which pre-solve branch the last Solve took: 0 none (warm), 1 LDLT, 2 COD */
int fccqp_oracle_presolve_path(const fccqp_oracle* o);
/* warm-start state (x_, mu_x_, mu_lambda_c_) in and out */
void fccqp_oracle_get_state(const fccqp_oracle* o, double* x, double* mu_x, double* mu_c);
void fccqp_oracle_set_state(fccqp_oracle* o, const double* x, const double* mu_x,
                            const double* mu_c);

double fccqp_oracle_solve_batch(int B, int n, int m, int nc, int lcs, int max_iter, double rho,
                                double eps_fcone, double eps_bound, int warm_mode, int nthreads,
                                const double* Qcm, const double* b, const double* Acm,
                                const double* beq, const double* mu, long mu_stride,
                                const double* lb, const double* ub, long bound_stride, double* z,
                                int* n_iter, int* status, double* details6);
void fccqp_oracle_solve_lanes(void** handles, int B, int warm, const double* Qcm, const double* b,
                              const double* Acm, const double* beq, const double* mu,
                              long mu_stride, const double* lb, const double* ub,
                              long bound_stride, double* z, int* n_iter, int* status,
                              double* details6);
int fccqp_oracle_hardware_threads(void);

/* constraint_utils.cpp restatements, exported for known-answer tests */
void fccqp_oracle_project_cone3(const double f[3], double mu, double out[3]);
double fccqp_oracle_cone_violation(const double* f, int nc, const double* mu);
double fccqp_oracle_bound_violation(const double* x, const double* lb, const double* ub, int n);

/* Test hook: per-iteration residuals (bound, friction cone) of the following solves into buf[2 * cap]; the
 * parity tests use it to show that an iteration-count difference sits on the exit threshold. */
void fccqp_oracle_set_trace(fccqp_oracle* o, double* buf, int cap);

/* Extension used only to check the product's opt-in over-relaxation (include/fccqp.h); 1.0 = the reference. */
void fccqp_oracle_set_relaxation(double alpha);
/* Same for the product's opt-in adaptive rho (fccqp_options::adapt_rho_interval); 0 = off = the reference. */
void fccqp_oracle_set_adaptive_rho(int interval);

#ifdef __cplusplus
}
#endif
#endif
