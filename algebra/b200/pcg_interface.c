/*
 * algebra/b200/pcg_interface.c -- LinSysSolver vtable (init / solve / update_matrices /
 * update_rho_vec / update_settings / warm_start / free) for the B200 reduced-KKT PCG.
 *
 * Contract reproduced from algebra/cuda/lin_sys/indirect/cuda_pcg_interface.cu:99-361
 * (SURVEY.md appendix B.8-B.9): solve() is in place on (x~, z~) and returns z~ = A x~; with
 * polishing it uses sigma = delta, rho = 1/delta and returns (A x - b2)/delta; the iterate is
 * always warm-started from the previous solution; update_rho_vec may get rho_vec == NULL.
 * The whole solve -- including the tolerance schedule of compute_tolerance (:32-64) -- runs
 * inside one persistent kernel (osqp_b200/csrc/pcg.cu); this file only forwards host scalars.
 */
#include "pcg_interface.h"
#include "glob_opts.h"
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec;
}

static const char* name_b200pcg(b200pcg_solver* s) {
  switch (s->precond_type) {
  case OSQP_NO_PRECONDITIONER:
    return "B200 persistent-kernel Conjugate Gradient - No preconditioner";
  case OSQP_DIAGONAL_PRECONDITIONER:
    return "B200 persistent-kernel Conjugate Gradient - Diagonal preconditioner";
  }
  return "B200 persistent-kernel Conjugate Gradient - Unknown preconditioner";
}

static void configure(b200pcg_solver* s) {
  b200_pcg_configure(s->pcg, s->sigma, s->rho, s->d_rho_vec,
                     s->precond_type == OSQP_DIAGONAL_PRECONDITIONER ? 1 : 0, (int)s->polishing);
}

static OSQPInt solve_linsys_b200pcg(b200pcg_solver* s, OSQPVectorf* b, OSQPInt admm_iter) {
  double pr = s->scaled_prim_res ? (double)*s->scaled_prim_res : 0.0;
  double dr = s->scaled_dual_res ? (double)*s->scaled_dual_res : 0.0;
  OSQPInt rc;
  b200_range_push("linsys solve");     /* OSQP_PROFILER_SEC_LINSYS_SOLVE */
  rc = b200_pcg_solve(s->pcg, b->d_val, (int)admm_iter, pr, dr, (int)s->max_iter,
                      (double)s->tol_fraction, (int)s->reduction_threshold);
  b200_range_pop();
  return rc;
}

static void update_settings_b200pcg(b200pcg_solver* s, const OSQPSettings* settings) {
  s->max_iter            = settings->cg_max_iter;
  s->reduction_threshold = settings->cg_tol_reduction;
  s->tol_fraction        = settings->cg_tol_fraction;
  if (s->precond_type != settings->cg_precond) {
    s->precond_type = settings->cg_precond;
    configure(s);
    b200_pcg_refresh_precond(s->pcg);
  }
}

static void warm_start_b200pcg(b200pcg_solver* s, const OSQPVectorf* x) {
  b200_pcg_warm_start(s->pcg, x->d_val);
}

static void free_b200pcg(b200pcg_solver* s) {
  if (s) {
    b200_pcg_destroy(s->pcg);
    c_free(s);
  }
}

static OSQPInt update_matrices_b200pcg(b200pcg_solver* s, const OSQPMatrix* P,
                                       const OSQPInt* Px_new_idx, OSQPInt P_new_n,
                                       const OSQPMatrix* A, const OSQPInt* Ax_new_idx,
                                       OSQPInt A_new_n) {
  /* the device solver borrows P / A / A': their value arrays are already updated, so only
     the fused operator and the Jacobi diagonal need rebuilding (cuda_pcg_interface.cu:336-349) */
  (void)P; (void)Px_new_idx; (void)P_new_n; (void)A; (void)Ax_new_idx; (void)A_new_n;
  b200_pcg_refresh_matrices(s->pcg);
  b200_pcg_refresh_precond(s->pcg);
  return 0;
}

static OSQPInt update_rho_vec_b200pcg(b200pcg_solver* s, const OSQPVectorf* rho_vec,
                                      OSQPFloat rho_sc) {
  /* rho_vec (if any) is the same device array we already hold */
  (void)rho_vec;
  s->rho = rho_sc;
  configure(s);
  b200_pcg_refresh_precond(s->pcg);
  return 0;
}

OSQPInt init_linsys_solver_b200pcg(b200pcg_solver** sp, const OSQPMatrix* P, const OSQPMatrix* A,
                                   const OSQPVectorf* rho_vec, const OSQPSettings* settings,
                                   OSQPFloat* scaled_prim_res, OSQPFloat* scaled_dual_res,
                                   OSQPInt polishing) {
  b200pcg_solver* s = (b200pcg_solver*)c_calloc(1, sizeof(b200pcg_solver));
  *sp = s;
  if (!s) return OSQP_MEM_ALLOC_ERROR;

  s->type     = OSQP_INDIRECT_SOLVER;
  s->nthreads = 1;
  s->n        = P->n;
  s->m        = A->m;

  s->polishing           = polishing;
  s->max_iter            = settings->cg_max_iter;
  s->precond_type        = settings->cg_precond;
  s->reduction_threshold = settings->cg_tol_reduction;
  s->tol_fraction        = settings->cg_tol_fraction;
  s->scaled_prim_res     = scaled_prim_res;
  s->scaled_dual_res     = scaled_dual_res;
  s->d_rho_vec           = rho_vec ? rho_vec->d_val : OSQP_NULL;

  if (polishing) {
    s->sigma = settings->delta;
    s->rho   = 1. / settings->delta;
  } else {
    s->sigma = settings->sigma;
    s->rho   = settings->rho;
  }

  s->name               = &name_b200pcg;
  s->solve              = &solve_linsys_b200pcg;
  s->warm_start         = &warm_start_b200pcg;
  s->adjoint_derivative = OSQP_NULL;
  s->free               = &free_b200pcg;
  s->update_matrices    = &update_matrices_b200pcg;
  s->update_rho_vec     = &update_rho_vec_b200pcg;
  s->update_settings    = &update_settings_b200pcg;

  double t0 = now_ms();
  b200_range_push("linsys init");      /* OSQP_PROFILER_SEC_LINSYS_INIT */
  s->pcg = b200_pcg_create(P->S, A->S, A->St, (int)s->n, (int)s->m);
  if (!s->pcg) { b200_range_pop(); return OSQP_MEM_ALLOC_ERROR; }

  configure(s);
  b200_pcg_refresh_matrices(s->pcg);
  b200_pcg_refresh_precond(s->pcg);
  b200_range_pop();
  if (getenv("B200_TRACE_SETUP")) { b200_sync(); fprintf(stderr, "[b200 trace] linsys init %.1f ms\n", now_ms() - t0); }
  return 0;
}

void b200pcg_get_stats(const LinSysSolver* ls, long long* total_iters, long long* n_solves) {
  const b200pcg_solver* s = (const b200pcg_solver*)ls;
  b200_pcg_stats(s->pcg, total_iters, n_solves, OSQP_NULL, OSQP_NULL, OSQP_NULL);
}
