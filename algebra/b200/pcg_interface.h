/*
 * algebra/b200/pcg_interface.h -- LinSysSolver vtable instance for the B200 PCG solver.
 * Struct-prefix layout must match `struct linsys_solver`
 * (/root/reference/include/private/types.h:243-279); reference counterpart:
 * algebra/cuda/lin_sys/indirect/cuda_pcg_interface.h:31-141.
 */
#ifndef B200_PCG_INTERFACE_H
#define B200_PCG_INTERFACE_H

#include "osqp.h"
#include "types.h"
#include "algebra_impl.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200pcg_solver_ {
  /* ---- must mirror struct linsys_solver, in order ---- */
  enum osqp_linsys_solver_type type;
  const char* (*name)(struct b200pcg_solver_* self);
  OSQPInt (*solve)(struct b200pcg_solver_* self, OSQPVectorf* b, OSQPInt admm_iter);
  void (*update_settings)(struct b200pcg_solver_* self, const OSQPSettings* settings);
  void (*warm_start)(struct b200pcg_solver_* self, const OSQPVectorf* x);
  OSQPInt (*adjoint_derivative)(struct b200pcg_solver_* self);
  void (*free)(struct b200pcg_solver_* self);
  OSQPInt (*update_matrices)(struct b200pcg_solver_* self, const OSQPMatrix* P,
                             const OSQPInt* Px_new_idx, OSQPInt P_new_n, const OSQPMatrix* A,
                             const OSQPInt* Ax_new_idx, OSQPInt A_new_n);
  OSQPInt (*update_rho_vec)(struct b200pcg_solver_* self, const OSQPVectorf* rho_vec,
                            OSQPFloat rho_sc);
  OSQPInt nthreads;

  /* ---- private state ---- */
  b200_pcg* pcg;                 /* device-side solver (kernels, work vectors, fused operator) */
  OSQPInt   n, m;
  OSQPInt   polishing;
  OSQPInt   max_iter;            /* settings->cg_max_iter      */
  OSQPInt   reduction_threshold; /* settings->cg_tol_reduction */
  OSQPFloat tol_fraction;        /* settings->cg_tol_fraction  */
  osqp_precond_type precond_type;
  OSQPFloat sigma, rho;
  const OSQPFloat* d_rho_vec;    /* borrowed from the core's rho_vec, NULL for scalar rho */
  OSQPFloat* scaled_prim_res;    /* host addresses inside OSQPWorkspace (types.h:199-200) */
  OSQPFloat* scaled_dual_res;
} b200pcg_solver;

OSQPInt init_linsys_solver_b200pcg(b200pcg_solver** sp, const OSQPMatrix* P, const OSQPMatrix* A,
                                   const OSQPVectorf* rho_vec, const OSQPSettings* settings,
                                   OSQPFloat* scaled_prim_res, OSQPFloat* scaled_dual_res,
                                   OSQPInt polishing);

/* total CG iterations / number of solves since init (synchronises) -- the metric
 * "mean PCG iterations per ADMM iteration" is not exposed by the reference */
void b200pcg_get_stats(const LinSysSolver* s, long long* total_iters, long long* n_solves);

#ifdef __cplusplus
}
#endif

#endif
