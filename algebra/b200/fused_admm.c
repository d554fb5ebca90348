/*
 * algebra/b200/fused_admm.c -- fused ADMM steps for the product build.
 *
 * The private algebra interface is per-op, so the unmodified core spends 5 kernels on
 * compute_rhs (src/auxil.c:136-158) and 16 on update_x / update_z / update_y (auxil.c:172-229)
 * where two suffice.  auxil.c stays BYTE-IDENTICAL: the Makefile compiles it with
 *     -Dupdate_xz_tilde=osqp_ref_update_xz_tilde -Dupdate_x=osqp_ref_update_x
 *     -Dupdate_z=osqp_ref_update_z -Dupdate_y=osqp_ref_update_y -Dupdate_info=osqp_ref_update_info
 *     -Dcheck_termination=osqp_ref_check_termination   (the macro also renames the OSQPSettings field of
 *      that name inside this one translation unit -- consistently, so the layout is untouched)
 *     -Dstore_solution=osqp_ref_store_solution
 * so that the reference definitions keep existing under the osqp_ref_ names (and can be selected
 * at run time with OSQP_B200_UNFUSED=1 for A/B parity runs), while the calls made by
 * src/osqp_api.c:708-726 bind to the versions below.  Expression order follows the reference's
 * add_scaled / add_scaled3 / ew_prod calls term by term.
 */
#include "osqp.h"
#include "auxil.h"
#include "types.h"
#include "algebra_impl.h"
#include "timing.h"
#include "glob_opts.h"
#include "scaling.h"

#include <stdlib.h>

void osqp_ref_update_xz_tilde(OSQPSolver* solver, OSQPInt admm_iter);
void osqp_ref_update_x(OSQPSolver* solver);
void osqp_ref_update_z(OSQPSolver* solver);
void osqp_ref_update_y(OSQPSolver* solver);
void osqp_ref_update_info(OSQPSolver* solver, OSQPInt iter, OSQPInt polishing);
OSQPInt osqp_ref_check_termination(OSQPSolver* solver, OSQPInt approximate);
void osqp_ref_store_solution(OSQPSolver* solver, OSQPSolution* solution);


/* A x carried through the relaxation step: valid from the first exact product of a
 * solve (update_info at iteration 1) until the solve ends; recomputed exactly every
 * AX_EXACT_EVERY-th check so that rounding drift cannot accumulate.  One record per host thread
 * (like the library context), keyed by the solver. */
#define AX_EXACT_EVERY 10
static _Thread_local struct {
  const OSQPSolver* solver;
  int               valid;
  int               checks;
} ax_carry;

static int no_ax_carry(void) {
  static int v = -1;
  if (v < 0) v = getenv("B200_NO_AX_CARRY") ? 1 : 0;
  return v;
}

static int unfused(void) {
  static int v = -1;
  if (v < 0) v = getenv("OSQP_B200_UNFUSED") ? 1 : 0;
  return v;
}

/* compute_rhs + LinSysSolver.solve (auxil.c:136-170) */
void update_xz_tilde(OSQPSolver* solver, OSQPInt admm_iter) {
  OSQPWorkspace* work     = solver->work;
  OSQPSettings*  settings = solver->settings;

  if (unfused()) {
    osqp_ref_update_xz_tilde(solver, admm_iter);
    return;
  }
  if (admm_iter == 1 || ax_carry.solver != solver) {   /* new solve: x, A, scaling may all have changed */
    ax_carry.solver = solver;
    ax_carry.valid  = 0;
    ax_carry.checks = 0;
  }
  b200_admm_compute_rhs(work->xtilde_view->d_val, work->ztilde_view->d_val, work->x_prev->d_val,
                        work->data->q->d_val, work->z_prev->d_val, work->y->d_val,
                        settings->rho_is_vec ? work->rho_inv_vec->d_val : OSQP_NULL, work->rho_inv,
                        settings->sigma, (int)work->data->n, (int)work->data->m);
  work->linsys_solver->solve(work->linsys_solver, work->xz_tilde, admm_iter);
}

/* update_x, update_z and update_y (auxil.c:172-229) in one kernel; osqp_solve calls the three
 * back to back (osqp_api.c:714-722), so the z and y halves are done here and the other two
 * entry points have nothing left to do. */
void update_x(OSQPSolver* solver) {
  OSQPWorkspace* work     = solver->work;
  OSQPSettings*  settings = solver->settings;
  int            vec      = settings->rho_is_vec ? 1 : 0;

  if (unfused()) {
    osqp_ref_update_x(solver);
    return;
  }
  b200_range_push("admm update");      /* OSQP_PROFILER_SEC_ADMM_UPDATE */
  b200_admm_update_xzy_carry(work->x->d_val, work->delta_x->d_val, work->z->d_val, work->y->d_val,
                             work->delta_y->d_val, work->xtilde_view->d_val, work->ztilde_view->d_val,
                             work->x_prev->d_val, work->z_prev->d_val, work->data->l->d_val,
                             work->data->u->d_val, vec ? work->rho_vec->d_val : OSQP_NULL,
                             vec ? work->rho_inv_vec->d_val : OSQP_NULL, settings->rho, work->rho_inv,
                             settings->alpha, (int)work->data->n, (int)work->data->m,
                             (ax_carry.valid && ax_carry.solver == solver && work->data->m > 0)
                                 ? work->Ax->d_val : OSQP_NULL);
  b200_range_pop();
}

void update_z(OSQPSolver* solver) {
  if (unfused()) osqp_ref_update_z(solver);
}

void update_y(OSQPSolver* solver) {
  if (unfused()) osqp_ref_update_y(solver);
}

/* update_info for the ADMM iterates (auxil.c:676-762 with compute_prim_res :308, compute_dual_res
 * :371 and compute_obj_val_dual_gap :231 inlined).  Side effects kept: work->Ax / Px / Aty are
 * materialised (read by later core code); the scratch uses of x_prev / z_prev are dropped (dead
 * until the swap at the top of the next iteration). */
void update_info(OSQPSolver* solver, OSQPInt iter, OSQPInt polishing) {
  OSQPInfo*      info     = solver->info;
  OSQPSettings*  settings = solver->settings;
  OSQPWorkspace* work     = solver->work;
  OSQPInt        n = work->data->n, m = work->data->m;
  int            unscale  = settings->scaling && !settings->scaled_termination;
  double         r[B200_RES_COUNT];
  OSQPFloat      quad_term, lin_term, sup_term, prim_res, dual_res;

  if (polishing || unfused()) {
    osqp_ref_update_info(solver, iter, polishing);
    return;
  }
  info->iter = iter;
  b200_range_push("termination check");

  if (m) {
    /* A x: carried by the fused x/z/y update since the last exact product of this solve */
    if (!(ax_carry.valid && ax_carry.solver == solver) || ++ax_carry.checks >= AX_EXACT_EVERY ||
        no_ax_carry()) {
      OSQPMatrix_Axpy(work->data->A, work->x, work->Ax, 1.0, 0.0);
      ax_carry.checks = 0;
      ax_carry.valid  = (ax_carry.solver == solver) && !no_ax_carry();
    }
  }
  OSQPMatrix_Axpy(work->data->P, work->x, work->Px, 1.0, 0.0);
  if (m) OSQPMatrix_Atxpy(work->data->A, work->y, work->Aty, 1.0, 0.0);

  b200_admm_residuals(work->x->d_val, work->y->d_val, work->z->d_val, work->Ax->d_val,
                      work->Px->d_val, work->Aty->d_val, work->data->q->d_val, work->data->l->d_val,
                      work->data->u->d_val, settings->scaling ? work->scaling->Einv->d_val : OSQP_NULL,
                      settings->scaling ? work->scaling->Dinv->d_val : OSQP_NULL,
                      OSQP_INFTY * OSQP_MIN_SCALING, OSQP_ZERO_DEADZONE, (int)n, (int)m, r);
  b200_range_pop();

  /* primal residual (compute_prim_res) */
  if (m == 0) {
    prim_res = 0.;
  } else {
    work->scaled_prim_res = (OSQPFloat)r[B200_RES_PRIM_S];
    prim_res = unscale ? (OSQPFloat)r[B200_RES_PRIM_U] : work->scaled_prim_res;
  }
  /* dual residual (compute_dual_res) */
  work->scaled_dual_res = (OSQPFloat)r[B200_RES_DUAL_S];
  dual_res = unscale ? work->scaling->cinv * (OSQPFloat)r[B200_RES_DUAL_U] : work->scaled_dual_res;
  info->prim_res = prim_res;
  info->dual_res = dual_res;

  /* objective and duality gap (compute_obj_val_dual_gap) */
  quad_term = (OSQPFloat)r[B200_RES_XPX];
  lin_term  = (OSQPFloat)r[B200_RES_QX];
  sup_term  = (OSQPFloat)r[B200_RES_SC];
  info->obj_val         = 0.5 * quad_term + lin_term;
  info->dual_obj_val    = -0.5 * quad_term - sup_term;
  work->scaled_dual_gap = quad_term + lin_term + sup_term;
  if (settings->scaling) {
    info->obj_val      *= work->scaling->cinv;
    info->dual_obj_val *= work->scaling->cinv;
    info->duality_gap   = work->scaling->cinv * work->scaled_dual_gap;
  } else {
    info->duality_gap = work->scaled_dual_gap;
  }
  work->xtPx = quad_term;
  work->qtx  = lin_term;
  work->SC   = sup_term;

  info->primdual_int += c_absval(info->duality_gap);
#ifdef OSQP_ENABLE_PROFILING
  info->solve_time = osqp_toc(work->timer);
#endif
  info->rel_kkt_error = c_max(c_max(info->dual_res, info->prim_res), info->duality_gap);
#ifdef OSQP_ENABLE_PRINTING
  work->summary_printed = 0;
#endif

  /* norms the core asks for next, valid until the next kernel or copy */
  b200_norm_cache_reset();
  if (m) {
    b200_norm_cache_put(OSQP_NULL, work->z->d_val, work->z->length, (OSQPFloat)r[B200_RES_Z_S]);
    b200_norm_cache_put(OSQP_NULL, work->Ax->d_val, work->Ax->length, (OSQPFloat)r[B200_RES_AX_S]);
    b200_norm_cache_put(OSQP_NULL, work->Aty->d_val, work->Aty->length, (OSQPFloat)r[B200_RES_ATY_S]);
  }
  b200_norm_cache_put(OSQP_NULL, work->data->q->d_val, work->data->q->length, (OSQPFloat)r[B200_RES_Q_S]);
  b200_norm_cache_put(OSQP_NULL, work->Px->d_val, work->Px->length, (OSQPFloat)r[B200_RES_PX_S]);
  if (settings->scaling) {
    if (m) {
      b200_norm_cache_put(work->scaling->Einv->d_val, work->z->d_val, work->z->length, (OSQPFloat)r[B200_RES_Z_U]);
      b200_norm_cache_put(work->scaling->Einv->d_val, work->Ax->d_val, work->Ax->length, (OSQPFloat)r[B200_RES_AX_U]);
      b200_norm_cache_put(work->scaling->Dinv->d_val, work->Aty->d_val, work->Aty->length, (OSQPFloat)r[B200_RES_ATY_U]);
    }
    b200_norm_cache_put(work->scaling->Dinv->d_val, work->data->q->d_val, work->data->q->length, (OSQPFloat)r[B200_RES_Q_U]);
    b200_norm_cache_put(work->scaling->Dinv->d_val, work->Px->d_val, work->Px->length, (OSQPFloat)r[B200_RES_PX_U]);
  }
  b200_norm_cache_seal();
}


/* check_termination (auxil.c:808-945), same decisions, fewer host round trips.  The tolerances are
 * evaluated with the very calls of compute_prim_tol / compute_dual_tol (auxil.c:334-458), which the
 * norm cache filled by update_info above answers without a kernel.  The infeasibility tests
 * (auxil.c:460-585) each start with two or three blocking reductions whose values alone decide
 * whether the test can return non-zero at all: ||delta_y|| > tol and u'dy+ + l'dy- < 0, resp.
 * ||delta_x|| > tol and q'dx < 0.  Those five scalars come from ONE kernel
 * (b200_admm_infeas_scalars); only when they leave the outcome open do the matrix products of the
 * tests follow, with the same calls as the reference minus the reductions already in hand. */
OSQPInt check_termination(OSQPSolver* solver, OSQPInt approximate) {
  OSQPInfo*      info     = solver->info;
  OSQPSettings*  settings = solver->settings;
  OSQPWorkspace* work     = solver->work;
  OSQPFloat eps_abs = settings->eps_abs, eps_rel = settings->eps_rel;
  OSQPFloat eps_prim_inf = settings->eps_prim_inf, eps_dual_inf = settings->eps_dual_inf;
  OSQPFloat eps_prim, eps_dual, eps_gap, mx, tmp;
  OSQPInt   exitflag = 0, prim_ok = 0, dual_ok = 0, gap_ok = 0, prim_inf = 0, dual_inf = 0;
  int       unscale = settings->scaling && !settings->scaled_termination;
  int       need_p = 0, need_d = 0;
  double    f[5];

  if (unfused()) return osqp_ref_check_termination(solver, approximate);

  if ((info->prim_res > OSQP_INFTY) || (info->dual_res > OSQP_INFTY)) {
    update_status(info, OSQP_NON_CVX);
    info->obj_val = OSQP_NAN;
    return 1;
  }
  if (approximate) {
    eps_abs *= 10; eps_rel *= 10; eps_prim_inf *= 10; eps_dual_inf *= 10;
  }

  /* residual tests (compute_prim_tol / compute_dual_tol) */
  if (work->data->m == 0) {
    prim_ok = 1;
  } else {
    if (unscale) {
      mx  = OSQPVectorf_scaled_norm_inf(work->scaling->Einv, work->z);
      tmp = OSQPVectorf_scaled_norm_inf(work->scaling->Einv, work->Ax);
    } else {
      mx  = OSQPVectorf_norm_inf(work->z);
      tmp = OSQPVectorf_norm_inf(work->Ax);
    }
    eps_prim = eps_abs + eps_rel * c_max(mx, tmp);
    if (info->prim_res < eps_prim) prim_ok = 1;
    else need_p = 1;
  }
  if (unscale) {
    mx  = OSQPVectorf_scaled_norm_inf(work->scaling->Dinv, work->data->q);
    tmp = OSQPVectorf_scaled_norm_inf(work->scaling->Dinv, work->Aty);
    mx  = c_max(mx, tmp);
    tmp = OSQPVectorf_scaled_norm_inf(work->scaling->Dinv, work->Px);
    mx  = c_max(mx, tmp) * work->scaling->cinv;
  } else {
    mx  = OSQPVectorf_norm_inf(work->data->q);
    tmp = OSQPVectorf_norm_inf(work->Aty);
    mx  = c_max(mx, tmp);
    tmp = OSQPVectorf_norm_inf(work->Px);
    mx  = c_max(mx, tmp);
  }
  eps_dual = eps_abs + eps_rel * mx;
  if (info->dual_res < eps_dual) dual_ok = 1;
  else need_d = 1;

  /* infeasibility tests: the reductions they start with from one kernel, the matrix products only when
     those leave the outcome open */
  if (need_p || need_d) {
    /* writes delta_y only: the norms parked by update_info stay valid for compute_rho_estimate */
    int live = b200_norm_cache_live();
    b200_admm_infeas_scalars(work->delta_y->d_val, work->data->l->d_val, work->data->u->d_val,
                             unscale ? work->scaling->E->d_val : OSQP_NULL, work->delta_x->d_val,
                             unscale ? work->scaling->D->d_val : OSQP_NULL, work->data->q->d_val,
                             OSQP_INFTY * OSQP_MIN_SCALING, (int)work->data->n, (int)work->data->m,
                             need_p, need_d, f);
    b200_norm_cache_after(live, work->delta_y->d_val, work->delta_y->length);
    /* is_primal_infeasible (auxil.c:460-514) from here on: ||A' dy|| < eps ||dy||, with the unscaling
       folded into the norm (max |Dinv_i v_i| is what ew_prod + norm_inf compute, product by product) */
    if (need_p && f[0] > OSQP_DIVISION_TOL && (f[1] + f[2]) < 0.0) {
      OSQPMatrix_Atxpy(work->data->A, work->delta_y, work->Atdelta_y, 1.0, 0.0);
      tmp = unscale ? OSQPVectorf_scaled_norm_inf(work->scaling->Dinv, work->Atdelta_y)
                    : OSQPVectorf_norm_inf(work->Atdelta_y);
      prim_inf = tmp < eps_prim_inf * (OSQPFloat)f[0];
    }
    /* is_dual_infeasible (auxil.c:516-585): ||P dx|| < c eps ||dx||, then A dx in the recession cone */
    if (need_d && f[3] > OSQP_DIVISION_TOL && f[4] < 0.0) {
      OSQPFloat cost_scaling = unscale ? work->scaling->c : 1.0;
      OSQPMatrix_Axpy(work->data->P, work->delta_x, work->Pdelta_x, 1.0, 0.0);
      tmp = unscale ? OSQPVectorf_scaled_norm_inf(work->scaling->Dinv, work->Pdelta_x)
                    : OSQPVectorf_norm_inf(work->Pdelta_x);
      if (tmp < cost_scaling * eps_dual_inf * (OSQPFloat)f[3]) {
        OSQPMatrix_Axpy(work->data->A, work->delta_x, work->Adelta_x, 1.0, 0.0);
        if (unscale) OSQPVectorf_ew_prod(work->Adelta_x, work->Adelta_x, work->scaling->Einv);
        dual_inf = OSQPVectorf_in_reccone(work->Adelta_x, work->data->l, work->data->u,
                                          OSQP_INFTY * OSQP_MIN_SCALING, eps_dual_inf * (OSQPFloat)f[3]);
      }
    }
  }

  /* duality gap (compute_duality_gap_tol) */
  if (settings->check_dualgap) {
    mx = c_absval(work->xtPx);
    mx = c_max(mx, c_absval(work->qtx));
    mx = c_max(mx, c_absval(work->SC));
    if (unscale) mx = work->scaling->cinv * mx;
    eps_gap = eps_abs + eps_rel * mx;
    if (unscale) { if (c_absval(info->duality_gap) < eps_gap) gap_ok = 1; }
    else         { if (c_absval(work->scaled_dual_gap) < eps_gap) gap_ok = 1; }
  } else {
    gap_ok = 1;
  }

  if (prim_ok && dual_ok && gap_ok) {
    update_status(info, approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED);
    exitflag = 1;
  } else if (prim_inf) {
    update_status(info, approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE);
    if (unscale) OSQPVectorf_ew_prod(work->delta_y, work->delta_y, work->scaling->E);
    info->obj_val = OSQP_INFTY;
    exitflag      = 1;
  } else if (dual_inf) {
    update_status(info, approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE);
    if (unscale) OSQPVectorf_ew_prod(work->delta_x, work->delta_x, work->scaling->D);
    info->obj_val = -OSQP_INFTY;
    exitflag      = 1;
  }
  return exitflag;
}


/* store_solution (auxil.c:598-675).  For a run that ended with a solution the reference fills two device
 * vectors with NaN and downloads them into the certificate arrays: (n + m) values over PCIe per solve --
 * 240 MB on BASELINE configs[3], 20-40 ms of a 330 ms solve.  The host arrays are filled with NaN here
 * directly, and not at all when they still hold the NaN of the previous call (first and last entry are
 * looked at: the arrays only ever hold all-OSQP_NAN, a certificate normalised to |v| <= 1, or the zeros of a fresh
 * allocation).  The device vectors are still set to NaN, as the reference leaves them.  Every other case
 * (no solution: NaN iterates, certificates; a caller-provided OSQPSolution) goes through the reference
 * function. */
static int all_nan_already(const OSQPFloat* a, OSQPInt len) {
  /* OSQP_NAN is the finite marker value (OSQPFloat)0x7fc00000; a normalised certificate never holds it */
  return len == 0 || (a[0] == OSQP_NAN && a[len - 1] == OSQP_NAN);
}

void store_solution(OSQPSolver* solver, OSQPSolution* solution) {
  OSQPWorkspace* work = solver->work;
  OSQPInt i, n, m;
  if (!solution) return;
  /* only the solver's own (host, calloc'ed by osqp_setup) solution arrays are written from the host; the
     arrays osqp_get_solution is handed may live in device memory (tests/basic_qp/test_cuda_io.cpp) */
  if (unfused() || solution != solver->solution || !has_solution(solver->info)) {
    osqp_ref_store_solution(solver, solution);
    return;
  }
  if (solver->settings->scaling) {
    unscale_solution(work->x_prev, work->z_prev, work->x, work->y, work);
    OSQPVectorf_to_raw(solution->x, work->x_prev);
    OSQPVectorf_to_raw(solution->y, work->z_prev);
  } else {
    OSQPVectorf_to_raw(solution->x, work->x);
    OSQPVectorf_to_raw(solution->y, work->y);
  }
  OSQPVectorf_set_scalar(work->delta_y, OSQP_NAN);
  OSQPVectorf_set_scalar(work->delta_x, OSQP_NAN);
  n = work->delta_x->length;
  m = work->delta_y->length;
  if (!all_nan_already(solution->prim_inf_cert, m))
    for (i = 0; i < m; i++) solution->prim_inf_cert[i] = OSQP_NAN;
  if (!all_nan_already(solution->dual_inf_cert, n))
    for (i = 0; i < n; i++) solution->dual_inf_cert[i] = OSQP_NAN;
}
