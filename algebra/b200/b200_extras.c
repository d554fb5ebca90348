/*
 * algebra/b200/b200_extras.c -- small exported helpers that need the private workspace
 * layout (include/private/types.h).  Not part of OSQP's interface; used by the Python host
 * mirror, bench.py and the tests for metrics the reference does not expose.
 */
#include "osqp.h"
#include "types.h"
#include "pcg_interface.h"
#include "algebra_impl.h"

/* total CG iterations and number of linear solves of this solver (synchronises) */
OSQPInt osqp_b200_cg_stats(const OSQPSolver* solver, long long* total_iters, long long* n_solves) {
  if (!solver || !solver->work || !solver->work->linsys_solver) return 1;
  b200pcg_get_stats(solver->work->linsys_solver, total_iters, n_solves);
  return 0;
}

/* problem sizes as stored by the backend: n, m, nnz(A), nnz(P triu) */
OSQPInt osqp_b200_dims(const OSQPSolver* solver, OSQPInt* n, OSQPInt* m, OSQPInt* nnzA,
                       OSQPInt* nnzP) {
  if (!solver || !solver->work || !solver->work->data) return 1;
  *n    = solver->work->data->n;
  *m    = solver->work->data->m;
  *nnzA = OSQPMatrix_get_nz(solver->work->data->A);
  *nnzP = OSQPMatrix_get_nz(solver->work->data->P);
  return 0;
}

/* sizes of the public structs this library was compiled with (ctypes layout check) */
OSQPInt osqp_b200_sizeof(OSQPInt which) {
  switch (which) {
  case 0: return (OSQPInt)sizeof(OSQPSettings);
  case 1: return (OSQPInt)sizeof(OSQPInfo);
  case 2: return (OSQPInt)sizeof(OSQPCscMatrix);
  case 3: return (OSQPInt)sizeof(OSQPFloat);
  case 4: return (OSQPInt)sizeof(OSQPInt);
  default: return -1;
  }
}

/* Declare the row-sharded layout before osqp_setup: n = global number of variables, m_local =
 * rows of A held by this rank (must differ from n).  m_local < 0 switches the mode off. */
OSQPInt osqp_b200_dist_configure(OSQPInt n, OSQPInt m_local) {
  if (m_local >= 0 && m_local == n) return 1;
  b200_dist_n       = n;
  b200_dist_mlocal  = m_local;
  b200_dist_nshared = -1;
  b200_dist_nglobal = -1;
  b200_dist_set_split(-1);
  return 0;
}

/* Column-split layout: this rank's QP has n_local = n_shared + (owned columns) variables and m_local
 * rows; n_global is the column count of the whole problem.  n_local must differ from m_local. */
OSQPInt osqp_b200_dist_configure_split(OSQPInt n_local, OSQPInt m_local, OSQPInt n_shared, OSQPInt n_global) {
  if (n_local == m_local || n_shared < 0 || n_shared > n_local) return 1;
  b200_dist_n       = n_local;
  b200_dist_mlocal  = m_local;
  b200_dist_nshared = n_shared;
  b200_dist_nglobal = n_global;
  b200_dist_set_split((int)n_shared);
  return 0;
}

/* Solve nb QPs that share P and A with the set-up `solver` (the template) and differ in their bounds
 * (l_batch, u_batch: nb x m, row-major, user units) and optionally their linear cost (q_batch: nb x n,
 * or NULL = the template's q): one CTA per QP, the whole osqp_solve loop on the device
 * (osqp_b200/csrc/batch.cu).  All pointers are HOST arrays; x_out (nb x n), y_out (nb x m), iters,
 * status, obj, prim_res, dual_res, cg_iters, rho_updates (nb each) receive the results.  The template's
 * current settings (rho, sigma, alpha, eps, check_termination, adaptive rho, CG options) apply to every
 * QP; the template's own iterates are not touched.  Returns 0 on success; a QP whose status is
 * OSQP_MAX_ITER_REACHED should be re-solved through osqp_update_data_vec / osqp_solve (the batch kernel
 * does not evaluate infeasibility certificates). */
static _Thread_local double last_batch_kernel_ms = 0.0;
/* device time (CUDA events on the library stream) of the kernel of the last osqp_b200_solve_batch call */
double osqp_b200_last_batch_kernel_ms(void) { return last_batch_kernel_ms; }

OSQPInt osqp_b200_solve_batch(OSQPSolver* solver, OSQPInt nb, const OSQPFloat* l_batch, const OSQPFloat* u_batch,
                              const OSQPFloat* q_batch, OSQPFloat* x_out, OSQPFloat* y_out, OSQPInt* iters,
                              OSQPInt* status, OSQPFloat* obj, OSQPFloat* prim_res, OSQPFloat* dual_res,
                              OSQPInt* cg_iters, OSQPInt* rho_updates) {
  OSQPWorkspace* work;
  OSQPSettings*  st;
  b200_batch_settings bs;
  OSQPInt n, m, rc = 1;
  size_t fn, fm;
  OSQPFloat *d_l = 0, *d_u = 0, *d_q = 0, *d_x = 0, *d_y = 0, *d_f = 0;
  OSQPInt   *d_i = 0;

  if (!solver || !solver->work || nb <= 0 || !l_batch || !u_batch) return 1;
  work = solver->work;
  st   = solver->settings;
  n    = work->data->n;
  m    = work->data->m;
  if (st->adaptive_rho != OSQP_ADAPTIVE_RHO_UPDATE_DISABLED && st->adaptive_rho != OSQP_ADAPTIVE_RHO_UPDATE_ITERATIONS)
    return 4;   /* time / KKT-error based rho updates are host decisions: not available in the batch kernel */
  fn = (size_t)nb * (size_t)n * sizeof(OSQPFloat);
  fm = (size_t)nb * (size_t)m * sizeof(OSQPFloat);
  d_l = (OSQPFloat*)b200_malloc(fm);  d_u = (OSQPFloat*)b200_malloc(fm);
  d_x = (OSQPFloat*)b200_malloc(fn);  d_y = (OSQPFloat*)b200_malloc(fm);
  d_f = (OSQPFloat*)b200_malloc(3 * (size_t)nb * sizeof(OSQPFloat));
  d_i = (OSQPInt*)b200_malloc(4 * (size_t)nb * sizeof(OSQPInt));
  if (q_batch) d_q = (OSQPFloat*)b200_malloc(fn);
  if (!d_l || !d_u || !d_x || !d_y || !d_f || !d_i || (q_batch && !d_q)) goto done;
  if (b200_copy_in(d_l, l_batch, fm) || b200_copy_in(d_u, u_batch, fm)) goto done;
  if (q_batch && b200_copy_in(d_q, q_batch, fn)) goto done;

  bs.rho = st->rho; bs.sigma = st->sigma; bs.alpha = st->alpha; bs.eps_abs = st->eps_abs; bs.eps_rel = st->eps_rel;
  bs.adaptive_rho_tolerance = st->adaptive_rho_tolerance;
  bs.rho_is_vec = st->rho_is_vec; bs.max_iter = st->max_iter; bs.check_termination = st->check_termination;
  bs.adaptive_rho = st->adaptive_rho; bs.adaptive_rho_interval = st->adaptive_rho_interval;
  bs.check_dualgap = st->check_dualgap; bs.scaled_termination = st->scaled_termination;
  bs.cg_max_iter = st->cg_max_iter; bs.cg_tol_reduction = st->cg_tol_reduction; bs.cg_tol_fraction = st->cg_tol_fraction;

  {
    void* e0 = b200_event_create();
    void* e1 = b200_event_create();
    b200_event_record(e0);
  rc = b200_batch_solve(work->data->P->S, work->data->A->S, work->data->A->St, (int)n, (int)m, (int)nb,
                        work->data->q->d_val, d_q, d_l, d_u,
                        st->scaling ? work->scaling->D->d_val : OSQP_NULL, st->scaling ? work->scaling->Dinv->d_val : OSQP_NULL,
                        st->scaling ? work->scaling->E->d_val : OSQP_NULL, st->scaling ? work->scaling->Einv->d_val : OSQP_NULL,
                        st->scaling ? work->scaling->c : (OSQPFloat)1.0, st->scaling ? work->scaling->cinv : (OSQPFloat)1.0,
                        &bs, d_x, d_y, d_i, d_i + nb, d_i + 2 * nb, d_i + 3 * nb, d_f, d_f + nb, d_f + 2 * nb);
    b200_event_record(e1);
    last_batch_kernel_ms = rc ? 0.0 : (double)b200_event_elapsed_ms(e0, e1);
    b200_event_destroy(e0);
    b200_event_destroy(e1);
  }
  if (rc) goto done;
  rc = 1;
  if (b200_copy_out(x_out, d_x, fn) || b200_copy_out(y_out, d_y, fm)) goto done;
  if (iters && b200_copy_out(iters, d_i, (size_t)nb * sizeof(OSQPInt))) goto done;
  if (status && b200_copy_out(status, d_i + nb, (size_t)nb * sizeof(OSQPInt))) goto done;
  if (cg_iters && b200_copy_out(cg_iters, d_i + 2 * nb, (size_t)nb * sizeof(OSQPInt))) goto done;
  if (rho_updates && b200_copy_out(rho_updates, d_i + 3 * nb, (size_t)nb * sizeof(OSQPInt))) goto done;
  if (obj && b200_copy_out(obj, d_f, (size_t)nb * sizeof(OSQPFloat))) goto done;
  if (prim_res && b200_copy_out(prim_res, d_f + nb, (size_t)nb * sizeof(OSQPFloat))) goto done;
  if (dual_res && b200_copy_out(dual_res, d_f + 2 * nb, (size_t)nb * sizeof(OSQPFloat))) goto done;
  rc = b200_last_error() ? 1 : 0;
done:
  b200_free(d_l); b200_free(d_u); b200_free(d_q); b200_free(d_x); b200_free(d_y); b200_free(d_f); b200_free(d_i);
  return rc;
}
