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
