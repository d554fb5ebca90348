/*
 * algebra/b200/algebra_libs.c -- library lifecycle + linear-system factory of the B200
 * backend.  Implements /root/reference/include/private/lin_alg.h:12-53
 * (reference CUDA counterpart: algebra/cuda/algebra_libs.cu:31-130).
 */
#include "osqp_api_constants.h"
#include "osqp_api_types.h"
#include "lin_alg.h"
#include "algebra_impl.h"
#include "pcg_interface.h"
#include "printing.h"

#include <stdio.h>

OSQPInt osqp_algebra_linsys_supported(void) {
  /* only the indirect (PCG) solver exists on the device */
  return OSQP_CAPABILITY_INDIRECT_SOLVER;
}

enum osqp_linsys_solver_type osqp_algebra_default_linsys(void) {
  return OSQP_INDIRECT_SOLVER;
}

OSQPInt osqp_algebra_init_libs(OSQPInt device) {
  /* 0 ok, 1 -> osqp_setup returns OSQP_ALGEBRA_LOAD_ERROR (osqp_api.c:378).  There is no
     CPU fallback: without a GPU the backend refuses to load. */
  return b200_init((int)device) ? 1 : 0;
}

void osqp_algebra_free_libs(void) {
  /* ref-counted: osqp_cleanup of one solver must not tear down the others
     (reference limitation: tests/basic_qp/test_basic_qp.cpp:845) */
  b200_shutdown();
}

OSQPInt osqp_algebra_name(char* name, OSQPInt nameLen) {
  return (OSQPInt)snprintf(name, (size_t)nameLen, "B200 (sm_100a)");
}

OSQPInt osqp_algebra_device_name(char* name, OSQPInt nameLen) {
  return (OSQPInt)b200_device_name(name, (int)nameLen);
}

OSQPInt osqp_algebra_init_linsys_solver(LinSysSolver**      s,
                                        const OSQPMatrix*   P,
                                        const OSQPMatrix*   A,
                                        const OSQPVectorf*  rho_vec,
                                        const OSQPSettings* settings,
                                        OSQPFloat*          scaled_prim_res,
                                        OSQPFloat*          scaled_dual_res,
                                        OSQPInt             polishing) {
  switch (settings->linsys_solver) {
  default:
  case OSQP_INDIRECT_SOLVER:
    return init_linsys_solver_b200pcg((b200pcg_solver**)s, P, A, rho_vec, settings,
                                      scaled_prim_res, scaled_dual_res, polishing);
  }
}
