/* ORACLE / BASELINE INFRASTRUCTURE.
 * Stand-in for the generated configure header, selecting the REFERENCE's own algebra/cuda backend
 * (cuSPARSE / cuBLAS / thrust) so that it can be timed on the same B200 as a second baseline
 * (BASELINE.md B2).  Double precision so that its numbers compare with the f64 product build;
 * -DREFCUDA_FLOAT gives the reference's default float32 build. */
#ifndef OSQP_CONFIGURE_H
#define OSQP_CONFIGURE_H
#define IS_LINUX
#define OSQP_ALGEBRA_CUDA
#define OSQP_ENABLE_PRINTING
#define OSQP_ENABLE_PROFILING
#ifdef REFCUDA_FLOAT
#define OSQP_USE_FLOAT
#endif
#endif
