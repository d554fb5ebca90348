/*
 * Build configuration for libosqp with the B200 algebra backend.  Stand-in for the header
 * CMake generates from /root/reference/configure/osqp_configure.h.in:13-49, with the new
 * backend symbol OSQP_ALGEBRA_B200 (the reference root CMakeLists.txt:94-100 only knows
 * builtin / mkl / cuda).  Precision is selected with -DB200_USE_FLOAT on the command line.
 */
#ifndef OSQP_CONFIGURE_H
#define OSQP_CONFIGURE_H
#define IS_LINUX
#define OSQP_ALGEBRA_B200
#define OSQP_ENABLE_PRINTING
#define OSQP_ENABLE_PROFILING
#ifdef B200_USE_FLOAT
#define OSQP_USE_FLOAT
#endif
#endif
