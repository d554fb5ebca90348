/* ORACLE / TEST INFRASTRUCTURE.
 * Hand-written stand-in for the header CMake generates from
 * /root/reference/configure/osqp_configure.h.in:13-49, selecting the reference
 * "builtin" (CSC loops + QDLDL) backend in double precision with 32-bit ints. */
#ifndef OSQP_CONFIGURE_H
#define OSQP_CONFIGURE_H
#define IS_LINUX
#define OSQP_ALGEBRA_BUILTIN
#define OSQP_ENABLE_PRINTING
#define OSQP_ENABLE_PROFILING
#endif
