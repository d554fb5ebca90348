/*
 * ORACLE / TEST INFRASTRUCTURE -- not part of the shipped B200 product path.
 *
 * Type header for the local restatement of QDLDL v0.1.8 (github.com/osqp/qdldl,
 * pinned by /root/reference/algebra/_common/lin_sys/qdldl/qdldl.cmake:8-10 and
 * absent from the reference tree).  The reference maps the QDLDL scalar types
 * onto OSQPInt / OSQPFloat / int
 * (algebra/_common/lin_sys/qdldl/qdldl_codegen_types.h.in:18-22); we do the same.
 */
#ifndef QDLDL_TYPES_H
#define QDLDL_TYPES_H

#include "osqp_api_types.h"
#include <limits.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef OSQPInt   QDLDL_int;
typedef OSQPFloat QDLDL_float;
typedef int       QDLDL_bool;

#ifdef OSQP_USE_LONG
#define QDLDL_INT_MAX LLONG_MAX
#else
#define QDLDL_INT_MAX INT_MAX
#endif

#ifdef __cplusplus
}
#endif

#endif
