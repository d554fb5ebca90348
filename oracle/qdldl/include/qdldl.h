/*
 * ORACLE / TEST INFRASTRUCTURE -- not part of the shipped B200 product path.
 *
 * Local restatement of the three QDLDL v0.1.8 entry points the reference calls:
 *   QDLDL_etree   qdldl_interface.c:94, :806
 *   QDLDL_factor  qdldl_interface.c:116-119, :485-487, :520-522, :816
 *   QDLDL_solve   qdldl_interface.c:409, :828, :846
 * Contract (inferred from those call sites):
 *   etree : returns sum(Lnz) >= 0, -1 if the matrix is not upper triangular or a
 *           column is empty, -2 on integer overflow of the count.
 *   factor: returns the number of strictly positive entries of D, or -1 when a
 *           zero pivot is met.
 *   solve : in-place solve of L D L' x = b with unit-lower L stored column-wise.
 * PARITY: unpinned at this boundary -- the upstream QDLDL source is not available
 * offline and the reference's only vtable-level known-answer test is disabled
 * (tests/lin_alg/lin_alg_tester.cpp:60-67).  The restatement is pinned instead
 * against scipy.sparse.linalg.splu and, end to end, against the reference's
 * golden solutions (tests/test_oracle_*.py).
 */
#ifndef QDLDL_H
#define QDLDL_H

#include "qdldl_types.h"
#include "qdldl_version.h"

#ifdef __cplusplus
extern "C" {
#endif

QDLDL_int QDLDL_etree(const QDLDL_int  n,
                      const QDLDL_int* Ap,
                      const QDLDL_int* Ai,
                      QDLDL_int*       work,
                      QDLDL_int*       Lnz,
                      QDLDL_int*       etree);

QDLDL_int QDLDL_factor(const QDLDL_int    n,
                       const QDLDL_int*   Ap,
                       const QDLDL_int*   Ai,
                       const QDLDL_float* Ax,
                       QDLDL_int*         Lp,
                       QDLDL_int*         Li,
                       QDLDL_float*       Lx,
                       QDLDL_float*       D,
                       QDLDL_float*       Dinv,
                       const QDLDL_int*   Lnz,
                       const QDLDL_int*   etree,
                       QDLDL_bool*        bwork,
                       QDLDL_int*         iwork,
                       QDLDL_float*       fwork);

void QDLDL_solve(const QDLDL_int    n,
                 const QDLDL_int*   Lp,
                 const QDLDL_int*   Li,
                 const QDLDL_float* Lx,
                 const QDLDL_float* Dinv,
                 QDLDL_float*       x);

void QDLDL_Lsolve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li,
                  const QDLDL_float* Lx, QDLDL_float* x);

void QDLDL_Ltsolve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li,
                   const QDLDL_float* Lx, QDLDL_float* x);

#ifdef __cplusplus
}
#endif

#endif
