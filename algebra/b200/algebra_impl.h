/*
 * algebra/b200 -- definitions of OSQP's opaque algebra types for the B200 backend.
 *
 * Counterpart of /root/reference/algebra/cuda/algebra_types.h:31-59 and
 * algebra/builtin/algebra_impl.h:16-45.  All numerical data lives in B200 HBM and is only
 * touched through the C-ABI in include/osqp_b200.h.
 */
#ifndef ALGEBRA_IMPL_H
#define ALGEBRA_IMPL_H

#include "osqp_api_types.h"
#include "osqp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* how a vector is laid out over the ranks of a row-sharded solve (B200_SHARD_*): fixed when the vector
 * is CREATED (constructors / views / copies in vector.c) and read by every reduction -- not re-derived
 * from its length at each use */
enum { B200_SHARD_NONE = 0, B200_SHARD_ROWS = 1, B200_SHARD_COLSPLIT = 2 };

struct OSQPVectori_ {
  OSQPInt* d_val;     /* device array */
  OSQPInt  length;
  OSQPInt  shard;     /* B200_SHARD_* */
};

struct OSQPVectorf_ {
  OSQPFloat* d_val;   /* device array (b200_float == OSQPFloat) */
  OSQPInt    length;
  OSQPInt    is_view; /* 1: d_val aliases a parent vector and is not freed */
  OSQPInt    shard;   /* B200_SHARD_* */
};

/*
 * A (m x n) is kept twice, as CSR of A and as CSR of A' (== the user's CSC arrays, so the
 * transpose costs nothing and A'x never needs atomics).  P (symmetric) is expanded from the
 * user's upper triangle to a full CSR with a structurally full diagonal.
 * The host maps translate positions in the USER's CSC value array into positions of the
 * device value arrays for OSQPMatrix_update_values (reference: d_A_to_At_ind /
 * d_P_triu_to_full_ind, algebra/cuda/algebra_types.h:51-59).
 */
struct OSQPMatrix_ {
  b200_csr* S;        /* CSR of the matrix itself (A, or full symmetric P)      */
  b200_csr* St;       /* CSR of A'; NULL for symmetric P                         */
  OSQPInt   m, n;     /* rows, columns                                           */
  OSQPInt   nnz_user; /* entries in the user's CSC (triu count for P)            */
  OSQPInt   is_symmetric;
  OSQPInt*  h_map;    /* user CSC k -> position in S (A: CSR pos; P: upper copy) */
  OSQPInt*  h_map2;   /* P only: position of the mirrored copy, -1 on diagonal   */
  OSQPInt*  d_map;    /* built on the device: the same map as h_map, in HBM (h_map NULL)   */
  OSQPInt*  d_map2;   /* P built on the device: the same map as h_map2, in HBM             */
};

/*
 * Row-sharded multi-GPU mode (one process per GPU).  The core runs unchanged
 * and identically on every rank; the backend is told the global column count n and the local row
 * count m_local once (osqp_b200_dist_configure, before osqp_setup) and from then on treats every
 * vector of length m_local as ROW-SHARDED (its reductions are combined across ranks) and every
 * other vector as replicated.  m_local must differ from n so that the classification is
 * unambiguous; the partitioner guarantees it.
 */
extern OSQPInt b200_dist_n;
extern OSQPInt b200_dist_mlocal;
#define B200_IS_SHARDED(len) (b200_dist_mlocal >= 0 && (len) == b200_dist_mlocal && b200_dist_world() > 1)
/*
 * Column-split refinement (osqp_b200_dist_configure_split): the rank's QP is
 * posed over [n_shared columns touched by several ranks ; the columns only this rank touches], so
 * b200_dist_n is the LOCAL column count and every vector of that length is COLUMN-SPLIT: its leading
 * n_shared entries are replicated on all ranks (reductions count them on rank 0 only), the rest is
 * owned.  n_local must differ from m_local (the partitioner pads a free row if needed).
 */
extern OSQPInt b200_dist_nshared;   /* -1: layout off */
extern OSQPInt b200_dist_nglobal;   /* global number of columns (mean over columns in scaling.c:123-124) */
#define B200_IS_COLSPLIT(len) (b200_dist_nshared >= 0 && (len) == b200_dist_n && b200_dist_world() > 1)
/* classification of a NEW vector of this length under the layout declared by osqp_b200_dist_configure*:
 * the core only ever creates vectors of n (columns), m (rows) and n + m entries between osqp_setup and
 * osqp_cleanup; polish (vectors of other lengths over subsets of the rows) is refused when sharded */
#define B200_SHARD_OF(len) (B200_IS_SHARDED(len) ? B200_SHARD_ROWS : (B200_IS_COLSPLIT(len) ? B200_SHARD_COLSPLIT : B200_SHARD_NONE))

/*
 * Scalar cache: the fused termination check (fused_admm.c) computes every norm the core asks for
 * right afterwards (compute_prim_tol / compute_dual_tol / compute_rho_estimate, src/auxil.c:14-47,
 * 334-458) in one kernel and parks them here keyed by (weight pointer, vector pointer).  An entry is
 * served only while b200_epoch() is unchanged, i.e. as long as no kernel or copy that could have
 * modified a vector has been issued since -- otherwise the reduction is recomputed as usual.
 */
#define B200_NORM_CACHE_MAX 12
typedef struct {
  unsigned long long epoch;
  int                count;
  const void*        s[B200_NORM_CACHE_MAX];   /* weight vector data pointer or NULL */
  const void*        v[B200_NORM_CACHE_MAX];
  size_t             bytes[B200_NORM_CACHE_MAX]; /* extent of v (and of s) */
  OSQPFloat          val[B200_NORM_CACHE_MAX];
} b200_norm_cache;

void b200_norm_cache_reset(void);
void b200_norm_cache_put(const void* s, const void* v, OSQPInt length, OSQPFloat val);
void b200_norm_cache_seal(void);                       /* stamp with the current epoch */
int  b200_norm_cache_get(const void* s, const void* v, OSQPFloat* val);
/* Operations of the termination check that are known to write only `dst` (or nothing, dst = NULL)
 * bracket themselves with these two calls: if the cache was live before the operation and `dst`
 * overlaps none of the cached vectors, the cache is re-stamped with the new epoch instead of dying.
 * Any operation that does NOT bracket itself still kills the cache (the safe default). */
int  b200_norm_cache_live(void);
void b200_norm_cache_after(int was_live, const void* dst, OSQPInt length);

#ifdef __cplusplus
}
#endif

#endif /* ALGEBRA_IMPL_H */
