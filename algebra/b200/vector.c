/*
 * algebra/b200/vector.c -- OSQPVectorf / OSQPVectori over B200 device memory.
 *
 * Implements /root/reference/include/private/algebra_vector.h:28-290.  Semantics follow the
 * CPU reference (algebra/builtin/vector.c), which is the parity oracle; the structure of the
 * container (device array + view flag, host-or-device raw pointers) follows the role of
 * algebra/cuda/vector.cu:43-237.  Every numerical operation is one call into the C-ABI
 * (include/osqp_b200.h) and therefore one sm_100a kernel on the library stream.
 */
#include "osqp.h"
#include "algebra_vector.h"
#include "algebra_impl.h"
#include "glob_opts.h"

/* ------------------------------------------------------------------ creation */

OSQPVectorf* OSQPVectorf_malloc(OSQPInt length) {
  OSQPVectorf* b = (OSQPVectorf*)c_malloc(sizeof(OSQPVectorf));
  if (!b) return OSQP_NULL;
  b->length  = length;
  b->is_view = 0;
  b->shard   = B200_SHARD_OF(length);
  b->d_val   = (OSQPFloat*)b200_malloc((size_t)length * sizeof(OSQPFloat));
  if (!b->d_val) {
    c_free(b);
    return OSQP_NULL;
  }
  return b;
}

OSQPVectorf* OSQPVectorf_calloc(OSQPInt length) {
  OSQPVectorf* b = (OSQPVectorf*)c_malloc(sizeof(OSQPVectorf));
  if (!b) return OSQP_NULL;
  b->length  = length;
  b->is_view = 0;
  b->shard   = B200_SHARD_OF(length);
  b->d_val   = (OSQPFloat*)b200_calloc((size_t)length * sizeof(OSQPFloat));
  if (!b->d_val) {
    c_free(b);
    return OSQP_NULL;
  }
  return b;
}

OSQPVectori* OSQPVectori_malloc(OSQPInt length) {
  OSQPVectori* b = (OSQPVectori*)c_malloc(sizeof(OSQPVectori));
  if (!b) return OSQP_NULL;
  b->length = length;
  b->shard  = B200_SHARD_OF(length);
  b->d_val  = (OSQPInt*)b200_malloc((size_t)length * sizeof(OSQPInt));
  if (!b->d_val) {
    c_free(b);
    return OSQP_NULL;
  }
  return b;
}

OSQPVectori* OSQPVectori_calloc(OSQPInt length) {
  OSQPVectori* b = (OSQPVectori*)c_malloc(sizeof(OSQPVectori));
  if (!b) return OSQP_NULL;
  b->length = length;
  b->shard  = B200_SHARD_OF(length);
  b->d_val  = (OSQPInt*)b200_calloc((size_t)length * sizeof(OSQPInt));
  if (!b->d_val) {
    c_free(b);
    return OSQP_NULL;
  }
  return b;
}

OSQPVectorf* OSQPVectorf_new(const OSQPFloat* a, OSQPInt length) {
  OSQPVectorf* out = OSQPVectorf_malloc(length);
  if (!out) return OSQP_NULL;
  if (length > 0) OSQPVectorf_from_raw(out, a);
  return out;
}

OSQPVectori* OSQPVectori_new(const OSQPInt* a, OSQPInt length) {
  OSQPVectori* out = OSQPVectori_malloc(length);
  if (!out) return OSQP_NULL;
  if (length > 0) OSQPVectori_from_raw(out, a);
  return out;
}

OSQPVectorf* OSQPVectorf_copy_new(const OSQPVectorf* a) {
  OSQPVectorf* b = OSQPVectorf_malloc(a->length);
  if (b) {
    b->shard = a->shard;      /* a copy is laid out like its source */
    OSQPVectorf_copy(b, a);
  }
  return b;
}

void OSQPVectorf_free(OSQPVectorf* a) {
  if (a) {
    if (!a->is_view) b200_free(a->d_val);
    c_free(a);
  }
}

void OSQPVectori_free(OSQPVectori* a) {
  if (a) {
    b200_free(a->d_val);
    c_free(a);
  }
}

/* views alias the parent's device storage (xtilde/ztilde inside xz_tilde, osqp_api.c:420-421) */
OSQPVectorf* OSQPVectorf_view(const OSQPVectorf* a, OSQPInt head, OSQPInt length) {
  OSQPVectorf* view = (OSQPVectorf*)c_malloc(sizeof(OSQPVectorf));
  if (view) {
    view->length  = length;
    view->is_view = 1;
    view->shard   = B200_SHARD_OF(length);   /* xtilde / ztilde inside xz_tilde: columns / rows */
    view->d_val   = a->d_val + head;
  }
  return view;
}

void OSQPVectorf_view_update(OSQPVectorf* a, const OSQPVectorf* b, OSQPInt head, OSQPInt length) {
  a->length = length;
  a->shard  = B200_SHARD_OF(length);
  a->d_val  = b->d_val + head;
}

void OSQPVectorf_view_free(OSQPVectorf* a) {
  if (a) c_free(a);
}

OSQPInt OSQPVectorf_length(const OSQPVectorf* a) { return a->length; }
OSQPInt OSQPVectori_length(const OSQPVectori* a) { return a->length; }

/* device pointer; only meaningful to device-aware callers */
OSQPFloat* OSQPVectorf_data(const OSQPVectorf* a) { return a->d_val; }

/* ---------------------------------------------------------------- raw copies */

void OSQPVectorf_copy(OSQPVectorf* b, const OSQPVectorf* a) {
  if (b == a) return;
  b200_copy_in(b->d_val, a->d_val, (size_t)a->length * sizeof(OSQPFloat));
}

/* `a` / `bv` may be host or device memory (reference: algebra/cuda/vector.cu:195-237) */
void OSQPVectorf_from_raw(OSQPVectorf* b, const OSQPFloat* a) {
  b200_copy_in(b->d_val, a, (size_t)b->length * sizeof(OSQPFloat));
}

void OSQPVectori_from_raw(OSQPVectori* b, const OSQPInt* a) {
  b200_copy_in(b->d_val, a, (size_t)b->length * sizeof(OSQPInt));
}

void OSQPVectorf_to_raw(OSQPFloat* bv, const OSQPVectorf* a) {
  b200_copy_out(bv, a->d_val, (size_t)a->length * sizeof(OSQPFloat));
}

void OSQPVectori_to_raw(OSQPInt* bv, const OSQPVectori* a) {
  b200_copy_out(bv, a->d_val, (size_t)a->length * sizeof(OSQPInt));
}

/* ------------------------------------------------------------- elementwise ops */

OSQPInt OSQPVectorf_is_eq(const OSQPVectorf* A, const OSQPVectorf* B, OSQPFloat tol) {
  if (A->length != B->length) return 0;
  return b200_vec_is_eq(A->d_val, B->d_val, tol, A->length);
}

void OSQPVectorf_set_scalar(OSQPVectorf* a, OSQPFloat sc) {
  b200_vec_set_scalar(a->d_val, sc, a->length);
}

void OSQPVectorf_set_scalar_conditional(OSQPVectorf* a, const OSQPVectori* test,
                                        OSQPFloat val_if_neg, OSQPFloat val_if_zero,
                                        OSQPFloat val_if_pos) {
  b200_vec_set_scalar_cond(a->d_val, test->d_val, val_if_neg, val_if_zero, val_if_pos, a->length);
}

void OSQPVectorf_round_to_zero(OSQPVectorf* a, OSQPFloat tol) {
  b200_vec_round_to_zero(a->d_val, tol, a->length);
}

void OSQPVectorf_mult_scalar(OSQPVectorf* a, OSQPFloat sc) {
  b200_vec_mult_scalar(a->d_val, sc, a->length);
}

void OSQPVectorf_plus(OSQPVectorf* x, const OSQPVectorf* a, const OSQPVectorf* b) {
  b200_vec_add_scaled(x->d_val, 1.0, a->d_val, 1.0, b->d_val, a->length);
}

void OSQPVectorf_minus(OSQPVectorf* x, const OSQPVectorf* a, const OSQPVectorf* b) {
  b200_vec_add_scaled(x->d_val, 1.0, a->d_val, -1.0, b->d_val, a->length);
}

void OSQPVectorf_add_scaled(OSQPVectorf* x, OSQPFloat sca, const OSQPVectorf* a, OSQPFloat scb,
                            const OSQPVectorf* b) {
  b200_vec_add_scaled(x->d_val, sca, a->d_val, scb, b->d_val, x->length);
}

void OSQPVectorf_add_scaled3(OSQPVectorf* x, OSQPFloat sca, const OSQPVectorf* a, OSQPFloat scb,
                             const OSQPVectorf* b, OSQPFloat scc, const OSQPVectorf* c) {
  b200_vec_add_scaled3(x->d_val, sca, a->d_val, scb, b->d_val, scc, c->d_val, x->length);
}

void OSQPVectorf_ew_prod(OSQPVectorf* c, const OSQPVectorf* a, const OSQPVectorf* b) {
  int live = b200_norm_cache_live();
  b200_vec_ew_prod(c->d_val, a->d_val, b->d_val, c->length);
  b200_norm_cache_after(live, c->d_val, c->length);
}

void OSQPVectorf_ew_bound_vec(OSQPVectorf* x, const OSQPVectorf* z, const OSQPVectorf* l,
                              const OSQPVectorf* u) {
  b200_vec_ew_bound(x->d_val, z->d_val, l->d_val, u->d_val, x->length);
}

void OSQPVectorf_project_polar_reccone(OSQPVectorf* y, const OSQPVectorf* l, const OSQPVectorf* u,
                                       OSQPFloat infval) {
  int live = b200_norm_cache_live();
  b200_vec_project_polar_reccone(y->d_val, l->d_val, u->d_val, infval, y->length);
  b200_norm_cache_after(live, y->d_val, y->length);
}

void OSQPVectorf_ew_reciprocal(OSQPVectorf* b, const OSQPVectorf* a) {
  b200_vec_ew_reciprocal(b->d_val, a->d_val, a->length);
}

void OSQPVectorf_ew_sqrt(OSQPVectorf* a) { b200_vec_ew_sqrt(a->d_val, a->length); }

void OSQPVectorf_ew_max_vec(OSQPVectorf* c, const OSQPVectorf* a, const OSQPVectorf* b) {
  b200_vec_ew_max(c->d_val, a->d_val, b->d_val, c->length);
}

void OSQPVectorf_ew_min_vec(OSQPVectorf* c, const OSQPVectorf* a, const OSQPVectorf* b) {
  b200_vec_ew_min(c->d_val, a->d_val, b->d_val, c->length);
}

void OSQPVectorf_set_scalar_if_lt(OSQPVectorf* x, const OSQPVectorf* z, OSQPFloat testval,
                                  OSQPFloat newval) {
  b200_vec_set_scalar_if_lt(x->d_val, z->d_val, testval, newval, x->length);
}

void OSQPVectorf_set_scalar_if_gt(OSQPVectorf* x, const OSQPVectorf* z, OSQPFloat testval,
                                  OSQPFloat newval) {
  b200_vec_set_scalar_if_gt(x->d_val, z->d_val, testval, newval, x->length);
}

/* ------------------------------------------------------------------ reductions
 * each returns a host scalar by value and is therefore one stream synchronisation */

OSQPInt b200_dist_n       = -1;
OSQPInt b200_dist_mlocal  = -1;
OSQPInt b200_dist_nshared = -1;
OSQPInt b200_dist_nglobal = -1;

/* Reductions over row-sharded vectors are combined across ranks inside the kernel library.  Over a
 * column-split vector a rank other than 0 skips the replicated leading slice: `expr` must address
 * its operands through RED_OFF / RED_CNT. */
#define DIST_REDUCE(vec, expr)                                          \
  do {                                                                  \
    OSQPInt RED_OFF = 0, RED_CNT = (vec)->length;                       \
    int     dist_   = b200_dist_world() > 1;                            \
    int     split_  = dist_ && (vec)->shard == B200_SHARD_COLSPLIT;     \
    if (split_ && b200_dist_rank() > 0) {                               \
      RED_OFF = b200_dist_nshared;                                      \
      RED_CNT = (vec)->length - RED_OFF;                                \
    }                                                                   \
    (void)RED_OFF; (void)RED_CNT;                                       \
    if ((dist_ && (vec)->shard == B200_SHARD_ROWS) || split_) {         \
      b200_dist_scope(1);                                               \
      expr;                                                             \
      b200_dist_scope(0);                                               \
    } else {                                                            \
      expr;                                                             \
    }                                                                   \
  } while (0)

static _Thread_local b200_norm_cache g_cache;   /* one per host thread, like the library context */

void b200_norm_cache_reset(void) { g_cache.count = 0; g_cache.epoch = 0; }

void b200_norm_cache_put(const void* s, const void* v, OSQPInt length, OSQPFloat val) {
  if (g_cache.count < B200_NORM_CACHE_MAX) {
    g_cache.s[g_cache.count]     = s;
    g_cache.v[g_cache.count]     = v;
    g_cache.bytes[g_cache.count] = (size_t)length * sizeof(OSQPFloat);
    g_cache.val[g_cache.count]   = val;
    g_cache.count++;
  }
}

int b200_norm_cache_live(void) { return g_cache.count > 0 && g_cache.epoch == b200_epoch(); }

static int ranges_overlap(const void* a, size_t na, const void* b, size_t nb) {
  const char* pa = (const char*)a;
  const char* pb = (const char*)b;
  return a && b && pa < pb + nb && pb < pa + na;
}

void b200_norm_cache_after(int was_live, const void* dst, OSQPInt length) {
  int    i;
  size_t nb = (size_t)(length > 0 ? length : 0) * sizeof(OSQPFloat);
  if (!was_live) return;
  if (dst && nb) {
    for (i = 0; i < g_cache.count; i++) {
      if (ranges_overlap(dst, nb, g_cache.v[i], g_cache.bytes[i]) ||
          ranges_overlap(dst, nb, g_cache.s[i], g_cache.bytes[i])) {
        g_cache.count = 0;
        return;
      }
    }
  }
  g_cache.epoch = b200_epoch();
}

void b200_norm_cache_seal(void) { g_cache.epoch = b200_epoch(); }

int b200_norm_cache_get(const void* s, const void* v, OSQPFloat* val) {
  int i;
  if (g_cache.count == 0 || g_cache.epoch != b200_epoch()) return 0;
  for (i = 0; i < g_cache.count; i++) {
    if (g_cache.s[i] == s && g_cache.v[i] == v) {
      *val = g_cache.val[i];
      return 1;
    }
  }
  return 0;
}

OSQPFloat OSQPVectorf_norm_inf(const OSQPVectorf* v) {
  OSQPFloat cached;
  int       live = b200_norm_cache_live();
  if (b200_norm_cache_get(OSQP_NULL, v->d_val, &cached)) return cached;
  DIST_REDUCE(v, cached = b200_vec_norm_inf(v->d_val + RED_OFF, RED_CNT));
  b200_norm_cache_after(live, OSQP_NULL, 0);     /* a reduction writes no vector */
  return cached;
}

OSQPFloat OSQPVectorf_scaled_norm_inf(const OSQPVectorf* S, const OSQPVectorf* v) {
  OSQPFloat cached;
  int       live = b200_norm_cache_live();
  if (b200_norm_cache_get(S->d_val, v->d_val, &cached)) return cached;
  DIST_REDUCE(v, cached = b200_vec_scaled_norm_inf(S->d_val + RED_OFF, v->d_val + RED_OFF, RED_CNT));
  b200_norm_cache_after(live, OSQP_NULL, 0);
  return cached;
}

OSQPFloat OSQPVectorf_norm_inf_diff(const OSQPVectorf* a, const OSQPVectorf* b) {
  OSQPFloat r;
  DIST_REDUCE(a, r = b200_vec_norm_inf_diff(a->d_val + RED_OFF, b->d_val + RED_OFF, RED_CNT));
  return r;
}

OSQPFloat OSQPVectorf_norm_1(const OSQPVectorf* a) {
  OSQPFloat r;
  DIST_REDUCE(a, r = b200_vec_norm_1(a->d_val + RED_OFF, RED_CNT));
  /* the only caller (scaling.c:123-124) divides by the LOCAL n to get a mean over columns: hand it
     the global sum rescaled so that the quotient is the global mean */
  if (b200_dist_world() > 1 && a->shard == B200_SHARD_COLSPLIT && b200_dist_nglobal > 0)
    r = r * (OSQPFloat)a->length / (OSQPFloat)b200_dist_nglobal;
  return r;
}

OSQPFloat OSQPVectorf_norm_2(const OSQPVectorf* a) { return b200_vec_norm_2(a->d_val, a->length); }

OSQPFloat OSQPVectorf_dot_prod(const OSQPVectorf* a, const OSQPVectorf* b) {
  OSQPFloat r;
  int       live = b200_norm_cache_live();
  DIST_REDUCE(a, r = b200_vec_dot(a->d_val + RED_OFF, b->d_val + RED_OFF, RED_CNT));
  b200_norm_cache_after(live, OSQP_NULL, 0);
  return r;
}

OSQPFloat OSQPVectorf_dot_prod_signed(const OSQPVectorf* a, const OSQPVectorf* b, OSQPInt sign) {
  OSQPFloat r;
  int       live = b200_norm_cache_live();
  DIST_REDUCE(a, r = b200_vec_dot_signed(a->d_val + RED_OFF, b->d_val + RED_OFF, (int)sign, RED_CNT));
  b200_norm_cache_after(live, OSQP_NULL, 0);
  return r;
}

OSQPInt OSQPVectorf_all_leq(const OSQPVectorf* l, const OSQPVectorf* u) {
  OSQPInt r;
  DIST_REDUCE(l, r = b200_vec_all_leq(l->d_val, u->d_val, l->length));
  return r;
}

OSQPInt OSQPVectorf_in_reccone(const OSQPVectorf* y, const OSQPVectorf* l, const OSQPVectorf* u,
                               OSQPFloat infval, OSQPFloat tol) {
  OSQPInt r;
  int     live = b200_norm_cache_live();
  DIST_REDUCE(y, r = b200_vec_in_reccone(y->d_val, l->d_val, u->d_val, infval, tol, y->length));
  b200_norm_cache_after(live, OSQP_NULL, 0);
  return r;
}

OSQPInt OSQPVectorf_ew_bounds_type(OSQPVectori* iseq, const OSQPVectorf* l, const OSQPVectorf* u,
                                   OSQPFloat tol, OSQPFloat infval) {
  OSQPInt r;
  DIST_REDUCE(iseq,
              r = b200_vec_bounds_type(iseq->d_val, l->d_val, u->d_val, tol, infval, iseq->length));
  return r;
}
