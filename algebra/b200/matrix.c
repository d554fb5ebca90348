/*
 * algebra/b200/matrix.c -- OSQPMatrix over B200 device CSR storage.
 *
 * Implements /root/reference/include/private/algebra_matrix.h:26-138.  Numerical contract =
 * CPU reference (algebra/builtin/matrix.c, algebra/_common/csc_math.c); storage design takes
 * over the role of algebra/cuda/matrix.cu:32-172 + src/cuda_csr.cu:489-714 (CSR of A, stored
 * transpose, full symmetric P with guaranteed diagonal, index maps for value updates) without
 * cuSPARSE/thrust:
 *   - the user's CSC of A *is* the CSR of A' -> one upload, identity value map;
 *   - CSR of A = transpose of that copy ON THE DEVICE (osqp_b200/csrc/transpose.cu: count, scan,
 *     scatter, per-row rank sort), index map kept in HBM;
 *   - full P = lower mirror + upper copy of the triu CSC, diagonal inserted where missing, also
 *     expanded on the device (b200_csr_symmetric_from_triu).
 * The host builders below (a parallel stable counting sort and a serial symmetric expansion) are
 * the fall-back for matrices with rows longer than the device rank-sort limit and for
 * B200_HOST_TRANSPOSE=1; both paths produce bit-identical matrices.
 */
#include "osqp.h"
#include "algebra_matrix.h"
#include "algebra_impl.h"
#include "glob_opts.h"
#include "printing.h"

#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

/* B200_TRACE_SETUP=1 prints where setup time goes (development aid) */
static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec;
}
static int trace_on(void) {
  static int v = -1;
  if (v < 0) v = getenv("B200_TRACE_SETUP") ? 1 : 0;
  return v;
}

/* ------------------------------------------------------------------ builders */

/* ---- fall-back: parallel stable counting sort (CSC -> CSR) on the host ------------------------
 * One core needs 205 ms for the 1.14e7-nnz Lasso (measured), so the sort is split over host threads:
 * thread t owns a contiguous chunk of COLUMNS; (1) it histograms the rows of its chunk,
 * (2) row offsets are the prefix sum over rows of the summed histograms, and every thread's
 * private cursor for row i starts after the entries of the threads before it, (3) it scatters its
 * chunk in order.  The result is identical to the serial sort (stable, columns ascending within
 * a row) whatever the thread count. */
#include <pthread.h>
#include <unistd.h>

typedef struct {
  int              tid, nthreads;
  OSQPInt          m, n;
  const OSQPInt*   Ap;
  const OSQPInt*   Ai;
  const OSQPFloat* Ax;
  OSQPInt*         rp;      /* m + 1 */
  OSQPInt*         ci;
  OSQPFloat*       vx;
  OSQPInt*         map;
  OSQPInt*         hist;    /* nthreads x m */
  OSQPInt          j0, j1;  /* column chunk */
  pthread_barrier_t* bar;
} tr_job;

static void* tr_worker(void* arg) {
  tr_job*  J = (tr_job*)arg;
  OSQPInt  m = J->m, i, j, k, t;
  OSQPInt* my = J->hist + (size_t)J->tid * m;

  memset(my, 0, (size_t)m * sizeof(OSQPInt));
  for (k = J->Ap[J->j0]; k < J->Ap[J->j1]; k++) my[J->Ai[k]]++;
  pthread_barrier_wait(J->bar);

  /* per-row totals -> rp[i + 1] (row range split over threads), then thread 0 scans */
  {
    OSQPInt r0 = (OSQPInt)((long long)m * J->tid / J->nthreads);
    OSQPInt r1 = (OSQPInt)((long long)m * (J->tid + 1) / J->nthreads);
    for (i = r0; i < r1; i++) {
      OSQPInt tot = 0;
      for (t = 0; t < J->nthreads; t++) {
        OSQPInt c = J->hist[(size_t)t * m + i];
        J->hist[(size_t)t * m + i] = tot;      /* exclusive prefix over threads */
        tot += c;
      }
      J->rp[i + 1] = tot;
    }
  }
  pthread_barrier_wait(J->bar);
  if (J->tid == 0) {
    J->rp[0] = 0;
    for (i = 0; i < m; i++) J->rp[i + 1] += J->rp[i];
  }
  pthread_barrier_wait(J->bar);

  /* scatter: cursor of (thread, row) = rp[row] + entries of earlier threads in that row */
  for (j = J->j0; j < J->j1; j++) {
    for (k = J->Ap[j]; k < J->Ap[j + 1]; k++) {
      OSQPInt row = J->Ai[k];
      OSQPInt pos = J->rp[row] + my[row]++;
      J->ci[pos] = j;
      J->vx[pos] = J->Ax[k];
      if (J->map) J->map[k] = pos;
    }
  }
  return OSQP_NULL;
}

static int host_threads(OSQPInt nnz) {
  long nc = sysconf(_SC_NPROCESSORS_ONLN);
  const char* env = getenv("B200_SETUP_THREADS");
  int t = env ? atoi(env) : (nc > 16 ? 16 : (int)nc);
  if (t < 1) t = 1;
  if (nnz < 200000) t = 1;        /* not worth the thread start-up */
  return t;
}

/* CSR of an m x n matrix from its CSC arrays; map[k] = CSR position of CSC entry k */
static b200_csr* csr_from_csc(OSQPInt m, OSQPInt n, const OSQPInt* Ap, const OSQPInt* Ai,
                              const OSQPFloat* Ax, OSQPInt* map) {
  OSQPInt   nnz = Ap[n];
  int       T   = host_threads(nnz), t;
  b200_csr* out = OSQP_NULL;
  OSQPInt*   rp   = (OSQPInt*)c_calloc((size_t)m + 2, sizeof(OSQPInt));
  OSQPInt*   hist = (OSQPInt*)c_malloc(((size_t)m + 1) * (size_t)T * sizeof(OSQPInt));
  OSQPInt*   ci   = (OSQPInt*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPInt));
  OSQPFloat* vx   = (OSQPFloat*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPFloat));
  tr_job*    jobs = (tr_job*)c_calloc((size_t)T, sizeof(tr_job));
  pthread_t* th   = (pthread_t*)c_calloc((size_t)T, sizeof(pthread_t));
  pthread_barrier_t bar;

  if (rp && hist && ci && vx && jobs && th) {
    OSQPInt j = 0;
    pthread_barrier_init(&bar, OSQP_NULL, (unsigned)T);
    for (t = 0; t < T; t++) {
      /* column chunks balanced by entries */
      OSQPInt target = (OSQPInt)((long long)nnz * (t + 1) / T);
      tr_job* J = &jobs[t];
      J->tid = t; J->nthreads = T; J->m = m; J->n = n; J->Ap = Ap; J->Ai = Ai; J->Ax = Ax;
      J->rp = rp; J->ci = ci; J->vx = vx; J->map = map; J->hist = hist; J->bar = &bar;
      J->j0 = j;
      if (t == T - 1) j = n;
      else while (j < n && Ap[j] < target) j++;
      J->j1 = j;
    }
    for (t = 1; t < T; t++) pthread_create(&th[t], OSQP_NULL, tr_worker, &jobs[t]);
    tr_worker(&jobs[0]);
    for (t = 1; t < T; t++) pthread_join(th[t], OSQP_NULL);
    pthread_barrier_destroy(&bar);
    out = b200_csr_create((int)m, (int)n, (int)nnz, rp, ci, vx);
  }
  c_free(rp); c_free(hist); c_free(ci); c_free(vx); c_free(jobs); c_free(th);
  return out;
}

/* full symmetric CSR (structurally full diagonal) from an upper-triangular CSC.
 * map_u[k]: position of entry k itself (row i, col j); map_l[k]: position of its mirror
 * (row j, col i) or -1 for diagonal entries.  Rows come out with sorted columns when the
 * CSC columns are sorted. */
static b200_csr* full_from_triu(OSQPInt n, const OSQPInt* Pp, const OSQPInt* Pi,
                                const OSQPFloat* Px, OSQPInt* map_u, OSQPInt* map_l) {
  OSQPInt   nnz = Pp[n];
  OSQPInt   i, j, k, pos, nnz_full;
  b200_csr* out = OSQP_NULL;
  OSQPInt* rp       = (OSQPInt*)c_calloc((size_t)n + 2, sizeof(OSQPInt));
  OSQPInt* next     = (OSQPInt*)c_malloc(((size_t)n + 1) * sizeof(OSQPInt));
  char*    has_diag = (char*)c_calloc((size_t)n + 1, 1);
  OSQPInt*   ci = OSQP_NULL;
  OSQPFloat* vx = OSQP_NULL;

  if (!rp || !next || !has_diag) goto done;

  for (j = 0; j < n; j++) {
    for (k = Pp[j]; k < Pp[j + 1]; k++) {
      i = Pi[k];
      rp[i + 1]++;                       /* (i, j) */
      if (i != j) rp[j + 1]++;           /* mirror (j, i) */
      else has_diag[j] = 1;
    }
  }
  for (i = 0; i < n; i++) {
    if (!has_diag[i]) rp[i + 1]++;       /* explicit zero on the diagonal */
    rp[i + 1] += rp[i];
  }
  nnz_full = rp[n];
  ci = (OSQPInt*)c_malloc(((size_t)nnz_full + 1) * sizeof(OSQPInt));
  vx = (OSQPFloat*)c_malloc(((size_t)nnz_full + 1) * sizeof(OSQPFloat));
  if (!ci || !vx) goto done;
  for (i = 0; i < n; i++) next[i] = rp[i];

  /* pass 1: strictly-lower mirrors, row j receives columns i < j in CSC order */
  for (j = 0; j < n; j++) {
    for (k = Pp[j]; k < Pp[j + 1]; k++) {
      i = Pi[k];
      if (i != j) {
        pos      = next[j]++;
        ci[pos]  = i;
        vx[pos]  = Px[k];
        map_l[k] = pos;
      } else {
        map_l[k] = -1;
      }
    }
  }
  /* missing diagonals sit between the lower and the upper part */
  for (i = 0; i < n; i++) {
    if (!has_diag[i]) {
      pos     = next[i]++;
      ci[pos] = i;
      vx[pos] = 0.0;
    }
  }
  /* pass 2: the upper triangle itself, row i receives columns j >= i in ascending j */
  for (j = 0; j < n; j++) {
    for (k = Pp[j]; k < Pp[j + 1]; k++) {
      i        = Pi[k];
      pos      = next[i]++;
      ci[pos]  = j;
      vx[pos]  = Px[k];
      map_u[k] = pos;
    }
  }
  (void)nnz;
  out = b200_csr_create((int)n, (int)n, (int)nnz_full, rp, ci, vx);

done:
  c_free(rp); c_free(next); c_free(has_diag); c_free(ci); c_free(vx);
  return out;
}

OSQPMatrix* OSQPMatrix_new_from_csc(const OSQPCscMatrix* M, OSQPInt is_triu) {
  double      t0  = now_ms(), t1;
  OSQPInt     nnz = M->p[M->n];
  OSQPMatrix* out = (OSQPMatrix*)c_calloc(1, sizeof(OSQPMatrix));
  if (!out) return OSQP_NULL;

  out->m            = M->m;
  out->n            = M->n;
  out->nnz_user     = nnz;
  out->is_symmetric = is_triu ? 1 : 0;

  if (is_triu) {
    out->St = OSQP_NULL;
    /* symmetric expansion on the device (upload the triangle once, transpose, merge); the host
       expansion remains for over-long rows, empty matrices and B200_HOST_TRANSPOSE=1 */
    if (!getenv("B200_HOST_TRANSPOSE"))
      out->S = b200_csr_symmetric_from_triu((int)M->n, M->p, M->i, M->x, &out->d_map, &out->d_map2);
    if (!out->S) {
      out->d_map = out->d_map2 = OSQP_NULL;
      out->h_map = (OSQPInt*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPInt));
      if (!out->h_map) goto fail;
      out->h_map2 = (OSQPInt*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPInt));
      if (!out->h_map2) goto fail;
      out->S = full_from_triu(M->n, M->p, M->i, M->x, out->h_map, out->h_map2);
    }
    if (!out->S) goto fail;
  } else {
    /* CSC(A) == CSR(A') */
    out->St = b200_csr_create((int)M->n, (int)M->m, (int)nnz, M->p, M->i, M->x);
    t1 = now_ms();
    if (trace_on()) fprintf(stderr, "[b200 trace] A: upload A' %.1f ms\n", t1 - t0);
    if (!out->St) goto fail;
    /* CSR(A): transposed on the device from the copy that is already there; the host counting
       sort + second upload remain for matrices with over-long rows (and B200_HOST_TRANSPOSE=1) */
    if (!getenv("B200_HOST_TRANSPOSE")) out->S = b200_csr_transpose(out->St, &out->d_map);
    if (!out->S) {
      out->d_map = OSQP_NULL;
      out->h_map = (OSQPInt*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPInt));
      if (!out->h_map) goto fail;
      out->S = csr_from_csc(M->m, M->n, M->p, M->i, M->x, out->h_map);
    }
    if (!out->S) goto fail;
  }
  if (trace_on()) { b200_sync(); fprintf(stderr, "[b200 trace] new_from_csc(%s) %.1f ms\n", is_triu ? "P" : "A", now_ms() - t0); }
  return out;

fail:
  OSQPMatrix_free(out);
  return OSQP_NULL;
}

void OSQPMatrix_free(OSQPMatrix* M) {
  if (M) {
    b200_csr_destroy(M->S);
    b200_csr_destroy(M->St);
    c_free(M->h_map);
    c_free(M->h_map2);
    b200_free(M->d_map);
    b200_free(M->d_map2);
    c_free(M);
  }
}

/* ------------------------------------------------------------------ accessors */

OSQPInt OSQPMatrix_get_m(const OSQPMatrix* M) { return M->m; }
OSQPInt OSQPMatrix_get_n(const OSQPMatrix* M) { return M->n; }
/* triu count for P: osqp_update_data_mat validates against it (osqp_api.c:1333-1347) */
OSQPInt OSQPMatrix_get_nz(const OSQPMatrix* M) { return M->nnz_user; }

OSQPInt OSQPMatrix_is_eq(const OSQPMatrix* A, const OSQPMatrix* B, OSQPFloat tol) {
  if (A->is_symmetric != B->is_symmetric) return 0;
  return b200_csr_is_eq(A->S, B->S, tol);
}

/* ---------------------------------------------------------------- value update
 * Mx_new_idx indexes the USER's CSC value array (triu for P); NULL = all values in CSC order
 * (algebra/_common/csc_math.c:27-45). */
void OSQPMatrix_update_values(OSQPMatrix* M, const OSQPFloat* Mx_new, const OSQPInt* Mx_new_idx,
                              OSQPInt M_new_n) {
  OSQPInt    k, cnt, nl;
  OSQPInt*   h_idx  = OSQP_NULL;
  OSQPInt*   h_idx2 = OSQP_NULL;
  OSQPInt*   h_sel  = OSQP_NULL;
  OSQPFloat* d_x    = OSQP_NULL;
  OSQPFloat* d_x2   = OSQP_NULL;
  OSQPInt*   d_idx  = OSQP_NULL;

  cnt = Mx_new_idx ? M_new_n : M->nnz_user;
  if (cnt <= 0) return;

  d_x   = (OSQPFloat*)b200_malloc((size_t)cnt * sizeof(OSQPFloat));
  d_idx = (OSQPInt*)b200_malloc((size_t)cnt * sizeof(OSQPInt));
  if (!d_x || !d_idx) goto done;
  b200_copy_in(d_x, Mx_new, (size_t)cnt * sizeof(OSQPFloat));

  if (M->d_map) {
    /* device-resident index maps: positions are looked up in HBM.  A: position in CSR(A), and the
       user's own ordering for A'.  P: the entry itself and its mirror (-1 on the diagonal). */
    if (!Mx_new_idx) {
      b200_vec_scatter(b200_csr_values(M->S), d_x, M->d_map, (int)cnt);
      if (M->is_symmetric) b200_vec_scatter_nonneg(b200_csr_values(M->S), d_x, M->d_map2, (int)cnt);
      else b200_copy_in(b200_csr_values(M->St), d_x, (size_t)cnt * sizeof(OSQPFloat));
    } else {
      OSQPInt* d_pos = (OSQPInt*)b200_malloc((size_t)cnt * sizeof(OSQPInt));
      if (!d_pos) goto done;
      b200_copy_in(d_idx, Mx_new_idx, (size_t)cnt * sizeof(OSQPInt));
      b200_veci_gather(d_pos, M->d_map, d_idx, (int)cnt);
      b200_vec_scatter(b200_csr_values(M->S), d_x, d_pos, (int)cnt);
      if (M->is_symmetric) {
        b200_veci_gather(d_pos, M->d_map2, d_idx, (int)cnt);
        b200_vec_scatter_nonneg(b200_csr_values(M->S), d_x, d_pos, (int)cnt);
      } else {
        b200_vec_scatter(b200_csr_values(M->St), d_x, d_idx, (int)cnt);
      }
      b200_free(d_pos);
    }
    goto done;
  }
  h_idx = (OSQPInt*)c_malloc((size_t)cnt * sizeof(OSQPInt));
  if (!h_idx) goto done;

  /* positions in S (for P: the upper copy) */
  for (k = 0; k < cnt; k++) h_idx[k] = M->h_map[Mx_new_idx ? Mx_new_idx[k] : k];
  b200_copy_in(d_idx, h_idx, (size_t)cnt * sizeof(OSQPInt));
  b200_vec_scatter(b200_csr_values(M->S), d_x, d_idx, (int)cnt);

  if (!M->is_symmetric) {
    /* A' shares the user's CSC ordering */
    if (!Mx_new_idx) {
      b200_copy_in(b200_csr_values(M->St), d_x, (size_t)cnt * sizeof(OSQPFloat));
    } else {
      b200_sync();   /* h_idx was the staging source of the previous copy */
      b200_copy_in(d_idx, Mx_new_idx, (size_t)cnt * sizeof(OSQPInt));
      b200_vec_scatter(b200_csr_values(M->St), d_x, d_idx, (int)cnt);
    }
  } else {
    /* mirrored (strictly lower) copies */
    h_idx2 = (OSQPInt*)c_malloc((size_t)cnt * sizeof(OSQPInt));
    h_sel  = (OSQPInt*)c_malloc((size_t)cnt * sizeof(OSQPInt));
    if (!h_idx2 || !h_sel) goto done;
    nl = 0;
    for (k = 0; k < cnt; k++) {
      OSQPInt pos = M->h_map2[Mx_new_idx ? Mx_new_idx[k] : k];
      if (pos >= 0) {
        h_idx2[nl] = pos;
        h_sel[nl]  = k;
        nl++;
      }
    }
    if (nl > 0) {
      d_x2 = (OSQPFloat*)b200_malloc((size_t)nl * sizeof(OSQPFloat));
      if (!d_x2) goto done;
      b200_sync();
      b200_copy_in(d_idx, h_sel, (size_t)nl * sizeof(OSQPInt));
      b200_vec_gather(d_x2, d_x, d_idx, (int)nl);
      b200_sync();
      b200_copy_in(d_idx, h_idx2, (size_t)nl * sizeof(OSQPInt));
      b200_vec_scatter(b200_csr_values(M->S), d_x2, d_idx, (int)nl);
    }
  }

done:
  b200_sync();
  b200_free(d_x); b200_free(d_x2); b200_free(d_idx);
  c_free(h_idx); c_free(h_idx2); c_free(h_sel);
}

/* ------------------------------------------------------------------- scalings */

void OSQPMatrix_mult_scalar(OSQPMatrix* A, OSQPFloat sc) {
  b200_csr_scale(A->S, sc);
  if (A->St) b200_csr_scale(A->St, sc);
}

/* A = diag(L) A : rows of A, columns of A' */
void OSQPMatrix_lmult_diag(OSQPMatrix* A, const OSQPVectorf* L) {
  b200_csr_scale_rows(A->S, L->d_val);
  if (A->St) b200_csr_scale_cols(A->St, L->d_val);
}

/* A = A diag(R) : columns of A, rows of A' */
void OSQPMatrix_rmult_diag(OSQPMatrix* A, const OSQPVectorf* R) {
  b200_csr_scale_cols(A->S, R->d_val);
  if (A->St) b200_csr_scale_rows(A->St, R->d_val);
}

/* ------------------------------------------------------------------- products */

/* y = alpha A x + beta y.  For P the full symmetric CSR makes this the symmetric product
 * that csc_Axpy_sym_triu computes on the CPU (csc_math.c:114-166). */
void OSQPMatrix_Axpy(const OSQPMatrix* A, const OSQPVectorf* x, OSQPVectorf* y, OSQPFloat alpha,
                     OSQPFloat beta) {
  int live;
  if (y->length <= 0) return;
  live = b200_norm_cache_live();
  b200_csr_spmv(A->S, x->d_val, y->d_val, alpha, beta);
  b200_norm_cache_after(live, y->d_val, y->length);
}

/* y = alpha A' x + beta y through the stored transpose (no atomics) */
void OSQPMatrix_Atxpy(const OSQPMatrix* A, const OSQPVectorf* x, OSQPVectorf* y, OSQPFloat alpha,
                      OSQPFloat beta) {
  if (y->length <= 0) return;
  if (!A->is_symmetric && b200_dist_world() > 1 && b200_dist_mlocal >= 0 && A->m == b200_dist_mlocal &&
      x->shard == B200_SHARD_ROWS) {   /* the rank's row block of the constraint matrix, not e.g. a polish submatrix */
    /* row-sharded A: A'x = sum over ranks of A_r' x_r -> one all-reduce of the length-n result.
       beta y must be added once, after the exchange. */
    /* column-split layout: only the shared leading slice has contributions from other ranks */
    int nx = b200_dist_nshared >= 0 ? (int)b200_dist_nshared : (int)y->length;
    if (beta == 0.0) {
      b200_csr_spmv(A->St, x->d_val, y->d_val, alpha, 0.0);
      b200_dist_allreduce_sum(y->d_val, nx);
    } else {
      OSQPFloat* tmp = (OSQPFloat*)b200_malloc((size_t)y->length * sizeof(OSQPFloat));
      if (!tmp) return;
      b200_csr_spmv(A->St, x->d_val, tmp, alpha, 0.0);
      b200_dist_allreduce_sum(tmp, nx);
      b200_vec_add_scaled(y->d_val, 1.0, tmp, beta, y->d_val, (int)y->length);
      b200_free(tmp);
    }
    return;
  }
  {
    int live = b200_norm_cache_live();
    b200_csr_spmv(A->is_symmetric ? A->S : A->St, x->d_val, y->d_val, alpha, beta);
    b200_norm_cache_after(live, y->d_val, y->length);
  }
}

/* ---------------------------------------------------------------------- norms */

void OSQPMatrix_col_norm_inf(const OSQPMatrix* M, OSQPVectorf* E) {
  /* columns of A are rows of A'.  For P the CPU reference takes the column norms of the stored
     UPPER TRIANGLE only (algebra/builtin/matrix.c:194-197 -> csc_col_norm_inf on the triu CSC),
     unlike the reference CUDA backend which uses the full symmetric matrix
     (algebra/cuda/matrix.cu:136-140).  The builtin backend is the parity oracle, so the Ruiz
     scaling (scaling.c:38) follows it. */
  if (M->is_symmetric) b200_csr_row_absmax_lower(M->S, E->d_val);
  else {
    b200_csr_row_absmax(M->St, E->d_val);
    /* row-sharded A: a column's norm is the max over the ranks' row blocks */
    if (b200_dist_world() > 1 && b200_dist_mlocal >= 0 && M->m == b200_dist_mlocal)
      b200_dist_allreduce_max(E->d_val, b200_dist_nshared >= 0 ? (int)b200_dist_nshared : (int)E->length);
  }
}

void OSQPMatrix_row_norm_inf(const OSQPMatrix* M, OSQPVectorf* E) {
  /* symmetric: full row norms (csc_row_norm_inf_sym_triu, csc_math.c:335-363) */
  b200_csr_row_absmax(M->S, E->d_val);
}

/* ------------------------------------------------------------ row extraction
 * keeps row j iff rows[j] != 0, order preserved (csc_utils.c:134-203); used by polish
 * (src/polish.c:317-372).  On the device: the kept rows of CSR(A) are compacted (flag scan + one warp
 * per row) and CSR(A_red') is the device transpose of the result; nothing is downloaded.  The host
 * filter below remains for the corner cases the device path declines (nothing kept, empty matrix,
 * a column of A_red longer than the rank-sort limit) and for B200_HOST_TRANSPOSE=1. */
static OSQPMatrix* submatrix_byrows_device(const OSQPMatrix* A, const OSQPVectori* rows) {
  int         mred = 0;
  OSQPMatrix* out;
  b200_csr*   S = b200_csr_select_rows(A->S, rows->d_val, &mred);
  if (!S) return OSQP_NULL;
  out = (OSQPMatrix*)c_calloc(1, sizeof(OSQPMatrix));
  if (!out) { b200_csr_destroy(S); return OSQP_NULL; }
  out->S  = S;
  out->St = b200_csr_transpose(S, OSQP_NULL);
  if (!out->St) { OSQPMatrix_free(out); return OSQP_NULL; }
  out->m = mred; out->n = A->n; out->nnz_user = b200_csr_nnz(S); out->is_symmetric = 0;
  return out;
}

OSQPMatrix* OSQPMatrix_submatrix_byrows(const OSQPMatrix* A, const OSQPVectori* rows) {
  OSQPInt        m = A->m, n = A->n, nnz = A->nnz_user;
  OSQPInt        i, j, k, mred = 0, nzred = 0;
  OSQPMatrix*    out   = OSQP_NULL;
  OSQPInt*       flags = OSQP_NULL;
  OSQPInt*       newrow = OSQP_NULL;
  OSQPInt *Ap = OSQP_NULL, *Ai = OSQP_NULL, *Rp = OSQP_NULL, *Ri = OSQP_NULL;
  OSQPFloat *Ax = OSQP_NULL, *Rx = OSQP_NULL;
  OSQPCscMatrix R;

  if (A->is_symmetric) {
    c_eprint("row selection not implemented for partially filled matrices");
    return OSQP_NULL;
  }
  if (!getenv("B200_HOST_TRANSPOSE") && A->m > 0) {
    out = submatrix_byrows_device(A, rows);
    if (out) return out;
  }

  flags  = (OSQPInt*)c_malloc(((size_t)m + 1) * sizeof(OSQPInt));
  newrow = (OSQPInt*)c_malloc(((size_t)m + 1) * sizeof(OSQPInt));
  Ap = (OSQPInt*)c_malloc(((size_t)n + 1) * sizeof(OSQPInt));
  Ai = (OSQPInt*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPInt));
  Ax = (OSQPFloat*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPFloat));
  Rp = (OSQPInt*)c_malloc(((size_t)n + 1) * sizeof(OSQPInt));
  Ri = (OSQPInt*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPInt));
  Rx = (OSQPFloat*)c_malloc(((size_t)nnz + 1) * sizeof(OSQPFloat));
  if (!flags || !newrow || !Ap || !Ai || !Ax || !Rp || !Ri || !Rx) goto done;

  if (m > 0) b200_copy_out(flags, rows->d_val, (size_t)m * sizeof(OSQPInt));
  /* CSR of A' == CSC of A (current, i.e. scaled, values) */
  if (b200_csr_download(A->St, Ap, Ai, Ax)) goto done;

  for (i = 0; i < m; i++) newrow[i] = flags[i] ? mred++ : -1;
  for (j = 0; j < n; j++) {
    Rp[j] = nzred;
    for (k = Ap[j]; k < Ap[j + 1]; k++) {
      if (newrow[Ai[k]] >= 0) {
        Ri[nzred] = newrow[Ai[k]];
        Rx[nzred] = Ax[k];
        nzred++;
      }
    }
  }
  Rp[n] = nzred;

  memset(&R, 0, sizeof(R));
  R.m = mred; R.n = n; R.p = Rp; R.i = Ri; R.x = Rx; R.nzmax = nzred; R.nz = -1;
  out = OSQPMatrix_new_from_csc(&R, 0);

done:
  c_free(flags); c_free(newrow); c_free(Ap); c_free(Ai); c_free(Ax);
  c_free(Rp); c_free(Ri); c_free(Rx);
  return out;
}
