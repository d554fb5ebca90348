/*
 * ORACLE / TEST INFRASTRUCTURE -- not part of the shipped B200 product path.
 *
 * Restatement of the sparse quasi-definite LDL' factorisation the reference's
 * CPU ("builtin") backend delegates to QDLDL v0.1.8 (un-vendored dependency,
 * /root/reference/algebra/_common/lin_sys/qdldl/qdldl.cmake:8-10).  Written from
 * the published algorithm (elimination tree + up-looking row-by-row LDL', see
 * T. Davis, "Direct Methods for Sparse Linear Systems", ch. 4, and the OSQP paper
 * section 5) and from the calling contract visible in the reference:
 *   etree : qdldl_interface.c:94-107      factor : qdldl_interface.c:116-130
 *   solve : qdldl_interface.c:394-416
 *
 * Storage: A = L D L' with A given as its upper triangle in CSC (diagonal entry
 * last in every column), L unit lower triangular stored by columns without the
 * unit diagonal, D diagonal.
 *
 * PARITY: unpinned against upstream QDLDL bits (source unavailable offline);
 * pinned against scipy splu and the reference's end-to-end golden solutions.
 */
#include "qdldl.h"

#define NODE_NONE   (-1)
#define MARK_CLEAR  (0)
#define MARK_SET    (1)

/*
 * Elimination tree of the upper-triangular CSC pattern (Ap, Ai) and the column
 * counts of L.  For every entry (i, j), i < j, walk from i towards the root
 * until a node already visited for column j is met; every node passed gains one
 * entry in its column of L (row j), and a parent-less node is attached to j.
 */
QDLDL_int QDLDL_etree(const QDLDL_int  n,
                      const QDLDL_int* Ap,
                      const QDLDL_int* Ai,
                      QDLDL_int*       work,
                      QDLDL_int*       Lnz,
                      QDLDL_int*       etree) {
  QDLDL_int col, k, node, total;

  for (col = 0; col < n; col++) {
    work[col]  = 0;
    Lnz[col]   = 0;
    etree[col] = NODE_NONE;
    /* a structurally empty column has no diagonal: not factorisable */
    if (Ap[col] == Ap[col + 1]) return -1;
  }

  for (col = 0; col < n; col++) {
    work[col] = col;
    for (k = Ap[col]; k < Ap[col + 1]; k++) {
      node = Ai[k];
      if (node > col) return -1;   /* entry below the diagonal */
      while (work[node] != col) {
        if (etree[node] == NODE_NONE) etree[node] = col;
        Lnz[node]++;
        work[node] = col;
        node = etree[node];
      }
    }
  }

  total = 0;
  for (col = 0; col < n; col++) {
    if (total > QDLDL_INT_MAX - Lnz[col]) return -2;
    total += Lnz[col];
  }
  return total;
}

/*
 * Numerical factorisation, one row of L at a time (up-looking).  Row k of L
 * solves the triangular system  L(0:k,0:k) D(0:k) y = A(0:k,k); its pattern is
 * the union of the etree paths started at the entries of column k of A, which
 * is enumerated in topological order before the sparse triangular solve.
 *
 * iwork is split into three n-slices (pattern of y, a path buffer, the next
 * free slot of every column of L); bwork marks pattern membership; fwork holds
 * the dense accumulator y.
 */
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
                       QDLDL_float*       fwork) {
  QDLDL_int    row, k, q, node, c, slot, npat, npath;
  QDLDL_int    n_pos = 0;
  QDLDL_int*   pattern   = iwork;
  QDLDL_int*   path      = iwork + n;
  QDLDL_int*   col_fill  = iwork + 2 * n;
  QDLDL_bool*  in_pat    = bwork;
  QDLDL_float* y         = fwork;
  QDLDL_float  y_c, l_kc;

  Lp[0] = 0;
  for (k = 0; k < n; k++) {
    Lp[k + 1]   = Lp[k] + Lnz[k];
    in_pat[k]   = MARK_CLEAR;
    y[k]        = 0.0;
    D[k]        = 0.0;
    col_fill[k] = Lp[k];
  }

  /* first pivot: the (0,0) entry is the only entry of column 0 */
  D[0] = Ax[0];
  if (D[0] == 0.0) return -1;
  if (D[0] > 0.0) n_pos++;
  Dinv[0] = 1.0 / D[0];

  for (row = 1; row < n; row++) {
    npat = 0;

    /* scatter column `row` of A into y and build the pattern of L(row,:) */
    for (k = Ap[row]; k < Ap[row + 1]; k++) {
      node = Ai[k];
      if (node == row) {
        D[row] = Ax[k];
        continue;
      }
      y[node] = Ax[k];

      if (in_pat[node] == MARK_CLEAR) {
        /* climb the tree until a marked node or the current row is reached */
        in_pat[node] = MARK_SET;
        path[0] = node;
        npath   = 1;
        q = etree[node];
        while (q != NODE_NONE && q < row) {
          if (in_pat[q] == MARK_SET) break;
          in_pat[q]     = MARK_SET;
          path[npath++] = q;
          q = etree[q];
        }
        /* append the path reversed so that, read backwards, the full pattern
           is in topological (ascending-dependency) order */
        while (npath) pattern[npat++] = path[--npath];
      }
    }

    /* sparse triangular solve over the pattern, last-pushed first */
    for (k = npat - 1; k >= 0; k--) {
      c    = pattern[k];
      slot = col_fill[c];
      y_c  = y[c];

      for (q = Lp[c]; q < slot; q++) y[Li[q]] -= Lx[q] * y_c;

      l_kc     = y_c * Dinv[c];
      Li[slot] = row;
      Lx[slot] = l_kc;
      D[row]  -= y_c * l_kc;
      col_fill[c] = slot + 1;

      y[c]      = 0.0;
      in_pat[c] = MARK_CLEAR;
    }

    if (D[row] == 0.0) return -1;
    if (D[row] > 0.0) n_pos++;
    Dinv[row] = 1.0 / D[row];
  }

  return n_pos;
}

/* x <- L^{-1} x  (unit lower triangular, columns) */
void QDLDL_Lsolve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li,
                  const QDLDL_float* Lx, QDLDL_float* x) {
  QDLDL_int col, k;
  for (col = 0; col < n; col++) {
    QDLDL_float xc = x[col];
    for (k = Lp[col]; k < Lp[col + 1]; k++) x[Li[k]] -= Lx[k] * xc;
  }
}

/* x <- L^{-T} x */
void QDLDL_Ltsolve(const QDLDL_int n, const QDLDL_int* Lp, const QDLDL_int* Li,
                   const QDLDL_float* Lx, QDLDL_float* x) {
  QDLDL_int col, k;
  for (col = n - 1; col >= 0; col--) {
    QDLDL_float xc = x[col];
    for (k = Lp[col]; k < Lp[col + 1]; k++) xc -= Lx[k] * x[Li[k]];
    x[col] = xc;
  }
}

/* x <- (L D L')^{-1} x */
void QDLDL_solve(const QDLDL_int    n,
                 const QDLDL_int*   Lp,
                 const QDLDL_int*   Li,
                 const QDLDL_float* Lx,
                 const QDLDL_float* Dinv,
                 QDLDL_float*       x) {
  QDLDL_int k;
  QDLDL_Lsolve(n, Lp, Li, Lx, x);
  for (k = 0; k < n; k++) x[k] *= Dinv[k];
  QDLDL_Ltsolve(n, Lp, Li, Lx, x);
}
