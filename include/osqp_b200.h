/*
 * osqp_b200.h -- C-ABI of the B200 (sm_100a) kernel library behind OSQP's
 * `algebra/b200` linear-algebra backend.
 *
 * Everything here is `extern "C"`, plain pointers and sizes.  The plain-C backend
 * in algebra/b200/ implements OSQP's private algebra interface
 * (/root/reference/include/private/{lin_alg,algebra_vector,algebra_matrix}.h and the
 * LinSysSolver vtable, include/private/types.h:243-279) on top of these entry
 * points; every function cites the reference interface / implementation whose
 * role it takes over.  No cuSPARSE / cuBLAS / thrust is used anywhere below.
 *
 * Conventions
 *   - `b200_float` is OSQPFloat (double, or float with -DB200_USE_FLOAT), `int` is
 *     OSQPInt (32-bit).  The library is compiled once per precision.
 *   - pointers named d_* are device pointers; h_* host pointers; others are
 *     "host or device" and are classified at run time (reference:
 *     algebra/cuda/src/cuda_memory.cu:72-112).
 *   - all work is enqueued on ONE library stream; functions that return a scalar
 *     by value synchronise that stream, nothing else does.
 *   - functions returning int return 0 on success, non-zero on failure, and never
 *     abort the process (the reference aborts: algebra/cuda/include/helper_cuda.h:584-597).
 */
#ifndef OSQP_B200_H
#define OSQP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef B200_USE_FLOAT
typedef float b200_float;
#else
typedef double b200_float;
#endif

/* ------------------------------------------------------------------ lifecycle
 * replaces osqp_algebra_init_libs / free_libs / device_name
 * (algebra/cuda/algebra_libs.cu:31-75, src/cuda_handler.cu:22-44).  Ref-counted so that
 * several solvers can be alive at once (osqp_cleanup calls free_libs, osqp_api.c:1068). */
int  b200_init(int device);              /* 0 ok; 1 = no usable GPU                 */
void b200_shutdown(void);
int  b200_device_name(char* name, int len);
int  b200_sm_count(void);
void b200_sync(void);                    /* wait for the library stream             */
void* b200_stream_handle(void);          /* cudaStream_t of the library stream      */
int  b200_last_error(void);              /* sticky CUDA error code, 0 if none       */
unsigned long long b200_launch_count(void); /* kernels launched since b200_init      */
/* CUDA-graph launches of the PCG loop since b200_init: each is counted once by b200_launch_count
 * but runs 1 + 3 k kernels (k = CG iterations of that solve, see b200_pcg_stats) */
unsigned long long b200_graph_launch_count(void);
/* launch tracer (development aid): with B200_TRACE_FILE set, every launch / collective records a
 * CUDA event + host time, dumped as CSV at b200_shutdown; this adds a named marker */
void b200_trace_mark(const char* tag);
/* NVTX ranges named after the reference's profiler sections (include/private/profilers.h:16-31:
 * "linsys init", "linsys solve", "admm update", "termination check"); active with B200_NVTX=1 */
void b200_range_push(const char* name);
void b200_range_pop(void);
/* CUDA-event timing on the library stream (bench.py / profiling only) */
void* b200_event_create(void);
void  b200_event_destroy(void* ev);
void  b200_event_record(void* ev);
float b200_event_elapsed_ms(void* ev_start, void* ev_stop);   /* waits for ev_stop */

/* --------------------------------------------------------------------- memory
 * replaces cuda_malloc/cuda_calloc/cuda_free + cuda_vec_copy_{h2d,d2h,d2d}
 * (algebra/cuda/src/cuda_memory.cu, src/cuda_lin_alg.cu:520-560). */
void* b200_malloc(size_t bytes);
void* b200_calloc(size_t bytes);
void  b200_free(void* d_ptr);
int   b200_copy_in (void* d_dst, const void* src, size_t bytes);   /* src: host or device */
int   b200_copy_out(void* dst, const void* d_src, size_t bytes);   /* dst: host or device */
int   b200_ptr_is_device(const void* ptr);

/* ------------------------------------------------------------- vector kernels
 * one templated grid-stride kernel family; replaces the 19 elementwise kernels of
 * algebra/cuda/src/cuda_lin_alg.cu:38-358 and the cublas axpy/scal/copy call sites
 * (:543-677).  In-place aliasing (x == a) is allowed everywhere
 * (include/private/algebra_vector.h:146-170,202-220). */
void b200_vec_set_scalar(b200_float* d_a, b200_float sc, int n);
void b200_vec_set_scalar_cond(b200_float* d_a, const int* d_test, b200_float neg,
                              b200_float zero, b200_float pos, int n);
void b200_vec_round_to_zero(b200_float* d_a, b200_float tol, int n);
void b200_vec_mult_scalar(b200_float* d_a, b200_float sc, int n);
void b200_vec_add_scaled(b200_float* d_x, b200_float sca, const b200_float* d_a,
                         b200_float scb, const b200_float* d_b, int n);
void b200_vec_add_scaled3(b200_float* d_x, b200_float sca, const b200_float* d_a,
                          b200_float scb, const b200_float* d_b,
                          b200_float scc, const b200_float* d_c, int n);
void b200_vec_ew_prod(b200_float* d_c, const b200_float* d_a, const b200_float* d_b, int n);
void b200_vec_ew_bound(b200_float* d_x, const b200_float* d_z, const b200_float* d_l,
                       const b200_float* d_u, int n);
void b200_vec_project_polar_reccone(b200_float* d_y, const b200_float* d_l,
                                    const b200_float* d_u, b200_float infval, int n);
void b200_vec_ew_reciprocal(b200_float* d_b, const b200_float* d_a, int n);
void b200_vec_ew_sqrt(b200_float* d_a, int n);
void b200_vec_ew_max(b200_float* d_c, const b200_float* d_a, const b200_float* d_b, int n);
void b200_vec_ew_min(b200_float* d_c, const b200_float* d_a, const b200_float* d_b, int n);
void b200_vec_set_scalar_if_lt(b200_float* d_x, const b200_float* d_z, b200_float testval,
                               b200_float newval, int n);
void b200_vec_set_scalar_if_gt(b200_float* d_x, const b200_float* d_z, b200_float testval,
                               b200_float newval, int n);
void b200_vec_scatter(b200_float* d_dst, const b200_float* d_src, const int* d_idx, int n);
void b200_vec_gather(b200_float* d_dst, const b200_float* d_src, const int* d_idx, int n);

/* reductions: warp-shuffle -> one partial per CTA -> last CTA finishes in fixed order
 * (deterministic); the scalar lands in a pinned slot and is returned by value.
 * replaces cublasI?amax(+abs_kernel), cublas?dot/asum/nrm2 and the cudaMalloc'ing
 * helpers at algebra/cuda/src/cuda_lin_alg.cu:679-780,792-958. */
b200_float b200_vec_norm_inf(const b200_float* d_v, int n);
b200_float b200_vec_scaled_norm_inf(const b200_float* d_s, const b200_float* d_v, int n);
b200_float b200_vec_norm_inf_diff(const b200_float* d_a, const b200_float* d_b, int n);
b200_float b200_vec_norm_1(const b200_float* d_v, int n);
b200_float b200_vec_norm_2(const b200_float* d_v, int n);
b200_float b200_vec_dot(const b200_float* d_a, const b200_float* d_b, int n);
b200_float b200_vec_dot_signed(const b200_float* d_a, const b200_float* d_b, int sign, int n);
int b200_vec_all_leq(const b200_float* d_l, const b200_float* d_u, int n);
int b200_vec_in_reccone(const b200_float* d_y, const b200_float* d_l, const b200_float* d_u,
                        b200_float infval, b200_float tol, int n);
int b200_vec_is_eq(const b200_float* d_a, const b200_float* d_b, b200_float tol, int n);
int b200_veci_is_eq(const int* d_a, const int* d_b, int n);
/* constraint classification; returns 1 iff any entry of d_iseq changed
 * (algebra/builtin/vector.c:888-922) */
int b200_vec_bounds_type(int* d_iseq, const b200_float* d_l, const b200_float* d_u,
                         b200_float tol, b200_float infval, int n);

/* ------------------------------------------------------------------ CSR matrix
 * One device CSR matrix (int32 indices) plus its row-block schedule for the
 * "CSR-stream" SpMV: contiguous row blocks of <= B200_SPMV_TILE nonzeros are staged
 * through shared memory by one CTA; rows longer than a tile are split over several
 * CTAs and summed in fixed order by the last arriving CTA (no floating-point atomics).
 * replaces `csr` + cusparseSpMV (algebra/cuda/include/csr_type.h:27-40,
 * src/cuda_lin_alg.cu:1053-1063) and the row-wise thrust reductions (:465-498). */
typedef struct b200_csr b200_csr;

b200_csr* b200_csr_create(int nrows, int ncols, int nnz, const int* h_row_ptr,
                          const int* h_col_ind, const b200_float* h_val);
void b200_csr_destroy(b200_csr* M);
/* CSR of the transpose, built on the device from a matrix that is already there (count / scan /
 * scatter / per-row rank sort: columns ascending inside every row, bit-identical to a host
 * counting sort).  *d_map_out (optional) receives a device array of nnz ints: position in the
 * result of every stored entry of Mt; free it with b200_free.  NULL when the result would have a row
 * longer than 4096 entries (the caller keeps its host path) or on failure.
 * replaces csr_transpose (algebra/cuda/src/cuda_csr.cu:489-560: thrust sort + cusparseCsr2cscEx2) */
b200_csr* b200_csr_transpose(const b200_csr* Mt, int** d_map_out);
void b200_veci_gather(int* d_dst, const int* d_src, const int* d_idx, int n);
/* The rows of M with d_flags[row] != 0 (order preserved) as a new matrix, built on the device (flag scan,
 * row-length scan, one warp per kept row).  NULL when nothing is kept or on failure (caller keeps its
 * host path).  replaces csr_submatrix_byrows (algebra/cuda/src/cuda_csr.cu:763-843) for
 * OSQPMatrix_submatrix_byrows (polish, src/polish.c:317-372) */
b200_csr* b200_csr_select_rows(const b200_csr* M, const int* d_flags, int* nrows_out);
/* Full symmetric CSR (structurally full diagonal: lower mirrors ++ zero diagonal if missing ++ upper
 * triangle, per row) from the upper-triangular CSC arrays of P (host pointers), expanded on the
 * device.  *d_map_u / *d_map_l: device arrays of nnz ints, position of every user entry itself and
 * of its mirror (-1 for diagonal entries); free with b200_free.  NULL = caller keeps its host path.
 * replaces csr_triu_to_full / csr_expand of algebra/cuda/src/cuda_csr.cu:562-628 */
b200_csr* b200_csr_symmetric_from_triu(int n, const int* h_p, const int* h_i, const b200_float* h_x,
                                       int** d_map_u, int** d_map_l);
/* d_dst[d_idx[i]] = d_src[i] for the i with d_idx[i] >= 0 */
void b200_vec_scatter_nonneg(b200_float* d_dst, const b200_float* d_src, const int* d_idx, int n);

int  b200_csr_nrows(const b200_csr* M);
int  b200_csr_ncols(const b200_csr* M);
int  b200_csr_nnz(const b200_csr* M);
b200_float* b200_csr_values(b200_csr* M);                 /* device value array        */
int  b200_csr_download(const b200_csr* M, int* h_row_ptr, int* h_col_ind, b200_float* h_val);
/* d_y = alpha * M * d_x + beta * d_y   (beta == 0 overwrites: csc_math.c:183) */
void b200_csr_spmv(const b200_csr* M, const b200_float* d_x, b200_float* d_y,
                   b200_float alpha, b200_float beta);
void b200_csr_scale(b200_csr* M, b200_float sc);                    /* M *= sc          */
void b200_csr_scale_rows(b200_csr* M, const b200_float* d_L);       /* M = diag(L) M    */
void b200_csr_scale_cols(b200_csr* M, const b200_float* d_R);       /* M = M diag(R)    */
void b200_csr_row_absmax(const b200_csr* M, b200_float* d_out);     /* max_j |M_ij|     */
/* symmetric M stored in full: d_out[j] = max_{i<=j} |M_ij|, i.e. the column norms of the UPPER
 * TRIANGLE only -- what the CPU reference computes for P (algebra/builtin/matrix.c:194-197 calls
 * csc_col_norm_inf on the triu CSC; the symmetric variant is only used for row norms) */
void b200_csr_row_absmax_lower(const b200_csr* M, b200_float* d_out);
/* d_out[i] = sum_j M_ij^2 * w_j   (w == NULL -> w_j = w_scalar): Jacobi diagonal of A' R A
 * when M = A' (algebra/cuda/lin_sys/indirect/cuda_pcg.cu:236-251) */
void b200_csr_row_wsumsq(const b200_csr* M, const b200_float* d_w, b200_float w_scalar,
                         b200_float* d_out);
void b200_csr_diag(const b200_csr* M, b200_float* d_out);            /* 0 where absent   */
int  b200_csr_is_eq(const b200_csr* A, const b200_csr* B, b200_float tol);

/* ------------------------------------------------------ reduced-KKT PCG solver
 * Jacobi-preconditioned CG on K = P + sigma I + A' diag(rho) A, run as ONE persistent
 * cooperative kernel per ADMM iteration: reduced right-hand side, tolerance schedule,
 * the CG loop (convergence decided on device) and z~ = A x~ all happen without a single
 * host synchronisation.  replaces solve_linsys_cudapcg / cuda_pcg_alg / mat_vec_prod /
 * compute_tolerance (algebra/cuda/lin_sys/indirect/cuda_pcg_interface.cu:32-92,229-273,
 * cuda_pcg.cu:50-208): ~17 launches + 1 device sync per CG iteration there. */
typedef struct b200_pcg b200_pcg;

/* P: full symmetric n x n CSR with structurally full diagonal; A: m x n CSR; At: n x m CSR.
 * The solver borrows the matrices (sees in-place value updates after
 * b200_pcg_refresh_matrices). */
b200_pcg* b200_pcg_create(const b200_csr* P, const b200_csr* A, const b200_csr* At,
                          int n, int m);
void b200_pcg_destroy(b200_pcg* s);
/* sigma, scalar rho, optional device rho vector (NULL -> scalar), preconditioner
 * (0 none / 1 Jacobi), polishing flag */
void b200_pcg_configure(b200_pcg* s, b200_float sigma, b200_float rho,
                        const b200_float* d_rho_vec, int precond, int polishing);
/* rebuild the fused operator [P + sigma I | A'] values / the Jacobi diagonal after P, A
 * or rho changed (cuda_pcg_update_precond, cuda_pcg.cu:211-284) */
void b200_pcg_refresh_matrices(b200_pcg* s);
void b200_pcg_refresh_precond(b200_pcg* s);
void b200_pcg_warm_start(b200_pcg* s, const b200_float* d_x);
/* In place on d_b (length n+m): in = KKT right-hand side (b1, b2); out = (x~, z~ = A x~),
 * or (x, (A x - b2) / delta) when polishing.  prim_res/dual_res are the host values of
 * work->scaled_prim_res / scaled_dual_res at call time (types.h:199-200). Asynchronous. */
int  b200_pcg_solve(b200_pcg* s, b200_float* d_b, int admm_iter, double prim_res,
                    double dual_res, int max_iter, double tol_fraction,
                    int reduction_threshold);
/* synchronising statistics read-back: total CG iterations, number of solves, iterations
 * and tolerance of the last solve */
void b200_pcg_stats(b200_pcg* s, long long* total_iters, long long* n_solves,
                    int* last_iters, double* last_eps, double* last_rnorm);
/* development aid: CUDA-event timing (microseconds, averaged over `reps` plain launches) of the
 * kernels of one CG iteration of the most recently created solver (graph driver, lean passes).
 * Leaves the iterate in an arbitrary state.  Returns 0, or -1 if there is nothing to profile. */
int  b200_pcg_profile_last(int reps, double* out_us, int nout);

/* ----------------------------------------------------------- fused ADMM steps
 * One kernel each instead of the 16 launches of update_x / update_z / update_y
 * (src/auxil.c:172-229) and the 5 of compute_rhs (src/auxil.c:136-158).
 * d_rho_vec / d_rho_inv_vec may be NULL (scalar rho). */
void b200_admm_compute_rhs(b200_float* d_xtilde, b200_float* d_ztilde,
                           const b200_float* d_x_prev, const b200_float* d_q,
                           const b200_float* d_z_prev, const b200_float* d_y,
                           const b200_float* d_rho_inv_vec, b200_float rho_inv,
                           b200_float sigma, int n, int m);
void b200_admm_update_xzy(b200_float* d_x, b200_float* d_delta_x, b200_float* d_z,
                          b200_float* d_y, b200_float* d_delta_y,
                          const b200_float* d_xtilde, const b200_float* d_ztilde,
                          const b200_float* d_x_prev, const b200_float* d_z_prev,
                          const b200_float* d_l, const b200_float* d_u,
                          const b200_float* d_rho_vec, const b200_float* d_rho_inv_vec,
                          b200_float rho, b200_float rho_inv, b200_float alpha, int n, int m);
/* same, and carries A x through the relaxation step when d_Ax != NULL (d_Ax holds A x_prev on
 * entry, A x on exit: A x+ = alpha z~ + (1 - alpha) A x since z~ = A x~): the termination check
 * then needs no SpMV for A x (compute_prim_res, src/auxil.c:308-332) */
void b200_admm_update_xzy_carry(b200_float* d_x, b200_float* d_delta_x, b200_float* d_z,
                                b200_float* d_y, b200_float* d_delta_y,
                                const b200_float* d_xtilde, const b200_float* d_ztilde,
                                const b200_float* d_x_prev, const b200_float* d_z_prev,
                                const b200_float* d_l, const b200_float* d_u,
                                const b200_float* d_rho_vec, const b200_float* d_rho_inv_vec,
                                b200_float rho, b200_float rho_inv, b200_float alpha, int n, int m,
                                b200_float* d_Ax);

/* All reductions of one termination check in ONE kernel + ONE device->host copy, instead of the
 * ~13 synchronising reductions and ~7 elementwise kernels of update_info / compute_prim_res /
 * compute_dual_res / compute_obj_val_dual_gap / compute_*_tol / compute_rho_estimate
 * (src/auxil.c:14-47,231-458,676-762).  Inputs: x, y, z and the products Ax, Px, Aty (already
 * computed), q, l, u and the scaling vectors (NULL when scaling is off).  h_out[17] receives, in
 * the order of B200_RES_*: maxima are inf-norms, *_U are the Einv/Dinv-weighted ones. */
enum {
  B200_RES_PRIM_S = 0, B200_RES_PRIM_U, B200_RES_Z_S, B200_RES_Z_U, B200_RES_AX_S, B200_RES_AX_U,
  B200_RES_SC, B200_RES_DUAL_S, B200_RES_DUAL_U, B200_RES_Q_S, B200_RES_Q_U, B200_RES_PX_S,
  B200_RES_PX_U, B200_RES_ATY_S, B200_RES_ATY_U, B200_RES_XPX, B200_RES_QX, B200_RES_COUNT
};
void b200_admm_residuals(const b200_float* d_x, const b200_float* d_y, const b200_float* d_z,
                         const b200_float* d_Ax, const b200_float* d_Px, const b200_float* d_Aty,
                         const b200_float* d_q, const b200_float* d_l, const b200_float* d_u,
                         const b200_float* d_Einv, const b200_float* d_Dinv, b200_float infval,
                         b200_float deadzone, int n, int m, double* h_out);
/* The five scalars that decide whether is_primal_infeasible / is_dual_infeasible (src/auxil.c:460-585)
 * can fire at all -- { ||E.*dy||_inf, u'max(dy,0), l'min(dy,0), ||D.*dx||_inf, q'dx }, dy projected on the
 * polar of the recession cone in place first -- in one kernel and one host round trip instead of five
 * blocking reductions per termination check.  d_E / d_D may be NULL. */
void b200_admm_infeas_scalars(b200_float* d_dy, const b200_float* d_l, const b200_float* d_u,
                              const b200_float* d_E, const b200_float* d_dx, const b200_float* d_D,
                              const b200_float* d_q, b200_float infval, int n, int m, int do_primal,
                              int do_dual, double* h_out);

/* ------------------------------------------------------ batches of small QPs (BASELINE configs[4])
 * nb independent QPs that share P and A (hence the scaling D, E, c of the set-up template problem) and
 * differ in their bounds (and optionally q): ONE CTA per QP runs the whole osqp_solve loop -- ADMM
 * steps, reduced-KKT PCG with the reference's tolerance schedule, update_info, check_termination,
 * adaptive rho -- in shared memory (osqp_b200/csrc/batch.cu cites the reference lines).  Matrices and
 * scaling vectors are the SCALED device data of the template; d_l_batch / d_u_batch / d_q_batch hold the
 * users' unscaled values, d_x / d_y receive unscaled solutions.  Status per QP: 1 solved, 2 solved
 * inaccurate, 7 maximum iterations reached, 9 non-convex (infeasibility certificates are not evaluated
 * here: unsolved QPs go through the ordinary API).  Returns 0, or 2 (m == 0) / 3 (iterates do not fit
 * shared memory): use the ordinary API.  Asynchronous on the library stream.
 * replaces: one osqp_solve per QP = ~100 launches per ADMM iteration through the per-op interface
 * (include/private/algebra_vector.h:28-290; docs/examples/mpc.rst:30-90 for the workload). */
typedef struct {
  b200_float rho, sigma, alpha, eps_abs, eps_rel, adaptive_rho_tolerance;
  int rho_is_vec, max_iter, check_termination, adaptive_rho, adaptive_rho_interval, check_dualgap,
      scaled_termination, cg_max_iter, cg_tol_reduction;
  double cg_tol_fraction;
} b200_batch_settings;
int b200_batch_solve(const b200_csr* P, const b200_csr* A, const b200_csr* At, int n, int m, int nb,
                     const b200_float* d_q, const b200_float* d_q_batch, const b200_float* d_l_batch,
                     const b200_float* d_u_batch, const b200_float* d_D, const b200_float* d_Dinv,
                     const b200_float* d_E, const b200_float* d_Einv, b200_float c, b200_float cinv,
                     const b200_batch_settings* st, b200_float* d_x, b200_float* d_y, int* d_iters, int* d_status,
                     int* d_cg_iters, int* d_rho_updates, b200_float* d_obj, b200_float* d_prim_res,
                     b200_float* d_dual_res);

/* ------------------------------------------------------ row-sharded multi-GPU mode
 * One process per GPU.  Every rank holds a block of ROWS of A (and the matching slices of the
 * m-vectors); n-vectors and P are replicated.  The only data-path exchange is one all-reduce of
 * a length-n vector per K.p (and per A'y).  New functionality: the reference has no multi-GPU
 * path.  NCCL is loaded lazily; the 128-byte unique id is created on rank
 * 0 and distributed by the host program (torch.distributed in bench.py / tests). */
int  b200_dist_unique_id(unsigned char* id128);
int  b200_dist_init(int rank, int world, const unsigned char* id128);   /* after b200_init */
void b200_dist_finalize(void);
int  b200_dist_world(void);
int  b200_dist_rank(void);
/* while suspended the process behaves as a single-GPU process (world 1): lets one rank of a sharded job
 * solve a whole problem on its own GPU, e.g. the single-GPU leg of a strong-scaling measurement */
void b200_dist_suspend(int on);
/* while the scope is 1, the scalar-returning reductions combine their result across ranks (max or
 * sum as appropriate) before handing it to the host: set by the backend around reductions over
 * row-sharded vectors */
void b200_dist_scope(int sharded);
/* column-split layout (SURVEY.md 8e "recommended refinement"): every n-vector of this rank is
 * [n_shared columns touched by several ranks, replicated ; the columns this rank owns].  Only the
 * shared slice is ever exchanged.  n_shared < 0 switches the layout off (plain row sharding with
 * fully replicated n-vectors). */
void b200_dist_set_split(int n_shared);
int  b200_dist_n_shared(void);
void b200_dist_allreduce_sum(b200_float* d_buf, int n);   /* in place, library stream */
void b200_dist_allreduce_max(b200_float* d_buf, int n);
void b200_dist_stats(unsigned long long* n_calls, unsigned long long* bytes);
/* Peer-memory exchange for the CG loop of the row-sharded solve (NVLink P2P stores instead of NCCL
 * calls).  After b200_dist_init: every rank exports the 64-byte CUDA IPC handle
 * of its exchange buffer, the host program all-gathers the handles (rank order) and every rank imports
 * them.  Without these two calls the sharded solve keeps its NCCL path. */
int  b200_dist_p2p_export(unsigned char* handle64);
int  b200_dist_p2p_import(const unsigned char* handles, int world);   /* world x 64 bytes */
int  b200_dist_p2p_enabled(void);
int  b200_dist_p2p_error(void);     /* 1: an exchange timed out (a peer never arrived); synchronises */

/* bumped by every kernel launch / device copy of the library: lets the backend cache scalars and
 * know when they went stale */
unsigned long long b200_epoch(void);

#ifdef __cplusplus
}
#endif

#endif /* OSQP_B200_H */
