// pcg.cu -- reduced-KKT Jacobi-preconditioned conjugate gradient as ONE persistent
// cooperative kernel per ADMM iteration.
//
// Solves  K x = b1 + A' (rho .* b2),   K = P + sigma I + A' diag(rho) A
// warm-started from the previous solution, to  ||K x - rhs||_inf <= eps  with the reference's
// tolerance schedule, then returns (x~, z~ = A x~) in place of (b1, b2).
//
// Role taken over from the reference CUDA backend:
//   solve_linsys_cudapcg  algebra/cuda/lin_sys/indirect/cuda_pcg_interface.cu:229-273
//   compute_tolerance     cuda_pcg_interface.cu:32-64   (now evaluated on the device)
//   compute_rhs           cuda_pcg_interface.cu:67-92
//   cuda_pcg_alg          cuda_pcg.cu:113-208            (17 launches + 1 sync per iteration)
//   mat_vec_prod          cuda_pcg.cu:50-106             (3 cusparseSpMV + copy + scal)
//   update_precond        cuda_pcg.cu:211-284
//
// Kernel plan (grid = co-resident CTAs, phases separated by grid.sync()):
//   [R]  only when admm_iter == 1 or polishing: ||b1 + A'(rho.*b2)||_inf for the tolerance
//   P1   t = rho .* (A x - b2)          elementwise from the carried A x (a pass over A only when
//                                       the carried product is stale: first solve, after warm_start /
//                                       matrix updates, and every kAxResync solves)
//   P2   r = [P+sigma I | A'] [x; t] - b1 ; p = -M^-1 r ; partials r'y, ||r||_inf
//   loop while ||r||_inf > eps and it < max_iter:
//     L1 w = A p ; t = rho .* w                       SpMV over A
//     L2 Kp = [P+sigma I | A'] [p; t] ; partial p'Kp  SpMV over the fused operator (n rows)
//     L3 x += a p ; r += a Kp ; Ax += a w ; partials r'y, ||r||_inf   (y = M^-1 r in registers)
//     L4 p  = beta p - M^-1 r
//   E1   b1 = x ; b2 = A x (carried; or (A x - b2)/delta when polishing)
// Carrying A x through the CG recurrence (A x_{k+1} = A x_k + a A p_k, and A p_k is computed in
// L1 anyway) removes two of the 2k+3 matrix passes of a solve; the product is recomputed
// exactly every kAxResync solves so rounding drift stays at the 1e-13 level.
// Every grid-wide scalar is a fixed-order sum of per-CTA partials -> deterministic for a
// given grid; no floating-point atomics; zero host synchronisation.
//
// Algorithmic HBM bytes (F = sizeof(T)), SpMV(r x c, nnz) = nnz (F+4) + (r+1) 4 + c F + r F:
//   per CG iteration : SpMV(A) + SpMV([P|A']) + 8 n F      (K.p + 8nF)
//   per launch, fixed: SpMV([P|A']) + (3n + 3m) F          (initial residual + rhs/write-back)
#include "pcg.cuh"

#include <cooperative_groups.h>
#include <vector>
#include <cstring>
#include <cstdlib>

namespace cg = cooperative_groups;
using namespace b200;

namespace {



__device__ __forceinline__ double grid_sum(const double* slot, int G, double* shr) {
  double a = 0.0;
  for (int i = threadIdx.x; i < G; i += kSpmvBlock) a += __ldcg(slot + i);
  return block_sum(a, shr);
}
__device__ __forceinline__ double grid_max(const double* slot, int G, double* shr) {
  double a = 0.0;
  for (int i = threadIdx.x; i < G; i += kSpmvBlock) a = fmax(a, __ldcg(slot + i));
  return block_max(a, shr);
}

__global__ void __launch_bounds__(kSpmvBlock, B200_PCG_MINBLOCKS) pcg_kernel(PcgArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double shr[33];
  Pipe pipe = pipe_init(dsm);

  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, m = a.m;
  const int gtid = cta * kSpmvBlock + tid, gstride = G * kSpmvBlock;
  T* const b1 = a.b;
  T* const b2 = a.b + n;
  T* const x = a.x; T* const p = a.p; T* const Kp = a.Kp; T* const r = a.r; T* const t = a.t;
  T* const Ax = a.Ax; T* const w = a.w;
  const T* const minv = a.minv;
  const T* const rho_vec = a.rho_vec;
  const T rho = a.rho;
  double* const red = a.red;

  // ------------------------------------------------------------- tolerance
  const PcgState st = *a.st;    // only rewritten by CTA 0 after the last grid.sync()
  double rf = st.reduction_factor, eps_prev = st.eps_prev, eps;
  int zero_iters = st.zero_iters;
  if (a.polishing || a.admm_iter == 1) {
    // [R] ||rhs||_inf, rhs = b1 + A' (rho .* b2)
    if (m > 0) {
      for (int j = gtid; j < m; j += gstride) t[j] = (rho_vec ? rho_vec[j] : rho) * b2[j];
      grid.sync();
    }
    double mx = 0.0;
    if (m > 0) {
      spmv_pass<SumOp>(
          a.At, cta, G, pipe, [&](int, int c, T v) { return v * t[c]; },
            [&](int row, T s) { mx = fmax(mx, fabs((double)(b1[row] + s))); });
    } else {
      for (int i = gtid; i < n; i += gstride) mx = fmax(mx, fabs((double)b1[i]));
    }
    mx = block_max(mx, shr);
    if (tid == 0) red[SLOT_RHS * G + cta] = mx;
    grid.sync();
    const double rhs_norm = grid_max(red + SLOT_RHS * G, G, shr);
    if (a.polishing) {
      eps = fmax(rhs_norm * kCgPolishTol, kCgTolMin);
    } else {
      rf       = a.tol_fraction;
      eps_prev = (rhs_norm < kCgTolMin) ? 1.0 : rhs_norm * rf;
      eps      = eps_prev;
    }
  } else {
    if (zero_iters >= a.reduction_threshold) {
      rf *= 0.5;
      zero_iters = 0;
    }
    eps      = rf * sqrt(a.prim_res * a.dual_res);
    eps      = fmax(fmin(eps, eps_prev), kCgTolMin);
    eps_prev = eps;
  }

  // ------------------------------------------------------------- P1: t = rho.*(A x - b2)
  if (m > 0) {
    if (a.ax_valid) {
      for (int j = gtid; j < m; j += gstride) t[j] = (rho_vec ? rho_vec[j] : rho) * (Ax[j] - b2[j]);
    } else {
      spmv_pass<SumOp>(
          a.A, cta, G, pipe, [&](int, int c, T v) { return v * x[c]; },
          [&](int row, T s) {
            Ax[row] = s;
            t[row]  = (rho_vec ? rho_vec[row] : rho) * (s - b2[row]);
          });
    }
    grid.sync();
  }

  // ------------------------------------------------------------- P2: initial residual
  double acc_rty = 0.0, acc_max = 0.0;
  spmv_pass<SumOp>(
          a.K2, cta, G, pipe, [&](int, int c, T v) { return v * (c < n ? x[c] : t[c - n]); },
        [&](int row, T s) {
          const T rr = s - b1[row];
          const T yy = minv[row] * rr;
          r[row] = rr;
          p[row] = -yy;
          acc_rty += (double)rr * (double)yy;
          acc_max = fmax(acc_max, fabs((double)rr));
        });
  acc_rty = block_sum(acc_rty, shr);
  acc_max = block_max(acc_max, shr);
  if (tid == 0) {
    red[SLOT_RTY * G + cta]  = acc_rty;
    red[SLOT_RMAX * G + cta] = acc_max;
  }
  grid.sync();
  double rTy   = grid_sum(red + SLOT_RTY * G, G, shr);
  double rnorm = grid_max(red + SLOT_RMAX * G, G, shr);

  // ------------------------------------------------------------- CG loop
  int it = 0;
  while (rnorm > eps && it < a.max_iter) {
    // L1: t = rho .* (A p)
    if (m > 0) {
      spmv_pass<SumOp>(
          a.A, cta, G, pipe, [&](int, int c, T v) { return v * p[c]; },
          [&](int row, T s) {
            w[row] = s;
            t[row] = (rho_vec ? rho_vec[row] : rho) * s;
          });
      grid.sync();
    }
    // L2: Kp = [P + sigma I | A'] [p; t], partial p'Kp
    double acc = 0.0;
    spmv_pass<SumOp>(
          a.K2, cta, G, pipe, [&](int, int c, T v) { return v * (c < n ? p[c] : t[c - n]); },
          [&](int row, T s) {
            Kp[row] = s;
            acc += (double)p[row] * (double)s;
          });
    acc = block_sum(acc, shr);
    if (tid == 0) red[SLOT_PKP * G + cta] = acc;
    grid.sync();
    const double pKp   = grid_sum(red + SLOT_PKP * G, G, shr);
    const T      alpha = (T)(rTy / pKp);

    // L3: x += alpha p ; r += alpha Kp ; y = M^-1 r (registers) ; partials
    acc_rty = 0.0; acc_max = 0.0;
    for (int i = gtid; i < n; i += gstride) {
      x[i] += alpha * p[i];
      const T rr = r[i] + alpha * Kp[i];
      r[i] = rr;
      const T yy = minv[i] * rr;
      acc_rty += (double)rr * (double)yy;
      acc_max = fmax(acc_max, fabs((double)rr));
    }
    for (int j = gtid; j < m; j += gstride) Ax[j] += alpha * w[j];
    acc_rty = block_sum(acc_rty, shr);
    acc_max = block_max(acc_max, shr);
    if (tid == 0) {
      red[SLOT_RTY * G + cta]  = acc_rty;
      red[SLOT_RMAX * G + cta] = acc_max;
    }
    grid.sync();
    const double rTy_new = grid_sum(red + SLOT_RTY * G, G, shr);
    rnorm                = grid_max(red + SLOT_RMAX * G, G, shr);
    const T beta         = (T)(rTy_new / rTy);
    rTy                  = rTy_new;

    // L4: p = beta p - y
    for (int i = gtid; i < n; i += gstride) p[i] = beta * p[i] - minv[i] * r[i];
    grid.sync();
    it++;
  }

  // ------------------------------------------------------------- E1: write back
  for (int i = gtid; i < n; i += gstride) b1[i] = x[i];
  if (m > 0) {
    const bool pol = a.polishing != 0;
    for (int j = gtid; j < m; j += gstride) b2[j] = pol ? rho * (Ax[j] - b2[j]) : Ax[j];
  }
  if (cta == 0 && tid == 0) {
    PcgState o = st;
    o.reduction_factor = rf;
    o.eps_prev         = eps_prev;
    o.zero_iters       = (it == 0) ? zero_iters + 1 : 0;
    o.last_iters       = it;
    o.last_eps         = eps;
    o.last_rnorm       = rnorm;
    o.total_iters      = st.total_iters + it;
    o.n_solves         = st.n_solves + 1;
    *a.st = o;
  }
}

// K2 row pointer: rp[i] = rpP[i] + rpAt[i]
__global__ void k2_rowptr_kernel(int n, const int* rpP, const int* rpAt, int* rp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) rp[i] = rpP[i] + (rpAt ? rpAt[i] : 0);
}

// one warp per row: copy P row (sigma added on the diagonal) then A' row (columns shifted by n)
__global__ void __launch_bounds__(kBlock) k2_fill_kernel(int n, T sigma, T pscale_head, int n_head, const int* rpP, const int* ciP,
                                                         const T* vP, const int* rpAt, const int* ciAt,
                                                         const T* vAt, const int* rp, int* ci, T* v) {
  const int lane = threadIdx.x & 31;
  const int wpb  = kBlock >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < n; row += gridDim.x * wpb) {
    int dst = rp[row];
    // rows < n_head (all rows under plain row sharding, the SHARED columns under the column-split
    // layout) carry P + sigma I on one rank only; rows owned by this rank always carry it
    const T pscale = row < n_head ? pscale_head : (T)1;
    const int s0 = rpP[row], e0 = rpP[row + 1];
    for (int k = s0 + lane; k < e0; k += 32) {
      const int c = ciP[k];
      ci[dst + k - s0] = c;
      v[dst + k - s0]  = pscale * (vP[k] + (c == row ? sigma : (T)0));
    }
    dst += e0 - s0;
    if (rpAt) {
      const int s1 = rpAt[row], e1 = rpAt[row + 1];
      for (int k = s1 + lane; k < e1; k += 32) {
        ci[dst + k - s1] = ciAt[k] + n;
        v[dst + k - s1]  = vAt[k];
      }
    }
  }
}

__global__ void precond_kernel(int n, T sigma, const T* pd, const T* ad, T* minv, int use) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    minv[i] = use ? (T)1.0 / (sigma + pd[i] + ad[i]) : (T)1.0;
}

}  // namespace

void b200_pcg_configure_kernels() {
  b200_enable_spmv_smem(pcg_kernel);
  b200_pcg_graph_configure_kernels();
}

extern "C" {

b200_pcg* b200_pcg_create(const b200_csr* P, const b200_csr* A, const b200_csr* At, int n, int m) {
  Context& c = ctx();
  b200_pcg* s = new b200_pcg();
  s->P = P; s->A = A; s->At = At; s->n = n; s->m = m;
  s->owner_slot = c.slot;
  bool ok = true;
  auto alloc = [&](T** p, size_t cnt) { ok &= B200_CHECK(dev_malloc(p, sizeof(T) * (cnt + 1))); };
  alloc(&s->d_x, n); alloc(&s->d_p, n); alloc(&s->d_Kp, n); alloc(&s->d_r, n);
  alloc(&s->d_t, m); alloc(&s->d_minv, n); alloc(&s->d_pd, n); alloc(&s->d_ad, n);
  alloc(&s->d_Ax, m); alloc(&s->d_w, m);
  ok &= B200_CHECK(dev_malloc(&s->d_state, sizeof(PcgState)));
  if (!ok) { b200_pcg_destroy(s); return nullptr; }
  B200_CHECK(cudaMemsetAsync(s->d_x, 0, sizeof(T) * (n + 1), c.stream));   // PCG iterate starts at 0
  B200_CHECK(cudaMemsetAsync(s->d_state, 0, sizeof(PcgState), c.stream));
  B200_CHECK(cudaMemsetAsync(s->d_ad, 0, sizeof(T) * (n + 1), c.stream));

  // fused operator K2 = [P + sigma I | A'] : pattern + schedule now, values in refresh_matrices
  std::vector<int> rpP(n + 1), rpAt(n + 1, 0), rp(n + 1);
  ok &= B200_CHECK(cudaMemcpyAsync(rpP.data(), P->d_row_ptr, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, c.stream));
  if (m > 0)
    ok &= B200_CHECK(cudaMemcpyAsync(rpAt.data(), At->d_row_ptr, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, c.stream));
  ok &= B200_CHECK(cudaStreamSynchronize(c.stream));
  for (int i = 0; i <= n; i++) rp[i] = rpP[i] + rpAt[i];
  b200_csr& K = s->K2;
  K.nrows = n; K.ncols = n + m; K.nnz = rp[n];
  ok &= B200_CHECK(dev_malloc(&K.d_row_ptr, sizeof(int) * ((size_t)n + 2 * kPad)));
  ok &= B200_CHECK(dev_malloc(&K.d_col_ind, sizeof(int) * ((size_t)K.nnz + 2 * kPad)));
  ok &= B200_CHECK(dev_malloc(&K.d_val, sizeof(T) * ((size_t)K.nnz + 2 * kPad)));
  if (!ok) { b200_pcg_destroy(s); return nullptr; }
  ok &= B200_CHECK(cudaMemcpyAsync(K.d_row_ptr, rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, c.stream));
  ok &= B200_CHECK(cudaStreamSynchronize(c.stream));
  if (!ok || b200_build_schedule(&K, rp.data()) != 0) { b200_pcg_destroy(s); return nullptr; }

  // cooperative grid: all CTAs must be co-resident
  int per_sm = 0;
  ok &= B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_kernel, kSpmvBlock,
                                                                 kSpmvSmemBytes));
  if (per_sm < 1) per_sm = 1;
  s->max_grid = per_sm * c.sm_count;
  long long want = K.nblocks;
  if (A && A->nblocks > want) want = A->nblocks;
  long long ew = ((long long)n + kSpmvBlock * 2 - 1) / (kSpmvBlock * 2);
  if (ew > want) want = ew;
  if (want < 1) want = 1;
  s->grid = (int)(want < s->max_grid ? want : s->max_grid);
  ok &= B200_CHECK(dev_malloc(&s->d_red, sizeof(double) * SLOT_COUNT * s->max_grid));
  if (!ok) { b200_pcg_destroy(s); return nullptr; }
  // driver choice: the graph driver (lean one-wave / one-CTA-per-tile kernels, WHILE node) costs ~6
  // launches per solve but runs its passes at full occupancy: measured on the 1.14e7-nnz Lasso
  // 113 ms against 121 ms per ADMM solve for the persistent kernel (profiles/r01_phase_profile.md),
  // so it is the default from kGraphDriverMinNnz stored entries; B200_PCG_DRIVER=graph|persistent
  // overrides.  It is also the structure a row-sharded solve needs: kernel boundaries where the
  // all-reduce goes.
  const char* env = getenv("B200_PCG_DRIVER");
  const long long work = (long long)K.nnz + (A ? (long long)A->nnz : 0);
  s->use_graph = env ? (strcmp(env, "graph") == 0) : (work >= kGraphDriverMinNnz);
  if (dist_active()) {
    s->sharded   = 1;
    s->use_graph = 1;                       // shares the lean kernels and device structs
    s->include_P = (b200_dist_rank() == 0);
  }
  if (s->use_graph && b200_pcg_graph_build(s) != 0) {
    fprintf(stderr, "[osqp_b200] graph PCG driver unavailable, using the persistent kernel\n");
    b200_pcg_graph_destroy(s);
    s->use_graph = 0;
  }
  b200_pcg_profile_register(s, true);
  return s;
}

void b200_pcg_destroy(b200_pcg* s) {
  if (!s) return;
  b200_pcg_profile_register(s, false);
  dev_free(s->d_x); dev_free(s->d_p); dev_free(s->d_Kp); dev_free(s->d_r); dev_free(s->d_t);
  dev_free(s->d_Ax); dev_free(s->d_w);
  dev_free(s->d_minv); dev_free(s->d_pd); dev_free(s->d_ad);
  b200_pcg_graph_destroy(s);
  dev_free(s->d_state); dev_free(s->d_red);
  dev_free(s->K2.d_row_ptr); dev_free(s->K2.d_col_ind); dev_free(s->K2.d_val);
  dev_free(s->K2.d_desc); dev_free(s->K2.d_long); dev_free(s->K2.d_long_partials);
  dev_free(s->K2.d_long_counters);
  delete s;
}

void b200_pcg_configure(b200_pcg* s, T sigma, T rho, const T* d_rho_vec, int precond, int polishing) {
  s->sigma = sigma; s->rho = rho; s->d_rho_vec = d_rho_vec;
  s->precond = precond; s->polishing = polishing;
}

void b200_pcg_refresh_matrices(b200_pcg* s) {
  Context& c = ctx();
  const int n = s->n;
  s->ax_valid = 0;   // A may have changed
  if (n <= 0) return;
  const bool hasA = s->m > 0;
  int grid = (n + (kBlock >> 5) - 1) / (kBlock >> 5);
  int cap  = c.sm_count * 8;
  if (grid > cap) grid = cap;
  k2_fill_kernel<<<grid, kBlock, 0, c.stream>>>(
      n, s->sigma, (T)(s->include_P ? 1 : 0), (s->sharded && dist_split()) ? dist_n_shared() : n, s->P->d_row_ptr, s->P->d_col_ind, s->P->d_val, hasA ? s->At->d_row_ptr : nullptr,
      hasA ? s->At->d_col_ind : nullptr, hasA ? s->At->d_val : nullptr, s->K2.d_row_ptr,
      s->K2.d_col_ind, s->K2.d_val);
  count_launch();
  b200_csr_diag(s->P, s->d_pd);
}

void b200_pcg_refresh_precond(b200_pcg* s) {
  Context& c = ctx();
  const int n = s->n;
  if (n <= 0) return;
  if (s->m > 0 && s->precond) b200_csr_row_wsumsq(s->At, s->d_rho_vec, s->rho, s->d_ad);
  // row-sharded: diag(A' R A) = sum over ranks of the local column sums
  if (s->sharded && s->precond) {
    if (s->m <= 0) B200_CHECK(cudaMemsetAsync(s->d_ad, 0, sizeof(T) * n, c.stream));
    b200_dist_allreduce_sum(s->d_ad, dist_split() ? dist_n_shared() : n);
  }
  precond_kernel<<<ew_grid(n), kBlock, 0, c.stream>>>(n, s->sigma, s->d_pd, s->d_ad, s->d_minv, s->precond);
  count_launch();
}

void b200_pcg_warm_start(b200_pcg* s, const T* d_x) {
  ctx().epoch++;
  s->ax_valid = 0;   // the iterate is replaced: the carried A x is stale
  if (s->n > 0)
    B200_CHECK(cudaMemcpyAsync(s->d_x, d_x, sizeof(T) * s->n, cudaMemcpyDeviceToDevice, ctx().stream));
}

int b200_pcg_solve(b200_pcg* s, T* d_b, int admm_iter, double prim_res, double dual_res, int max_iter,
                   double tol_fraction, int reduction_threshold) {
  if (s->n <= 0) return 0;
  if (ctx().refcount <= 0 || ctx().slot != s->owner_slot) {
    // the context (stream, constant argument block, mailbox) is per host thread: a solve issued from
    // another thread would run on the wrong stream against the wrong argument block
    fprintf(stderr, "[osqp_b200] solver used from a thread other than the one that created it\n");
    return 1;
  }
  PcgArgs a;
  memset(&a, 0, sizeof(a));
  a.K2 = s->K2.view();
  if (s->m > 0) { a.A = s->A->view(); a.At = s->At->view(); }   // m == 0: no A phases at all
  a.n = s->n; a.m = s->m;
  a.n_shared = (s->sharded && dist_split()) ? dist_n_shared() : 0;
  a.x = s->d_x; a.p = s->d_p; a.Kp = s->d_Kp; a.r = s->d_r; a.t = s->d_t; a.b = d_b;
  a.Ax = s->d_Ax; a.w = s->d_w;
  // carried A x: recomputed exactly when stale, when polishing, and every kAxResync solves
  if (s->polishing || s->solves_since_sync >= kAxResync) s->ax_valid = 0;
  a.ax_valid = s->ax_valid;
  if (!s->ax_valid) s->solves_since_sync = 0;
  s->solves_since_sync++;
  s->ax_valid = 1;
  a.minv = s->d_minv; a.rho_vec = s->d_rho_vec; a.rho = s->rho;
  a.admm_iter = admm_iter; a.max_iter = max_iter; a.polishing = s->polishing;
  a.reduction_threshold = reduction_threshold;
  a.prim_res = prim_res; a.dual_res = dual_res; a.tol_fraction = tol_fraction;
  a.st = s->d_state; a.red = s->d_red;
  if (s->sharded) return b200_pcg_sharded_solve(s, a);
  if (s->use_graph) return b200_pcg_graph_solve(s, a);
  void* args[] = {&a};
  bool ok = B200_CHECK(cudaLaunchCooperativeKernel((const void*)pcg_kernel, dim3(s->grid), dim3(kSpmvBlock),
                                                   args, kSpmvSmemBytes, ctx().stream));
  count_launch();
  return ok ? 0 : 1;
}

void b200_pcg_stats(b200_pcg* s, long long* total_iters, long long* n_solves, int* last_iters,
                    double* last_eps, double* last_rnorm) {
  PcgState h;
  memset(&h, 0, sizeof(h));
  B200_CHECK(cudaMemcpyAsync(&h, s->d_state, sizeof(h), cudaMemcpyDeviceToHost, ctx().stream));
  B200_CHECK(cudaStreamSynchronize(ctx().stream));
  if (total_iters) *total_iters = h.total_iters;
  if (n_solves) *n_solves = h.n_solves;
  if (last_iters) *last_iters = h.last_iters;
  if (last_eps) *last_eps = h.last_eps;
  if (last_rnorm) *last_rnorm = h.last_rnorm;
}

}  // extern "C"
