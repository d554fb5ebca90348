// pcg_graph.cu -- reduced-KKT PCG for LARGE problems: lean one-CTA-per-tile kernels with the CG
// loop expressed as a CUDA-graph WHILE node whose condition is written on the device.
//
// Why a second driver: the SpMV passes are latency/gather bound (csr.cuh), so they want many small
// resident CTAs.  The persistent cooperative kernel (pcg.cu) needs 64 registers/thread for the
// union of its phases and is pinned to 4 CTAs/SM with a static round-robin of tiles; the same pass
// as a stand-alone 32-40 register kernel with one CTA per tile runs 15-25 % faster (measured,
// tools/micro/spmv_variants.cu).  A conditional graph node keeps what made the persistent kernel
// attractive: the convergence test `||r||_inf > eps && it < max_iter` is evaluated on the device
// (cudaGraphSetConditional in the last kernel of the loop body), so a solve is still enqueued
// without a single host synchronisation, whatever the iteration count turns out to be.
//
// Same algorithm, same tolerance schedule, same carried A x as pcg.cu (see its header).
// Grid-wide scalars: every CTA publishes a partial, the LAST CTA to arrive (integer ticket) folds
// them in index order -> deterministic.
#include "pcg.cuh"

#include <cstring>
#include <vector>

using namespace b200;

namespace {

// fold `part` (already CTA-reduced, valid in thread 0) into slot `slot`; returns true in ALL
// threads of the last CTA to arrive, with the folded total in `total`.
template <bool IS_MAX>
__device__ __forceinline__ bool publish(double part, double* red, int stride, int slot, unsigned* ticket,
                                        bool count, double* shr, double& total) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    red[slot * stride + blockIdx.x] = part;
    if (count) {
      __threadfence();
      const unsigned t = atomicAdd(ticket, 1u);
      s_last = (t == gridDim.x - 1);
    }
  }
  if (!count) return false;
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double a = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    const double v = __ldcg(red + slot * stride + i);
    a = IS_MAX ? fmax(a, v) : a + v;
  }
  total = IS_MAX ? block_max(a, shr) : block_sum(a, shr);
  return true;
}
// fold a second slot inside the last CTA (no ticket traffic)
template <bool IS_MAX>
__device__ __forceinline__ double fold(const double* red, int stride, int slot, double* shr) {
  double a = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    const double v = __ldcg(red + slot * stride + i);
    a = IS_MAX ? fmax(a, v) : a + v;
  }
  return IS_MAX ? block_max(a, shr) : block_sum(a, shr);
}

__global__ void g_set_args(PcgArgs* dst, PcgArgs a) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *dst = a;
}

// t = rho .* b2   (only for the ||rhs|| of the tolerance at admm_iter == 1 / polishing)
__global__ void __launch_bounds__(kBlock) g_rhs_t(const PcgArgs* ap) {
  const PcgArgs& a = *ap;
  const T* b2 = a.b + a.n;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.m; j += stride)
    a.t[j] = (a.rho_vec ? a.rho_vec[j] : a.rho) * b2[j];
}

// ||b1 + A' t||_inf  -> run->rhs_norm
__global__ void __launch_bounds__(kSpmvBlock, B200_SPMV_MINBLOCKS) g_rhs_norm(const PcgArgs* ap, PcgRun* run, double* red,
                                                         int stride) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double shr[33];
  const PcgArgs& a = *ap;
  Pipe pipe = pipe_init(dsm);
  const T* b1 = a.b;
  const T* t = a.t;
  const CsrView M = a.At;   // local copy: the argument block lives in global memory
  double mx = 0.0;
  if (a.m > 0) {
    spmv_pass<SumOp>(
        M, blockIdx.x, gridDim.x, pipe, [&](int, int c, T v) { return v * t[c]; },
        [&](int row, T s) { mx = fmax(mx, fabs((double)(b1[row] + s))); });
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x)
      mx = fmax(mx, fabs((double)b1[i]));
  }
  mx = block_max(mx, shr);
  double tot;
  if (publish<true>(mx, red, stride, SLOT_RHS, &run->ticket[SLOT_RHS], true, shr, tot) && threadIdx.x == 0) {
    run->rhs_norm = tot;
    run->ticket[SLOT_RHS] = 0;
  }
}

// tolerance schedule of compute_tolerance (cuda_pcg_interface.cu:32-64), evaluated on the device
__global__ void g_tolerance(const PcgArgs* ap, PcgRun* run) {
  if (threadIdx.x || blockIdx.x) return;
  const PcgArgs& a = *ap;
  const PcgState st = *a.st;
  double rf = st.reduction_factor, eps_prev = st.eps_prev, eps;
  int zero_iters = st.zero_iters;
  if (a.polishing) {
    eps = fmax(run->rhs_norm * kCgPolishTol, kCgTolMin);
  } else if (a.admm_iter == 1) {
    rf       = a.tol_fraction;
    eps_prev = (run->rhs_norm < kCgTolMin) ? 1.0 : run->rhs_norm * rf;
    eps      = eps_prev;
  } else {
    if (zero_iters >= a.reduction_threshold) {
      rf *= 0.5;
      zero_iters = 0;
    }
    eps      = rf * sqrt(a.prim_res * a.dual_res);
    eps      = fmax(fmin(eps, eps_prev), kCgTolMin);
    eps_prev = eps;
  }
  run->eps = eps; run->rf = rf; run->eps_prev = eps_prev; run->zero_iters = zero_iters;
  run->it = 0;
}

// P1 from the carried product: t = rho .* (Ax - b2)
__global__ void __launch_bounds__(kBlock) g_p1_carried(const PcgArgs* ap) {
  const PcgArgs& a = *ap;
  const T* b2 = a.b + a.n;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.m; j += stride)
    a.t[j] = (a.rho_vec ? a.rho_vec[j] : a.rho) * (a.Ax[j] - b2[j]);
}

// pass over A.  MODE 0: P1 with exact recomputation (Ax = A x ; t = rho .* (Ax - b2))
//               MODE 1: L1 (w = A p ; t = rho .* w)
template <int MODE>
__global__ void __launch_bounds__(kSpmvBlock, B200_SPMV_MINBLOCKS) g_pass_A(const PcgArgs* ap) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const PcgArgs& a = *ap;
  Pipe pipe = pipe_init(dsm);
  const T* src = (MODE == 0) ? a.x : a.p;
  const T* b2 = a.b + a.n;
  const T* rho_vec = a.rho_vec;
  const T rho = a.rho;
  T* t = a.t;
  T* out = (MODE == 0) ? a.Ax : a.w;
  const CsrView M = a.A;    // local copy: the argument block lives in global memory
  spmv_pass<SumOp>(
      M, blockIdx.x, gridDim.x, pipe, [&](int, int c, T v) { return v * src[c]; },
      [&](int row, T s) {
        out[row] = s;
        const T rr = rho_vec ? rho_vec[row] : rho;
        t[row] = (MODE == 0) ? rr * (s - b2[row]) : rr * s;
      });
}

// pass over the fused operator [P + sigma I | A'].
//   MODE 0: P2  r = K2 [x; t] - b1 ; p = -M^-1 r ; totals r'y, ||r||_inf -> run
//   MODE 1: L2  Kp = K2 [p; t] ; total p'Kp -> run
//   MODE 2/3 (row-sharded): Kp = this rank's PARTIAL K2 [x; t] / K2 [p; t]; the all-reduce and
//           the scalars follow in separate kernels (g_resid_init / g_dot_pKp)
template <int MODE>
__global__ void __launch_bounds__(kSpmvBlock, B200_SPMV_MINBLOCKS) g_pass_K(const PcgArgs* ap, PcgRun* run, double* red,
                                                       int stride) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double shr[33];
  const PcgArgs& a = *ap;
  Pipe pipe = pipe_init(dsm);
  const int n = a.n;
  const T* src = (MODE == 0 || MODE == 2) ? a.x : a.p;
  const T* t = a.t;
  const T* b1 = a.b;
  const T* minv = a.minv;
  T* r = a.r; T* p = a.p; T* Kp = a.Kp;
  double acc0 = 0.0, acc1 = 0.0;
  const CsrView M = a.K2;   // local copy: the argument block lives in global memory
  spmv_pass<SumOp>(
      M, blockIdx.x, gridDim.x, pipe,
      [&](int, int c, T v) { return v * (c < n ? src[c] : t[c - n]); },
      [&](int row, T s) {
        if (MODE == 0) {
          const T rr = s - b1[row];
          const T yy = minv[row] * rr;
          r[row] = rr;
          p[row] = -yy;
          acc0 += (double)rr * (double)yy;
          acc1 = fmax(acc1, fabs((double)rr));
        } else if (MODE == 1) {
          Kp[row] = s;
          acc0 += (double)p[row] * (double)s;
        } else {
          Kp[row] = s;
        }
      });
  if (MODE >= 2) return;
  double tot;
  if (MODE == 0) {
    acc0 = block_sum(acc0, shr);
    acc1 = block_max(acc1, shr);
    publish<true>(acc1, red, stride, SLOT_RMAX, nullptr, false, shr, tot);
    if (publish<false>(acc0, red, stride, SLOT_RTY, &run->ticket[SLOT_RTY], true, shr, tot)) {
      const double rmax = fold<true>(red, stride, SLOT_RMAX, shr);
      if (threadIdx.x == 0) {
        run->rTy = tot;
        run->rnorm = rmax;
        run->ticket[SLOT_RTY] = 0;
      }
    }
  } else {
    acc0 = block_sum(acc0, shr);
    if (publish<false>(acc0, red, stride, SLOT_PKP, &run->ticket[SLOT_PKP], true, shr, tot) &&
        threadIdx.x == 0) {
      run->pKp = tot;
      run->ticket[SLOT_PKP] = 0;
    }
  }
}

// row-sharded: Kp[i] = (A_r' t)_i partial, for the ||rhs|| of the tolerance
__global__ void __launch_bounds__(kSpmvBlock, B200_SPMV_MINBLOCKS) g_pass_At(const PcgArgs* ap) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const PcgArgs& a = *ap;
  Pipe pipe = pipe_init(dsm);
  const T* t = a.t;
  T* Kp = a.Kp;
  const CsrView M = a.At;
  spmv_pass<SumOp>(
      M, blockIdx.x, gridDim.x, pipe, [&](int, int c, T v) { return v * t[c]; },
      [&](int row, T s) { Kp[row] = s; });
}

// row-sharded: ||b1 + Kp||_inf -> run->rhs_norm   (Kp = all-reduced A'(rho .* b2))
__global__ void __launch_bounds__(kBlock) g_rhs_norm_sum(const PcgArgs* ap, PcgRun* run, double* red,
                                                         int stride, int have_At) {
  __shared__ double shr[33];
  const PcgArgs& a = *ap;
  double mx = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x)
    mx = fmax(mx, fabs((double)(a.b[i] + (have_At ? a.Kp[i] : (T)0))));
  mx = block_max(mx, shr);
  double tot;
  if (publish<true>(mx, red, stride, SLOT_RHS, &run->ticket[SLOT_RHS], true, shr, tot) && threadIdx.x == 0) {
    run->rhs_norm = tot;
    run->ticket[SLOT_RHS] = 0;
  }
}

// row-sharded P2 tail: r = Kp - b1 ; p = -M^-1 r ; totals r'y, ||r||_inf (n-vectors are replicated,
// so every rank computes the same totals and no scalar exchange is needed)
__global__ void __launch_bounds__(kBlock) g_resid_init(const PcgArgs* ap, PcgRun* run, double* red,
                                                       int stride) {
  __shared__ double shr[33];
  const PcgArgs& a = *ap;
  double acc0 = 0.0, acc1 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const T rr = a.Kp[i] - a.b[i];
    const T yy = a.minv[i] * rr;
    a.r[i] = rr;
    a.p[i] = -yy;
    acc0 += (double)rr * (double)yy;
    acc1 = fmax(acc1, fabs((double)rr));
  }
  acc0 = block_sum(acc0, shr);
  acc1 = block_max(acc1, shr);
  double tot;
  publish<true>(acc1, red, stride, SLOT_RMAX, nullptr, false, shr, tot);
  if (publish<false>(acc0, red, stride, SLOT_RTY, &run->ticket[SLOT_RTY], true, shr, tot)) {
    const double rmax = fold<true>(red, stride, SLOT_RMAX, shr);
    if (threadIdx.x == 0) {
      run->rTy = tot;
      run->rnorm = rmax;
      run->ticket[SLOT_RTY] = 0;
    }
  }
}

// row-sharded L2 tail: p'Kp of the all-reduced Kp
__global__ void __launch_bounds__(kBlock) g_dot_pKp(const PcgArgs* ap, PcgRun* run, double* red, int stride) {
  __shared__ double shr[33];
  const PcgArgs& a = *ap;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x)
    acc += (double)a.p[i] * (double)a.Kp[i];
  acc = block_sum(acc, shr);
  double tot;
  if (publish<false>(acc, red, stride, SLOT_PKP, &run->ticket[SLOT_PKP], true, shr, tot) && threadIdx.x == 0) {
    run->pKp = tot;
    run->ticket[SLOT_PKP] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Lean passes for the CG loop body (matrices without over-long rows): the same CSR-stream tile as
// spmv_pass, written out flat -- one CTA of 512 threads per 2048-entry tile, a single batch of 4
// (col, val, gather) chains per thread, no long-row branch, no generic functors -- so that ptxas
// keeps it at 32 registers (full occupancy).  Measured on the Lasso matrix: 58 us per pass against
// 73-80 us for the generic instantiation (tools/micro/spmv_variants.cu).
//   MODE 0: L1   w = A p ; t = rho .* w
//   MODE 1: L2   Kp = [P + sigma I | A'] [p; t] ; total p'Kp -> run
constexpr int kLeanBlock = 512;
static_assert(kTile == 4 * kLeanBlock, "lean pass assumes one batch of 4 per thread");

template <int MODE>
__global__ void __launch_bounds__(kLeanBlock, 3) g_lean_pass(const PcgArgs* ap, PcgRun* run, double* red,
                                                             int stride) {
  __shared__ T sm[kTile];
  __shared__ int srp[kMaxRows + 1];
  __shared__ double shr[33];
  const PcgArgs& a = *ap;
  // everything the tile needs is pulled out of the (global-memory) argument block once
  const int* __restrict__ row_ptr = (MODE == 0) ? a.A.row_ptr : a.K2.row_ptr;
  const int* __restrict__ col_ind = (MODE == 0) ? a.A.col_ind : a.K2.col_ind;
  const T* __restrict__   val     = (MODE == 0) ? a.A.val : a.K2.val;
  const int4* __restrict__ desc   = (MODE == 0) ? a.A.desc : a.K2.desc;
  const T* __restrict__ p = a.p;
  const T* __restrict__ t = a.t;
  const T* __restrict__ rho_vec = a.rho_vec;
  const T rho = a.rho;
  T* __restrict__ out0 = (MODE == 0) ? a.w : a.Kp;
  T* __restrict__ out1 = a.t;
  const int n = a.n;
  const int tid = threadIdx.x;
  const int4 d = __ldg(desc + blockIdx.x);
  const int nnz0 = d.z, cnt = d.w, nrows = d.y & 0xffffff, lg = d.y >> 24;
  for (int i = tid; i <= nrows; i += kLeanBlock) srp[i] = ld_stream(row_ptr + d.x + i) - nnz0;
  int c[4];
  T   v[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int k = u * kLeanBlock + tid;
    if (k < cnt) {
      c[u] = ld_stream(col_ind + nnz0 + k);
      v[u] = ld_stream(val + nnz0 + k);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int k = u * kLeanBlock + tid;
    if (k < cnt) {
      if (MODE == 0) sm[k] = v[u] * p[c[u]];
      else sm[k] = v[u] * (c[u] < n ? p[c[u]] : t[c[u] - n]);
    }
  }
  __syncthreads();
  const int g = 1 << lg, gid = tid >> lg, lig = tid & (g - 1), ngroup = kLeanBlock >> lg;
  double acc = 0.0;
  for (int base = 0; base < nrows; base += ngroup) {   // one trip except for tiles of 1-2 entry rows
    const int r = base + gid;
    T sum = 0;
    if (r < nrows) {
      const int e = srp[r + 1];
      for (int k = srp[r] + lig; k < e; k += g) sum += sm[k];
    }
    for (int o = g >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (r < nrows && lig == 0) {
      const int row = d.x + r;
      if (MODE == 0) {
        out0[row] = sum;
        out1[row] = (rho_vec ? rho_vec[row] : rho) * sum;
      } else {
        out0[row] = sum;
        acc += (double)p[row] * (double)sum;
      }
    }
  }
  if (MODE == 1) {
    acc = block_sum(acc, shr);
    double tot;
    if (publish<false>(acc, red, stride, SLOT_PKP, &run->ticket[SLOT_PKP], true, shr, tot) &&
        threadIdx.x == 0) {
      run->pKp = tot;
      run->ticket[SLOT_PKP] = 0;
    }
  }
}

// the flat kernel does not handle chunks of over-long rows
static bool lean_ok(const b200_csr& M, const std::vector<int4>& desc) {
  if (M.nlong > 0) return false;
  for (const int4& d : desc)
    if (d.y < 0) return false;
  return true;
}

// first node of the loop graph: arm the WHILE condition from the initial residual
__global__ void g_loop_init(const PcgArgs* ap, PcgRun* run, cudaGraphConditionalHandle h) {
  if (threadIdx.x || blockIdx.x) return;
  cudaGraphSetConditional(h, (run->rnorm > run->eps && run->it < ap->max_iter) ? 1u : 0u);
}

// L3: x += a p ; r += a Kp ; Ax += a w ; totals r'y, ||r||_inf ; last CTA: beta, it++, condition
__global__ void __launch_bounds__(kBlock) g_update(const PcgArgs* ap, PcgRun* run, double* red, int stride,
                                                   cudaGraphConditionalHandle h) {
  __shared__ double shr[33];
  const PcgArgs& a = *ap;
  const int n = a.n, m = a.m;
  const T alpha = (T)(run->rTy / run->pKp);
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  double acc_rty = 0.0, acc_max = 0.0;
  for (int i = gtid; i < n; i += gstride) {
    a.x[i] += alpha * a.p[i];
    const T rr = a.r[i] + alpha * a.Kp[i];
    a.r[i] = rr;
    const T yy = a.minv[i] * rr;
    acc_rty += (double)rr * (double)yy;
    acc_max = fmax(acc_max, fabs((double)rr));
  }
  for (int j = gtid; j < m; j += gstride) a.Ax[j] += alpha * a.w[j];
  acc_rty = block_sum(acc_rty, shr);
  acc_max = block_max(acc_max, shr);
  double tot;
  publish<true>(acc_max, red, stride, SLOT_RMAX, nullptr, false, shr, tot);
  if (publish<false>(acc_rty, red, stride, SLOT_RTY, &run->ticket[SLOT_RTY], true, shr, tot)) {
    const double rmax = fold<true>(red, stride, SLOT_RMAX, shr);
    if (threadIdx.x == 0) {
      run->beta  = tot / run->rTy;
      run->rTy   = tot;
      run->rnorm = rmax;
      run->it   += 1;
      run->ticket[SLOT_RTY] = 0;
      if (h) cudaGraphSetConditional(h, (rmax > run->eps && run->it < a.max_iter) ? 1u : 0u);
    }
  }
}

// L4: p = beta p - M^-1 r
__global__ void __launch_bounds__(kBlock) g_direction(const PcgArgs* ap, const PcgRun* run) {
  const PcgArgs& a = *ap;
  const T beta = (T)run->beta;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
    a.p[i] = beta * a.p[i] - a.minv[i] * a.r[i];
}

// E1: b1 = x ; b2 = A x (carried) or (A x - b2)/delta when polishing ; persist the schedule state
__global__ void __launch_bounds__(kBlock) g_epilogue(const PcgArgs* ap, PcgRun* run) {
  const PcgArgs& a = *ap;
  T* b1 = a.b;
  T* b2 = a.b + a.n;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  for (int i = gtid; i < a.n; i += gstride) b1[i] = a.x[i];
  const bool pol = a.polishing != 0;
  for (int j = gtid; j < a.m; j += gstride) b2[j] = pol ? a.rho * (a.Ax[j] - b2[j]) : a.Ax[j];
  if (gtid == 0) {
    PcgState o = *a.st;
    o.reduction_factor = run->rf;
    o.eps_prev         = run->eps_prev;
    o.zero_iters       = (run->it == 0) ? run->zero_iters + 1 : 0;
    o.last_iters       = run->it;
    o.last_eps         = run->eps;
    o.last_rnorm       = run->rnorm;
    o.total_iters     += run->it;
    o.n_solves        += 1;
    *a.st = o;
  }
}

inline int pass_grid(const b200_csr& M, int cap) {
  int g = M.nblocks < cap ? M.nblocks : cap;
  return g > 0 ? g : 1;
}

}  // namespace

void b200_pcg_graph_configure_kernels() {
  b200_enable_spmv_smem(g_rhs_norm);
  b200_enable_spmv_smem(g_pass_A<0>);
  b200_enable_spmv_smem(g_pass_A<1>);
  b200_enable_spmv_smem(g_pass_K<0>);
  b200_enable_spmv_smem(g_pass_K<1>);
  b200_enable_spmv_smem(g_pass_K<2>);
  b200_enable_spmv_smem(g_pass_K<3>);
  b200_enable_spmv_smem(g_pass_At);
}

int b200_pcg_graph_build(b200_pcg* s) {
  Context& c = ctx();
  bool ok = true;
  s->gred_stride = c.sm_count * 32;
  if (s->K2.nblocks + 8 > s->gred_stride) s->gred_stride = s->K2.nblocks + 8;
  if (s->A && s->A->nblocks + 8 > s->gred_stride) s->gred_stride = s->A->nblocks + 8;
  ok &= B200_CHECK(dev_malloc(&s->d_args, sizeof(PcgArgs)));
  ok &= B200_CHECK(dev_malloc(&s->d_run, sizeof(PcgRun)));
  ok &= B200_CHECK(dev_malloc(&s->d_gred, sizeof(double) * SLOT_COUNT * s->gred_stride));
  if (!ok) return 1;
  B200_CHECK(cudaMemsetAsync(s->d_run, 0, sizeof(PcgRun), c.stream));
  if (s->sharded) return 0;   // host-driven loop with an all-reduce per iteration: no graph

  cudaGraph_t g = nullptr;
  if (!B200_CHECK(cudaGraphCreate(&g, 0))) return 1;
  cudaGraphConditionalHandle h;
  if (!B200_CHECK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault))) return 1;

  const PcgArgs* d_args = s->d_args;
  PcgRun* d_run = s->d_run;
  double* d_red = s->d_gred;
  int stride = s->gred_stride;

  // node 0: arm the condition
  cudaGraphNode_t n_init;
  {
    void* args[] = {(void*)&d_args, (void*)&d_run, (void*)&h};
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = (void*)g_loop_init; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.kernelParams = args;
    ok &= B200_CHECK(cudaGraphAddKernelNode(&n_init, g, nullptr, 0, &kp));
  }
  // node 1: WHILE
  cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
  cp.conditional.handle = h;
  cp.conditional.type   = cudaGraphCondTypeWhile;
  cp.conditional.size   = 1;
  cudaGraphNode_t n_while;
  ok &= B200_CHECK(cudaGraphAddNode(&n_while, g, &n_init, 1, &cp));
  if (!ok) return 1;
  cudaGraph_t body = cp.conditional.phGraph_out[0];

  // body: L1 -> L2 -> L3 -> L4
  cudaGraphNode_t prev = nullptr;
  auto add = [&](void* func, dim3 grid, dim3 block, size_t smem, void** args) {
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = func; kp.gridDim = grid; kp.blockDim = block; kp.sharedMemBytes = (unsigned)smem;
    kp.kernelParams = args;
    cudaGraphNode_t node;
    ok &= B200_CHECK(cudaGraphAddKernelNode(&node, body, prev ? &prev : nullptr, prev ? 1 : 0, &kp));
    prev = node;
  };
  const int cap = s->gred_stride;
  // lean flat kernels when neither matrix has over-long rows (checked on the host schedules)
  bool lean = getenv("B200_PCG_NO_LEAN") == nullptr;
  if (lean) {
    auto fetch = [&](const b200_csr& M) {
      std::vector<int4> h(M.nblocks > 0 ? M.nblocks : 0);
      if (M.nblocks > 0) {
        B200_CHECK(cudaMemcpyAsync(h.data(), M.d_desc, sizeof(int4) * M.nblocks, cudaMemcpyDeviceToHost, c.stream));
        B200_CHECK(cudaStreamSynchronize(c.stream));
      }
      return h;
    };
    lean = lean_ok(s->K2, fetch(s->K2)) && (s->m == 0 || lean_ok(*s->A, fetch(*s->A))) &&
           s->K2.nblocks <= cap && (s->m == 0 || s->A->nblocks <= cap);
  }
  s->lean = lean ? 1 : 0;
  if (getenv("B200_TRACE_SETUP"))
    fprintf(stderr, "[b200 trace] graph PCG driver: %s passes, K2 tiles %d, A tiles %d, partial stride %d\n",
            lean ? "lean" : "generic", s->K2.nblocks, s->m > 0 ? s->A->nblocks : 0, cap);
  if (s->m > 0) {
    if (lean) {
      void* a1[] = {(void*)&d_args, (void*)&d_run, (void*)&d_red, (void*)&stride};
      add((void*)g_lean_pass<0>, dim3(s->A->nblocks), dim3(kLeanBlock), 0, a1);
    } else {
      void* a1[] = {(void*)&d_args};
      add((void*)g_pass_A<1>, dim3(pass_grid(*s->A, cap)), dim3(kSpmvBlock), kSpmvSmemBytes, a1);
    }
  }
  {
    void* a2[] = {(void*)&d_args, (void*)&d_run, (void*)&d_red, (void*)&stride};
    if (lean) add((void*)g_lean_pass<1>, dim3(s->K2.nblocks), dim3(kLeanBlock), 0, a2);
    else add((void*)g_pass_K<1>, dim3(pass_grid(s->K2, cap)), dim3(kSpmvBlock), kSpmvSmemBytes, a2);
  }
  {
    void* a3[] = {(void*)&d_args, (void*)&d_run, (void*)&d_red, (void*)&stride, (void*)&h};
    int nm = s->n > s->m ? s->n : s->m;
    add((void*)g_update, dim3(ew_grid(nm)), dim3(kBlock), 0, a3);
  }
  {
    void* a4[] = {(void*)&d_args, (void*)&d_run};
    add((void*)g_direction, dim3(ew_grid(s->n)), dim3(kBlock), 0, a4);
  }
  if (!ok) return 1;
  cudaGraphExec_t exec = nullptr;
  if (!B200_CHECK(cudaGraphInstantiate(&exec, g, 0))) return 1;
  s->graph = (void*)g;
  s->graph_exec = (void*)exec;
  return 0;
}

void b200_pcg_graph_destroy(b200_pcg* s) {
  if (s->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)s->graph_exec);
  if (s->graph) cudaGraphDestroy((cudaGraph_t)s->graph);
  dev_free(s->d_args);
  dev_free(s->d_run);
  dev_free(s->d_gred);
  s->graph_exec = s->graph = nullptr;
  s->d_args = nullptr; s->d_run = nullptr; s->d_gred = nullptr;
}

int b200_pcg_graph_solve(b200_pcg* s, const PcgArgs& a) {
  Context& c = ctx();
  cudaStream_t st = c.stream;
  const int cap = s->gred_stride;
  const int n = s->n, m = s->m;
  g_set_args<<<1, 32, 0, st>>>(s->d_args, a);
  count_launch();
  const PcgArgs* d_args = s->d_args;
  if (a.polishing || a.admm_iter == 1) {
    if (m > 0) {
      g_rhs_t<<<ew_grid(m), kBlock, 0, st>>>(d_args);
      count_launch();
    }
    const int g = m > 0 ? pass_grid(*s->At, cap) : (ew_grid(n) < cap ? ew_grid(n) : cap);
    g_rhs_norm<<<g, kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, s->d_run, s->d_gred, cap);
    count_launch();
  }
  g_tolerance<<<1, 32, 0, st>>>(d_args, s->d_run);
  count_launch();
  if (m > 0) {
    if (a.ax_valid) g_p1_carried<<<ew_grid(m), kBlock, 0, st>>>(d_args);
    else g_pass_A<0><<<pass_grid(*s->A, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
    count_launch();
  }
  g_pass_K<0><<<pass_grid(s->K2, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, s->d_run, s->d_gred, cap);
  count_launch();
  bool ok = B200_CHECK(cudaGraphLaunch((cudaGraphExec_t)s->graph_exec, st));
  count_launch();
  const int nm = n > m ? n : m;
  g_epilogue<<<ew_grid(nm), kBlock, 0, st>>>(d_args, s->d_run);
  count_launch();
  return ok ? 0 : 1;
}

// ------------------------------------------------------------------ row-sharded driver
// Every rank holds a row block A_r (CSR), A_r' and the matching slices of the m-vectors; x, p, r,
// Kp, M^-1 are replicated.  K p = (P + sigma I) p [rank 0 only] + A_r' (rho .* (A_r p)) summed by ONE
// NCCL all-reduce of the length-n partial per CG iteration; all CG scalars are then computed
// redundantly from replicated vectors, so they are bit-identical on every rank and the ranks take
// the same branch without any scalar exchange.  The loop is driven by the host, which reads
// (||r||_inf, eps, it) back once per iteration.
int b200_pcg_sharded_solve(b200_pcg* s, const PcgArgs& a) {
  Context& c = ctx();
  cudaStream_t st = c.stream;
  const int cap = s->gred_stride;
  const int n = s->n, m = s->m;
  const int gn = ew_grid(n) < cap ? ew_grid(n) : cap;
  g_set_args<<<1, 32, 0, st>>>(s->d_args, a);
  count_launch();
  const PcgArgs* d_args = s->d_args;
  PcgRun* run = s->d_run;
  cudaGraphConditionalHandle none = 0;

  if (a.polishing || a.admm_iter == 1) {
    if (m > 0) {
      g_rhs_t<<<ew_grid(m), kBlock, 0, st>>>(d_args);
      count_launch();
      g_pass_At<<<pass_grid(*s->At, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
      count_launch();
    } else {
      B200_CHECK(cudaMemsetAsync(s->d_Kp, 0, sizeof(T) * n, st));
    }
    b200_dist_allreduce_sum(s->d_Kp, n);
    g_rhs_norm_sum<<<gn, kBlock, 0, st>>>(d_args, run, s->d_gred, cap, 1);
    count_launch();
  }
  g_tolerance<<<1, 32, 0, st>>>(d_args, run);
  count_launch();
  if (m > 0) {
    if (a.ax_valid) g_p1_carried<<<ew_grid(m), kBlock, 0, st>>>(d_args);
    else g_pass_A<0><<<pass_grid(*s->A, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
    count_launch();
  }
  g_pass_K<2><<<pass_grid(s->K2, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, run, s->d_gred, cap);
  count_launch();
  b200_dist_allreduce_sum(s->d_Kp, n);
  g_resid_init<<<gn, kBlock, 0, st>>>(d_args, run, s->d_gred, cap);
  count_launch();

  PcgRun h;
  bool ok = true;
  for (;;) {
    ok &= B200_CHECK(cudaMemcpyAsync(&h, run, sizeof(PcgRun), cudaMemcpyDeviceToHost, st));
    ok &= B200_CHECK(cudaStreamSynchronize(st));
    if (!ok || !(h.rnorm > h.eps && h.it < a.max_iter)) break;
    if (m > 0) {
      g_pass_A<1><<<pass_grid(*s->A, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
      count_launch();
    }
    g_pass_K<3><<<pass_grid(s->K2, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, run, s->d_gred, cap);
    count_launch();
    b200_dist_allreduce_sum(s->d_Kp, n);
    g_dot_pKp<<<gn, kBlock, 0, st>>>(d_args, run, s->d_gred, cap);
    count_launch();
    const int nm = n > m ? n : m;
    g_update<<<ew_grid(nm) < cap ? ew_grid(nm) : cap, kBlock, 0, st>>>(d_args, run, s->d_gred, cap, none);
    count_launch();
    g_direction<<<ew_grid(n), kBlock, 0, st>>>(d_args, run);
    count_launch();
  }
  const int nm = n > m ? n : m;
  g_epilogue<<<ew_grid(nm), kBlock, 0, st>>>(d_args, run);
  count_launch();
  return ok ? 0 : 1;
}
