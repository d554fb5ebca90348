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
#include "xchg.cuh"

#include <cstring>
#include <cstdlib>
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

// sums of three values over the CTA with one pair of barriers; results valid in thread 0
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* sh /* >= 3*16 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  if (lane == 0) { sh[w] = a; sh[16 + w] = b; sh[32 + w] = c; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    a = warp_sum(lane < nw ? sh[lane] : 0.0);
    b = warp_sum(lane < nw ? sh[16 + lane] : 0.0);
    c = warp_sum(lane < nw ? sh[32 + lane] : 0.0);
  }
}
__device__ __forceinline__ void block_sum_max(double& a, double& mx, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_sum(a); mx = warp_max(mx);
  if (lane == 0) { sh[w] = a; sh[16 + w] = mx; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    a  = warp_sum(lane < nw ? sh[lane] : 0.0);
    mx = warp_max(lane < nw ? sh[16 + lane] : 0.0);
  }
}

// thread 0 publishes up to three partials of this CTA; the last CTA to arrive folds every slot in
// index order.  Returns true in all threads of that CTA, totals valid in all its threads.
template <int NSLOT, bool LAST_IS_MAX>
__device__ __forceinline__ bool publish_fold(const double* part, const int* slots, double* red, int stride,
                                             unsigned* ticket, double* shr, double* totals) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NSLOT; i++) red[slots[i] * stride + blockIdx.x] = part[i];
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int i = 0; i < NSLOT; i++) {
    const bool is_max = LAST_IS_MAX && (i == NSLOT - 1);
    double acc = 0.0;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) {
      const double v = __ldcg(red + slots[i] * stride + j);
      acc = is_max ? fmax(acc, v) : acc + v;
    }
    totals[i] = is_max ? block_max(acc, shr) : block_sum(acc, shr);
  }
  return true;
}

// The argument block of the current solve lives in CONSTANT memory: every kernel of the driver
// (the graph's kernel nodes have frozen parameters) reads pointers and scalars as constant-bank
// operands -- no registers, no L1 traffic.  (Reading them through a pointer to global memory cost
// one extra L1 wavefront per use in the per-row epilogues: 17 us of a 70 us pass.)
__constant__ PcgArgs c_args[kMaxContexts];   // one block per library context (host thread)

// ---- peer-memory exchange (row-sharded solve; buffers and protocol: common.cuh / dist.cu) -------------
__constant__ XchgView c_xchg;
// ALL threads of ONE CTA: this rank's (a, b) goes to every peer, the peers' pairs come back; returns
// the sum of the a's and the maximum of the b's, folded in rank order (bit-identical on all ranks)
__device__ __forceinline__ void xchg_scalars(double a_loc, double b_loc, double& a_sum, double& b_max) {
  const XchgView& X = c_xchg;
  XchgState* S = X.state;
  const int me = X.rank, world = X.world, tid = threadIdx.x;
  const unsigned long long seq = *(volatile unsigned long long*)&S->sseq + 1;
  const int set = (int)(seq & 1ull);
  if (tid < world && tid != me) {
    double* dst = X.peer[tid] + xchg_sc(set, me);
    dst[0] = a_loc;
    dst[1] = b_loc;
    __threadfence_system();
    st_release_sys((unsigned long long*)(X.peer[tid] + xchg_scflag(set, me)), seq);
    if (!xchg_wait(S, (const unsigned long long*)(X.mine + xchg_scflag(set, tid)), seq)) S->err = 1;
  }
  __syncthreads();
  double s = 0.0, mx = 0.0;
  for (int r = 0; r < world; r++) {
    const double av = (r == me) ? a_loc : __ldcg(X.mine + xchg_sc(set, r));
    const double bv = (r == me) ? b_loc : __ldcg(X.mine + xchg_sc(set, r) + 1);
    s += av;
    mx = fmax(mx, bv);
  }
  a_sum = s;
  b_max = mx;
  __syncthreads();
  if (tid == 0) S->sseq = seq;
}

inline bool set_args(const PcgArgs& a_in, cudaStream_t st) {
  // Between two termination checks consecutive solves of an ADMM run have identical argument
  // blocks (same vectors, same rho, prim_res / dual_res refreshed only at the checks; admm_iter is
  // only ever tested for == 1): the ~8 us constant-memory copy is skipped when nothing changed.
  static thread_local PcgArgs last;
  static thread_local bool have_last = false;
  static thread_local int last_slot = -1;
  PcgArgs a = a_in;
  if (a.admm_iter > 2) a.admm_iter = 2;
  if (have_last && ctx().args_valid && last_slot == ctx().slot && memcmp(&last, &a, sizeof(PcgArgs)) == 0) return true;
  // pageable source: the runtime stages the bytes before returning; stream-ordered with the
  // kernels of the previous solve
  const bool ok = B200_CHECK(cudaMemcpyToSymbolAsync(c_args, &a, sizeof(PcgArgs), sizeof(PcgArgs) * (size_t)ctx().slot,
                                                     cudaMemcpyHostToDevice, st));
  last = a;
  last_slot = ctx().slot;
  have_last = ok;
  ctx().args_valid = ok ? 1 : 0;
  return ok;
}

// step length and (predicted) direction coefficient of one CG iteration, from the three dots of
// the fused-operator pass.  beta uses the expansion of r+'M^-1 r+ so that x, r AND p can be
// updated by ONE vector kernel (no second grid-wide reduction before the direction update); the
// exactly reduced r+'y of that kernel replaces the prediction as the next iteration's r'y.
__device__ __forceinline__ void cg_step_scalars(PcgRun* run, double pKp, double rkp, double kpkp) {
  const double rTy   = run->rTy;
  const double alpha = rTy / pKp;
  const double pred  = rTy + alpha * (2.0 * rkp + alpha * kpkp);
  double beta = pred / rTy;
  if (!(beta > 0.0)) beta = 0.0;          // cancellation at convergence: restart direction
  run->alpha = alpha;
  run->beta  = beta;
}

// t = rho .* b2   (only for the ||rhs|| of the tolerance at admm_iter == 1 / polishing)
__global__ void __launch_bounds__(kBlock) g_rhs_t(int slot) {
  const PcgArgs& a = c_args[slot];
  const T* b2 = a.b + a.n;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.m; j += stride)
    a.t[j] = (a.rho_vec ? a.rho_vec[j] : a.rho) * b2[j];
}

// ||b1 + A' t||_inf  -> run->rhs_norm
__global__ void __launch_bounds__(kSpmvBlock, B200_SPMV_MINBLOCKS) g_rhs_norm(int slot, PcgRun* run, double* red,
                                                         int stride) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double shr[33];
  const PcgArgs& a = c_args[slot];
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
__device__ __forceinline__ void tolerance_step(PcgRun* run, int slot) {
  const PcgArgs& a = c_args[slot];
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
__global__ void g_tolerance(int slot, PcgRun* run) {
  if (threadIdx.x || blockIdx.x) return;
  tolerance_step(run, slot);
}

// P1 from the carried product: t = rho .* (Ax - b2)
__global__ void __launch_bounds__(kBlock) g_p1_carried(int slot) {
  const PcgArgs& a = c_args[slot];
  const T* b2 = a.b + a.n;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.m; j += stride)
    a.t[j] = (a.rho_vec ? a.rho_vec[j] : a.rho) * (a.Ax[j] - b2[j]);
}

// pass over A.  MODE 0: P1 with exact recomputation (Ax = A x ; t = rho .* (Ax - b2))
//               MODE 1: L1 (w = A p ; t = rho .* w)
template <int MODE>
__global__ void __launch_bounds__(kSpmvBlock, B200_SPMV_MINBLOCKS) g_pass_A(int slot) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const PcgArgs& a = c_args[slot];
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
__global__ void __launch_bounds__(kSpmvBlock, B200_SPMV_MINBLOCKS) g_pass_K(int slot, PcgRun* run, double* red,
                                                       int stride) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double shr[33];
  const PcgArgs& a = c_args[slot];
  Pipe pipe = pipe_init(dsm);
  const int n = a.n;
  const T* src = (MODE == 0 || MODE == 2) ? a.x : a.p;
  const T* t = a.t;
  const T* b1 = a.b;
  const T* minv = a.minv;
  T* r = a.r; T* p = a.p; T* Kp = a.Kp;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
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
          const T yk = minv[row] * s;
          acc0 += (double)p[row] * (double)s;
          acc1 += (double)r[row] * (double)yk;
          acc2 += (double)s * (double)yk;
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
    acc1 = block_sum(acc1, shr);
    acc2 = block_sum(acc2, shr);
    publish<false>(acc1, red, stride, SLOT_RKP, nullptr, false, shr, tot);
    publish<false>(acc2, red, stride, SLOT_KPKP, nullptr, false, shr, tot);
    if (publish<false>(acc0, red, stride, SLOT_PKP, &run->ticket[SLOT_PKP], true, shr, tot)) {
      const double rkp  = fold<false>(red, stride, SLOT_RKP, shr);
      const double kpkp = fold<false>(red, stride, SLOT_KPKP, shr);
      if (threadIdx.x == 0) {
        run->pKp = tot;
        cg_step_scalars(run, tot, rkp, kpkp);
        run->ticket[SLOT_PKP] = 0;
      }
    }
  }
}

// row-sharded: Kp[i] = (A_r' t)_i partial, for the ||rhs|| of the tolerance
__global__ void __launch_bounds__(kSpmvBlock, B200_SPMV_MINBLOCKS) g_pass_At(int slot) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const PcgArgs& a = c_args[slot];
  Pipe pipe = pipe_init(dsm);
  const T* t = a.t;
  T* Kp = a.Kp;
  const CsrView M = a.At;
  spmv_pass<SumOp>(
      M, blockIdx.x, gridDim.x, pipe, [&](int, int c, T v) { return v * t[c]; },
      [&](int row, T s) { Kp[row] = s; });
}

// row-sharded: ||b1 + Kp||_inf -> run->rhs_norm   (Kp = all-reduced A'(rho .* b2))
__global__ void __launch_bounds__(kBlock) g_rhs_norm_sum(int slot, PcgRun* run, double* red,
                                                         int stride, int have_At, int off) {
  __shared__ double shr[33];
  const PcgArgs& a = c_args[slot];
  double mx = 0.0;
  for (int i = off + blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x)
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
__global__ void __launch_bounds__(kBlock) g_resid_init(int slot, PcgRun* run, double* red,
                                                       int stride, int off, int p2p) {
  __shared__ double shr[33];
  const PcgArgs& a = c_args[slot];
  double acc0 = 0.0, acc1 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const T rr = a.Kp[i] - a.b[i];
    const T yy = a.minv[i] * rr;
    a.r[i] = rr;
    a.p[i] = -yy;
    if (i >= off) {       // column split: the shared slice is counted by rank 0 only
      acc0 += (double)rr * (double)yy;
      acc1 = fmax(acc1, fabs((double)rr));
    }
  }
  acc0 = block_sum(acc0, shr);
  acc1 = block_max(acc1, shr);
  double tot;
  publish<true>(acc1, red, stride, SLOT_RMAX, nullptr, false, shr, tot);
  if (publish<false>(acc0, red, stride, SLOT_RTY, &run->ticket[SLOT_RTY], true, shr, tot)) {
    double rmax = fold<true>(red, stride, SLOT_RMAX, shr);
    if (p2p) xchg_scalars(tot, rmax, tot, rmax);     // peer-memory exchange of (r'y, ||r||_inf)
    if (threadIdx.x == 0) {
      run->rTy = tot;
      run->rnorm = rmax;
      run->ticket[SLOT_RTY] = 0;
    }
  }
}

// row-sharded, peer-memory path: push the n_shared-long head of this rank's partial K p (MODE 1: and the
// three dot-product partials over the columns it owns) into every peer's exchange buffer, wait for the
// peers' pushes, fold the heads in rank order into Kp, and (MODE 1) finish the three dots -- the shared
// columns are counted here, identically on every rank -- and with them alpha and beta.  One kernel:
// the transfer over NVLink overlaps the tail of the pushing CTAs, nothing returns to the host.
// The grid must be co-resident (CTAs spin on the peers' sequence words): <= one CTA per SM.
template <int MODE>
__global__ void __launch_bounds__(kBlock) g_xchg_vector(int slot, PcgRun* run, double* red, int stride) {
  __shared__ double shr[48];
  __shared__ int s_lastx;
  const PcgArgs& a = c_args[slot];
  const XchgView& X = c_xchg;
  XchgState* S = X.state;
  const int ns = a.n_shared, me = X.rank, world = X.world;
  const unsigned long long seq = *(volatile unsigned long long*)&S->vseq + 1;
  const int set = (int)(seq & 1ull);
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  T* __restrict__ Kp = a.Kp;
  for (int q = 0; q < world; q++) {
    if (q == me) continue;
    double* dst = X.peer[q] + xchg_vec(set, me);
    for (int i = gtid; i < ns; i += gstride) dst[i] = (double)Kp[i];
    if (MODE == 1 && gtid < 3) dst[ns + gtid] = run->dots[gtid];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_lastx = (atomicAdd(&S->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_lastx) {           // every CTA of this rank has pushed (and fenced): publish the sequence number
    __threadfence_system();
    if ((int)threadIdx.x < world && (int)threadIdx.x != me)
      st_release_sys((unsigned long long*)(X.peer[threadIdx.x] + xchg_vflag(set, me)), seq);
    if (threadIdx.x == 0) S->ticket = 0;
  }
  if ((int)threadIdx.x < world && (int)threadIdx.x != me) {
    if (!xchg_wait(S, (const unsigned long long*)(X.mine + xchg_vflag(set, threadIdx.x)), seq)) S->err = 1;
  }
  __syncthreads();
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  for (int i = gtid; i < ns; i += gstride) {
    double sum = 0.0;
    for (int r = 0; r < world; r++) sum += (r == me) ? (double)Kp[i] : __ldcg(X.mine + xchg_vec(set, r) + i);
    const T kp = (T)sum;
    Kp[i] = kp;
    if (MODE == 1) {
      const T yk = a.minv[i] * kp;
      acc0 += (double)a.p[i] * (double)kp;
      acc1 += (double)a.r[i] * (double)yk;
      acc2 += (double)kp * (double)yk;
    }
  }
  block_sum3(acc0, acc1, acc2, shr);
  const double part[3] = {acc0, acc1, acc2};
  const int slots[3] = {SLOT_PKP, SLOT_RKP, SLOT_KPKP};
  double tot[3];
  if (publish_fold<3, false>(part, slots, red, stride, &S->ticket2, shr, tot) && threadIdx.x == 0) {
    if (MODE == 1) {
      for (int r = 0; r < world; r++)
        for (int j = 0; j < 3; j++)
          tot[j] += (r == me) ? run->dots[j] : __ldcg(X.mine + xchg_vec(set, r) + ns + j);
      run->pKp = tot[0];
      cg_step_scalars(run, tot[0], tot[1], tot[2]);
    }
    S->ticket2 = 0;
    S->vseq = seq;
  }
}

// row-sharded L2 tail: the three dots of the exchanged Kp (p'Kp, r'M^-1 Kp, Kp'M^-1 Kp) over the
// columns this rank counts; summed over the ranks afterwards when the layout is column-split
__global__ void __launch_bounds__(kBlock) g_dots3(int slot, PcgRun* run, double* red, int stride, int off) {
  __shared__ double shr[48];
  const PcgArgs& a = c_args[slot];
  double acc = 0.0, acc1 = 0.0, acc2 = 0.0;
  for (int i = off + blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const T kp = a.Kp[i];
    const T yk = a.minv[i] * kp;
    acc  += (double)a.p[i] * (double)kp;
    acc1 += (double)a.r[i] * (double)yk;
    acc2 += (double)kp * (double)yk;
  }
  block_sum3(acc, acc1, acc2, shr);
  const double part[3] = {acc, acc1, acc2};
  const int slots[3] = {SLOT_PKP, SLOT_RKP, SLOT_KPKP};
  double tot[3];
  if (publish_fold<3, false>(part, slots, red, stride, &run->ticket[SLOT_PKP], shr, tot) && threadIdx.x == 0) {
    run->dots[0] = tot[0];
    run->dots[1] = tot[1];
    run->dots[2] = tot[2];
    run->ticket[SLOT_PKP] = 0;
  }
}
// column-split layout: r'y (a sum) and ||r||_inf (a maximum) of the ranks travel in ONE sum
// all-reduce: every rank writes its pair into its own slot of a zeroed array ...
__global__ void g_slots_put(PcgRun* run, int rank, int world) {
  if (threadIdx.x || blockIdx.x) return;
  for (int r = 0; r < world; r++) {
    run->slots[2 * r]     = (r == rank) ? run->rTy : 0.0;
    run->slots[2 * r + 1] = (r == rank) ? run->rnorm : 0.0;
  }
}
// ... and folds the gathered pairs in rank order afterwards (identical on every rank)
__global__ void g_slots_fold(PcgRun* run, int world) {
  if (threadIdx.x || blockIdx.x) return;
  double s = 0.0, mx = 0.0;
  for (int r = 0; r < world; r++) {
    s += run->slots[2 * r];
    mx = fmax(mx, run->slots[2 * r + 1]);
  }
  run->rTy = s;
  run->rnorm = mx;
}
// alpha, beta from the (exchanged) dots
__global__ void g_step_scalars(PcgRun* run) {
  if (threadIdx.x || blockIdx.x) return;
  run->pKp = run->dots[0];
  cg_step_scalars(run, run->dots[0], run->dots[1], run->dots[2]);
}

// ---------------------------------------------------------------------------------------------
// Lean passes: the same CSR-stream tile as spmv_pass, written out flat -- 512 threads per 2048-entry
// tile, a single batch of 4 (col, val, gather) chains per thread, no generic functors (<= 42
// registers, 3 CTAs per SM); chunks of rows longer than a tile are summed CTA-wide and folded by the
// last chunk to arrive.  The grid is
// ONE WAVE (3 CTAs per SM) looping over the tiles round-robin: the per-CTA tail -- block
// reduction of the dot-product partials, publication, ticket -- is then paid 444 times per pass
// instead of once per tile (6100 times for the Lasso operator), which measured 45 us of a 120 us
// pass (profiles/r01_phase_profile.md).
//   MODE 0: L1   w = A p ; t = rho .* w
//   MODE 1: L2   Kp = [P + sigma I | A'] [p; t] ; totals p'Kp, r'M^-1 Kp, Kp'M^-1 Kp -> alpha, beta
//   MODE 2: P2   r = K2 [x; t] - b1 ; p = -M^-1 r ; totals r'y, ||r||_inf
//   MODE 3: P1   Ax = A x ; t = rho .* (Ax - b2)       (exact recomputation of the carried product)
//   MODE 4: Kp = A' t   MODE 5: Kp = K2 [p; t]   MODE 6: Kp = K2 [x; t]   (plain stores: the row-sharded
//           driver exchanges the partial Kp before any scalar is formed; also the phase profile)
//   MODE 7: Kp = K2 [p; t] and the three dot partials over the rows >= n_shared (the columns this rank
//           OWNS, complete without any exchange) -> run->dots  (row-sharded peer-memory path)
constexpr int kLeanBlock = 512;
constexpr int kLeanCtasPerSm = 3;
static_assert(kTile == 4 * kLeanBlock, "lean pass assumes one batch of 4 per thread");

template <int MODE>
__global__ void __launch_bounds__(kLeanBlock, kLeanCtasPerSm) g_lean_pass(int slot, PcgRun* run, double* red,
                                                                          int stride) {
  __shared__ T sm[kTile];
  __shared__ int srp[kMaxRows + 1];
  __shared__ double shr[48];
  const PcgArgs& a = c_args[slot];
  // everything the tiles need is pulled out of the (global-memory) argument block once
  constexpr bool kOverA = (MODE == 0 || MODE == 3 || MODE == 4);
  const CsrView& M = (MODE == 4) ? a.At : (kOverA ? a.A : a.K2);
  const int* __restrict__ row_ptr = M.row_ptr;
  const int* __restrict__ col_ind = M.col_ind;
  const T* __restrict__   val     = M.val;
  const int4* __restrict__ desc   = M.desc;
  const int nblocks = M.nblocks;
  const T* __restrict__ src = (MODE == 4) ? a.t : ((MODE == 0 || MODE == 1 || MODE == 5 || MODE == 7) ? a.p : a.x);   // 2, 3, 6: x
  const int n_shared = a.n_shared;
  const T* __restrict__ t = a.t;
  const int n = a.n;
  const int tid = threadIdx.x;
  double acc = 0.0, acc1 = 0.0, acc2 = 0.0;
  // what happens to a finished row sum (exactly once per row, by one thread)
  auto emit = [&](int row, T sum) {
    if (MODE == 0) {
      a.w[row] = sum;
      a.t[row] = (a.rho_vec ? a.rho_vec[row] : a.rho) * sum;
    } else if (MODE == 3) {
      a.Ax[row] = sum;
      a.t[row]  = (a.rho_vec ? a.rho_vec[row] : a.rho) * (sum - a.b[n + row]);
    } else if (MODE == 7) {
      a.Kp[row] = sum;
      if (row >= n_shared) {
        const T yk = a.minv[row] * sum;
        acc  += (double)src[row] * (double)sum;
        acc1 += (double)a.r[row] * (double)yk;
        acc2 += (double)sum * (double)yk;
      }
    } else if (MODE >= 4) {
      a.Kp[row] = sum;
    } else if (MODE == 1) {
      // Kp and the three dots that fix alpha AND beta before the vector update:
      //   r+ = r + alpha Kp  =>  r+' M^-1 r+ = r'y + 2 alpha r'M^-1 Kp + alpha^2 Kp'M^-1 Kp
      a.Kp[row] = sum;
      const T yk = a.minv[row] * sum;
      acc  += (double)src[row] * (double)sum;
      acc1 += (double)a.r[row] * (double)yk;
      acc2 += (double)sum * (double)yk;
    } else {
      const T rr = sum - a.b[row];
      const T yy = a.minv[row] * rr;
      a.r[row] = rr;
      a.p[row] = -yy;
      acc  += (double)rr * (double)yy;
      acc1 = fmax(acc1, fabs((double)rr));
    }
  };
  for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const int4 d = __ldg(desc + b);
    const int nnz0 = d.z, cnt = d.w;
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
    if (d.y < 0) {
      // ---- one <= kTile-entry chunk of a row longer than a tile (e.g. the 1e4-entry feature rows of A' in
      // BASELINE configs[3]): CTA-wide sum, partial published, the LAST chunk to arrive folds the partials in
      // chunk order (integer ticket: deterministic, no floating-point atomics) and emits the row
      const int  lr   = -d.y - 1;
      const int4 info = __ldg(M.long_rows + lr);       // {row, first block, chunks, -}
      double part = 0.0;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int k = u * kLeanBlock + tid;
        if (k < cnt) part += (double)(kOverA ? v[u] * src[c[u]] : v[u] * (c[u] < n ? src[c[u]] : t[c[u] - n]));
      }
      part = block_sum(part, shr);
      __shared__ int s_lastchunk;
      if (tid == 0) {
        M.long_partials[b] = part;
        __threadfence();
        s_lastchunk = (atomicAdd(&M.long_counters[lr], 1u) == (unsigned)(info.z - 1));
      }
      __syncthreads();
      if (s_lastchunk) {
        __threadfence();
        double tot = 0.0;
        for (int k = tid; k < info.z; k += kLeanBlock) tot += __ldcg(&M.long_partials[info.y + k]);
        tot = block_sum(tot, shr);
        if (tid == 0) {
          M.long_counters[lr] = 0;
          emit(info.x, (T)tot);
        }
      }
      __syncthreads();
      continue;
    }
    const int nrows = d.y & 0xffffff, lg = d.y >> 24;
    for (int i = tid; i <= nrows; i += kLeanBlock) srp[i] = ld_stream(row_ptr + d.x + i) - nnz0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int k = u * kLeanBlock + tid;
      if (k < cnt) {
        if (kOverA) sm[k] = v[u] * src[c[u]];
        else sm[k] = v[u] * (c[u] < n ? src[c[u]] : t[c[u] - n]);
      }
    }
    __syncthreads();
    const int g = 1 << lg, gid = tid >> lg, lig = tid & (g - 1), ngroup = kLeanBlock >> lg;
    for (int base = 0; base < nrows; base += ngroup) {   // one trip except for tiles of 1-2 entry rows
      const int r = base + gid;
      T sum = 0;
      if (r < nrows) {
        const int e = srp[r + 1];
        for (int k = srp[r] + lig; k < e; k += g) sum += sm[k];
      }
      for (int o = g >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (r < nrows && lig == 0) emit(d.x + r, sum);
    }
    __syncthreads();   // sm / srp are reused by the next tile
  }
  if (MODE == 1) {
    block_sum3(acc, acc1, acc2, shr);
    const double part[3] = {acc, acc1, acc2};
    const int slots[3] = {SLOT_PKP, SLOT_RKP, SLOT_KPKP};
    double tot[3];
    if (publish_fold<3, false>(part, slots, red, stride, &run->ticket[SLOT_PKP], shr, tot) && tid == 0) {
      run->pKp = tot[0];
      cg_step_scalars(run, tot[0], tot[1], tot[2]);
      run->ticket[SLOT_PKP] = 0;
    }
  } else if (MODE == 7) {
    block_sum3(acc, acc1, acc2, shr);
    const double part[3] = {acc, acc1, acc2};
    const int slots[3] = {SLOT_PKP, SLOT_RKP, SLOT_KPKP};
    double tot[3];
    if (publish_fold<3, false>(part, slots, red, stride, &run->ticket[SLOT_PKP], shr, tot) && tid == 0) {
      run->dots[0] = tot[0];
      run->dots[1] = tot[1];
      run->dots[2] = tot[2];
      run->ticket[SLOT_PKP] = 0;
    }
  } else if (MODE == 2) {
    block_sum_max(acc, acc1, shr);
    const double part[2] = {acc, acc1};
    const int slots[2] = {SLOT_RTY, SLOT_RMAX};
    double tot[2];
    if (publish_fold<2, true>(part, slots, red, stride, &run->ticket[SLOT_RTY], shr, tot) && tid == 0) {
      run->rTy = tot[0];
      run->rnorm = tot[1];
      run->ticket[SLOT_RTY] = 0;
    }
  }
}

// passes that end in a grid-wide reduction run as one wave; passes without one (the A passes) run
// one CTA per tile, which the hardware schedules dynamically (58 vs 65 us on the Lasso A)
inline int lean_grid(const b200_csr& M, bool reduces = true) {
  const int wave = ctx().sm_count * kLeanCtasPerSm;
  const int g = (reduces && M.nblocks > wave) ? wave : M.nblocks;
  return g > 0 ? g : 1;
}

// the flat kernel takes every tile the schedule builder produces (normal blocks of <= kTile entries and
// <= kTile-entry chunks of over-long rows); B200_PCG_LEAN_NO_LONG=1 restores the round-1 restriction
static bool lean_ok(const b200_csr& M, const std::vector<int4>& desc) {
  static const bool no_long = getenv("B200_PCG_LEAN_NO_LONG") != nullptr;
  for (const int4& d : desc) {
    if (d.w > kTile) return false;
    if (d.y < 0 && no_long) return false;
  }
  return !(no_long && M.nlong > 0);
}

// first node of the loop graph: arm the WHILE condition from the initial residual
__global__ void g_loop_init(int slot, PcgRun* run, cudaGraphConditionalHandle h) {
  if (threadIdx.x || blockIdx.x) return;
  tolerance_step(run, slot);      // the schedule only needs ||rhs|| (first solve / polish), known by now
  cudaGraphSetConditional(h, (run->rnorm > run->eps && run->it < c_args[slot].max_iter) ? 1u : 0u);
}

// L3+L4 in one kernel (graph driver): x += a p ; r += a Kp ; p = beta p - M^-1 r ; Ax += a w ;
// totals r'y (exact), ||r||_inf ; last CTA: it++, loop condition.  8 n F + 3 m F bytes.
__global__ void __launch_bounds__(kBlock) g_update_fused(int slot, PcgRun* run, double* red, int stride,
                                                         cudaGraphConditionalHandle h, int off, int p2p) {
  __shared__ double shr[33];
  const PcgArgs& a = c_args[slot];
  const int n = a.n, m = a.m;
  const T alpha = (T)run->alpha, beta = (T)run->beta;
  T* __restrict__ x = a.x; T* __restrict__ p = a.p; T* __restrict__ r = a.r;
  const T* __restrict__ Kp = a.Kp; const T* __restrict__ minv = a.minv;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  double acc_rty = 0.0, acc_max = 0.0;
  for (int i = gtid; i < n; i += gstride) {
    const T pi = p[i];
    x[i] += alpha * pi;
    const T rr = r[i] + alpha * Kp[i];
    r[i] = rr;
    const T yy = minv[i] * rr;
    p[i] = beta * pi - yy;
    if (i >= off) {      // column-split layout: the shared slice is counted by rank 0 only
      acc_rty += (double)rr * (double)yy;
      acc_max = fmax(acc_max, fabs((double)rr));
    }
  }
  T* __restrict__ Ax = a.Ax; const T* __restrict__ w = a.w;
  for (int j = gtid; j < m; j += gstride) Ax[j] += alpha * w[j];
  acc_rty = block_sum(acc_rty, shr);
  acc_max = block_max(acc_max, shr);
  double tot;
  publish<true>(acc_max, red, stride, SLOT_RMAX, nullptr, false, shr, tot);
  if (publish<false>(acc_rty, red, stride, SLOT_RTY, &run->ticket[SLOT_RTY], true, shr, tot)) {
    double rmax = fold<true>(red, stride, SLOT_RMAX, shr);
    if (p2p) xchg_scalars(tot, rmax, tot, rmax);     // row-sharded: (r'y, ||r||_inf) of all ranks, over NVLink
    if (threadIdx.x == 0) {
      run->rTy   = tot;
      run->rnorm = rmax;
      run->it   += 1;
      run->ticket[SLOT_RTY] = 0;
      if (h) cudaGraphSetConditional(h, (rmax > run->eps && run->it < a.max_iter) ? 1u : 0u);
    }
  }
}

// E1: b1 = x ; b2 = A x (carried) or (A x - b2)/delta when polishing ; persist the schedule state
__global__ void __launch_bounds__(kBlock) g_epilogue(int slot, PcgRun* run) {
  const PcgArgs& a = c_args[slot];
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

// grid of the exchange kernel: its CTAs spin on the peers' sequence words, so all of them must be
// resident at once -- at most one per SM
inline int xchg_grid(int n_shared) {
  int g = (n_shared + kBlock * 2 - 1) / (kBlock * 2);
  if (g > ctx().sm_count) g = ctx().sm_count;
  return g > 0 ? g : 1;
}

// one wave of co-resident CTAs looping over the tiles (see the note on per-CTA tails above)
inline int pass_grid(const b200_csr& M, int cap) {
  const int wave = ctx().sm_count * B200_SPMV_MINBLOCKS;
  if (cap > wave) cap = wave;
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
    lean = lean_ok(s->K2, fetch(s->K2)) && (s->m == 0 || lean_ok(*s->A, fetch(*s->A)));
  }
  s->lean = lean ? 1 : 0;
  if (getenv("B200_TRACE_SETUP"))
    fprintf(stderr, "[b200 trace] graph PCG driver: %s passes, K2 tiles %d, A tiles %d, partial stride %d\n",
            lean ? "lean" : "generic", s->K2.nblocks, s->m > 0 ? s->A->nblocks : 0, s->gred_stride);
  // Row-sharded solve.  With the peer-memory exchange (NVLink P2P stores inside the kernels, dist.cu) the
  // loop stays a CUDA-graph WHILE node: A pass, operator pass with the owned-column dots, ONE exchange
  // kernel (vector head + dots -> alpha, beta), fused update whose last CTA trades (r'y, ||r||_inf) with
  // the peers and sets the loop condition.  Otherwise (no P2P, plain row layout, over-long rows, head
  // larger than an exchange slot): host-driven loop with NCCL all-reduces.
  bool p2p = s->sharded && lean && dist_p2p_ready() && dist_split() && getenv("B200_DIST_NO_P2P") == nullptr &&
             dist_n_shared() + 8 <= kXchgCap && s->m > 0;
  if (s->sharded) {
    // the ranks must agree (a rank whose shard has an over-long row cannot take the lean path): one
    // MAX all-reduce of "I cannot" at solver creation, which every rank reaches in lockstep
    double veto = p2p ? 0.0 : 1.0;
    double* d_veto = reinterpret_cast<double*>(s->d_run);     // scratch: zeroed again below
    ok &= B200_CHECK(cudaMemcpyAsync(d_veto, &veto, sizeof(double), cudaMemcpyHostToDevice, c.stream));
    dist_allreduce_f64(d_veto, 1, true);
    ok &= B200_CHECK(cudaMemcpyAsync(&veto, d_veto, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    ok &= B200_CHECK(cudaStreamSynchronize(c.stream));
    ok &= B200_CHECK(cudaMemsetAsync(s->d_run, 0, sizeof(PcgRun), c.stream));
    if (!ok) return 1;
    p2p = p2p && veto == 0.0;
  }
  s->p2p = p2p ? 1 : 0;
  if (s->sharded && !p2p) return 0;
  if (p2p) {
    const XchgView v = dist_xchg_view();
    if (!B200_CHECK(cudaMemcpyToSymbolAsync(c_xchg, &v, sizeof(v), 0, cudaMemcpyHostToDevice, c.stream))) return 1;
    if (!B200_CHECK(cudaStreamSynchronize(c.stream))) return 1;
  }
  const int x_off = p2p ? dist_col_off() : 0;
  const int x_p2p = p2p ? 1 : 0;

  cudaGraph_t g = nullptr;
  if (!B200_CHECK(cudaGraphCreate(&g, 0))) return 1;
  cudaGraphConditionalHandle h;
  if (!B200_CHECK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault))) return 1;

  const int d_args = ctx().slot;
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

  // body: L1 -> L2 -> L3+L4
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
  if (s->m > 0) {
    if (lean) {
      void* a1[] = {(void*)&d_args, (void*)&d_run, (void*)&d_red, (void*)&stride};
      add((void*)g_lean_pass<0>, dim3(lean_grid(*s->A, false)), dim3(kLeanBlock), 0, a1);
    } else {
      void* a1[] = {(void*)&d_args};
      add((void*)g_pass_A<1>, dim3(pass_grid(*s->A, cap)), dim3(kSpmvBlock), kSpmvSmemBytes, a1);
    }
  }
  {
    void* a2[] = {(void*)&d_args, (void*)&d_run, (void*)&d_red, (void*)&stride};
    if (p2p) {
      add((void*)g_lean_pass<7>, dim3(lean_grid(s->K2)), dim3(kLeanBlock), 0, a2);
      add((void*)g_xchg_vector<1>, dim3(xchg_grid(dist_n_shared())), dim3(kBlock), 0, a2);
    } else if (lean) add((void*)g_lean_pass<1>, dim3(lean_grid(s->K2)), dim3(kLeanBlock), 0, a2);
    else add((void*)g_pass_K<1>, dim3(pass_grid(s->K2, cap)), dim3(kSpmvBlock), kSpmvSmemBytes, a2);
  }
  {
    // L3 + L4 in one kernel: alpha AND beta are known after the fused-operator pass
    int zero_off = x_off, use_p2p = x_p2p;
    void* a3[] = {(void*)&d_args, (void*)&d_run, (void*)&d_red, (void*)&stride, (void*)&h, (void*)&zero_off, (void*)&use_p2p};
    int nm = s->n > s->m ? s->n : s->m;
    int gu = ew_grid(nm) < cap ? ew_grid(nm) : cap;
    add((void*)g_update_fused, dim3(gu), dim3(kBlock), 0, a3);
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
  set_args(a, st);
  ctx().epoch++;
  const int d_args = ctx().slot;
  if (a.polishing || a.admm_iter == 1) {
    if (m > 0) {
      g_rhs_t<<<ew_grid(m), kBlock, 0, st>>>(d_args);
      count_launch("g_rhs_t");
    }
    const int g = m > 0 ? pass_grid(*s->At, cap) : (ew_grid(n) < cap ? ew_grid(n) : cap);
    g_rhs_norm<<<g, kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, s->d_run, s->d_gred, cap);
    count_launch("g_rhs_norm");
  }
  if (m > 0) {
    if (a.ax_valid) g_p1_carried<<<ew_grid(m), kBlock, 0, st>>>(d_args);
    else if (s->lean) g_lean_pass<3><<<lean_grid(*s->A, false), kLeanBlock, 0, st>>>(d_args, s->d_run, s->d_gred, cap);
    else g_pass_A<0><<<pass_grid(*s->A, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
    count_launch("g_lean_pass<3>");
  }
  if (s->lean) g_lean_pass<2><<<lean_grid(s->K2), kLeanBlock, 0, st>>>(d_args, s->d_run, s->d_gred, cap);
  else g_pass_K<0><<<pass_grid(s->K2, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, s->d_run, s->d_gred, cap);
  count_launch("g_lean_pass<2>");
  bool ok = true;
  static const bool hostloop = getenv("B200_PCG_HOSTLOOP") != nullptr;
  if (hostloop) {
    // profiling aid: ncu cannot see kernel nodes inside a conditional graph, so run the very same
    // loop body as plain launches, with the condition read back by the host once per iteration
    cudaGraphConditionalHandle none = 0;
    g_tolerance<<<1, 32, 0, st>>>(d_args, s->d_run);
    count_launch("g_tolerance");
    const int nm0 = n > m ? n : m;
    const int gu = ew_grid(nm0) < cap ? ew_grid(nm0) : cap;
    PcgRun h;
    for (;;) {
      ok &= B200_CHECK(cudaMemcpyAsync(&h, s->d_run, sizeof(PcgRun), cudaMemcpyDeviceToHost, st));
      ok &= B200_CHECK(cudaStreamSynchronize(st));
      if (!ok || !(h.rnorm > h.eps && h.it < a.max_iter)) break;
      if (m > 0) {
        if (s->lean) g_lean_pass<0><<<lean_grid(*s->A, false), kLeanBlock, 0, st>>>(d_args, s->d_run, s->d_gred, cap);
        else g_pass_A<1><<<pass_grid(*s->A, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
        count_launch("L1 pass A");
      }
      if (s->lean) g_lean_pass<1><<<lean_grid(s->K2), kLeanBlock, 0, st>>>(d_args, s->d_run, s->d_gred, cap);
      else g_pass_K<1><<<pass_grid(s->K2, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, s->d_run, s->d_gred, cap);
      count_launch("L2 pass K2");
      g_update_fused<<<gu, kBlock, 0, st>>>(d_args, s->d_run, s->d_gred, cap, none, 0, 0);
      count_launch("L3+L4 update");
    }
  } else {
    ok = B200_CHECK(cudaGraphLaunch((cudaGraphExec_t)s->graph_exec, st));
    count_launch("graph(loop)");
    ctx().graph_launches++;
  }
  const int nm = n > m ? n : m;
  g_epilogue<<<ew_grid(nm), kBlock, 0, st>>>(d_args, s->d_run);
  count_launch("g_epilogue");
  return ok ? 0 : 1;
}

// ------------------------------------------------------------------ row-sharded driver
// Every rank holds a row block A_r (CSR), A_r' and the matching slices of the m-vectors.
//   * plain row sharding: the n-vectors x, p, r, Kp, M^-1 are replicated; K p = (P + sigma I) p
//     [rank 0 only] + A_r' (rho .* (A_r p)) summed by one all-reduce of the length-n partial per CG
//     iteration; every CG scalar is then computed redundantly from replicated vectors.
//   * column-split layout (dist_split()): the n-vectors are [shared ; local]; only
//     the SHARED slice of the partial K p is all-reduced (the local columns are complete on their
//     owner, which also carries their P + sigma I), and the CG scalars -- sums over shared columns
//     once plus every rank's local columns -- take two small all-reduces per iteration (the three
//     dots that fix alpha and beta; then r'y and ||r||_inf of the updated residual).
// The loop is driven by the host, which reads (||r||_inf, eps, it) back once per iteration.
namespace {
inline void exchange_vector(T* d_v, int n) {
  b200_dist_allreduce_sum(d_v, dist_split() ? dist_n_shared() : n);
}
// scalars of `run` that were reduced over this rank's columns only
inline void exchange_scalars(double* d_first, int count, bool is_max) {
  if (dist_split()) dist_allreduce_f64(d_first, count, is_max);
}
// (r'y, ||r||_inf): one collective for the pair
inline void exchange_residual_scalars(PcgRun* run, cudaStream_t st) {
  if (!dist_split()) return;
  const int world = b200_dist_world(), rank = b200_dist_rank();
  if (world > 8) {       // slots[] holds 8 ranks; beyond that fall back to two collectives
    dist_allreduce_f64(&run->rTy, 1, false);
    dist_allreduce_f64(&run->rnorm, 1, true);
    return;
  }
  g_slots_put<<<1, 32, 0, st>>>(run, rank, world);
  count_launch("g_slots_put");
  dist_allreduce_f64(run->slots, 2 * world, false);
  g_slots_fold<<<1, 32, 0, st>>>(run, world);
  count_launch("g_slots_fold");
}
}  // namespace

int b200_pcg_sharded_solve(b200_pcg* s, const PcgArgs& a) {
  Context& c = ctx();
  cudaStream_t st = c.stream;
  const int cap = s->gred_stride;
  const int n = s->n, m = s->m;
  const int gn = ew_grid(n) < cap ? ew_grid(n) : cap;
  const int off = dist_col_off();
  set_args(a, st);
  ctx().epoch++;
  const int d_args = ctx().slot;
  PcgRun* run = s->d_run;
  cudaGraphConditionalHandle none = 0;

  if (a.polishing || a.admm_iter == 1) {
    if (m > 0) {
      g_rhs_t<<<ew_grid(m), kBlock, 0, st>>>(d_args);
      count_launch("g_rhs_t");
      g_pass_At<<<pass_grid(*s->At, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
      count_launch("g_rhs_t");
    } else {
      B200_CHECK(cudaMemsetAsync(s->d_Kp, 0, sizeof(T) * n, st));
    }
    exchange_vector(s->d_Kp, n);
    g_rhs_norm_sum<<<gn, kBlock, 0, st>>>(d_args, run, s->d_gred, cap, 1, off);
    count_launch("g_rhs_norm_sum");
    exchange_scalars(&run->rhs_norm, 1, true);
  }
  if (s->p2p) {
    // peer-memory path: no NCCL call and no host read-back from here on; the loop is the graph
    if (a.ax_valid) g_p1_carried<<<ew_grid(m), kBlock, 0, st>>>(d_args);
    else g_lean_pass<3><<<lean_grid(*s->A, false), kLeanBlock, 0, st>>>(d_args, run, s->d_gred, cap);
    count_launch("P1");
    g_lean_pass<6><<<lean_grid(s->K2, false), kLeanBlock, 0, st>>>(d_args, run, s->d_gred, cap);
    count_launch("P2 partial");
    g_xchg_vector<0><<<xchg_grid(dist_n_shared()), kBlock, 0, st>>>(d_args, run, s->d_gred, cap);
    count_launch("xchg(vector)");
    g_resid_init<<<gn, kBlock, 0, st>>>(d_args, run, s->d_gred, cap, off, 1);
    count_launch("g_resid_init+xchg(scalars)");
    const bool okg = B200_CHECK(cudaGraphLaunch((cudaGraphExec_t)s->graph_exec, st));
    count_launch("graph(loop)");
    ctx().graph_launches++;
    const int nm1 = n > m ? n : m;
    g_epilogue<<<ew_grid(nm1), kBlock, 0, st>>>(d_args, run);
    count_launch("g_epilogue");
    return okg ? 0 : 1;
  }
  g_tolerance<<<1, 32, 0, st>>>(d_args, run);
  count_launch("g_tolerance");
  if (m > 0) {
    if (a.ax_valid) g_p1_carried<<<ew_grid(m), kBlock, 0, st>>>(d_args);
    else if (s->lean) g_lean_pass<3><<<lean_grid(*s->A, false), kLeanBlock, 0, st>>>(d_args, run, s->d_gred, cap);
    else g_pass_A<0><<<pass_grid(*s->A, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
    count_launch("g_p1_carried");
  }
  if (s->lean) g_lean_pass<6><<<lean_grid(s->K2, false), kLeanBlock, 0, st>>>(d_args, run, s->d_gred, cap);
  else g_pass_K<2><<<pass_grid(s->K2, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, run, s->d_gred, cap);
  count_launch("g_p1_carried");
  exchange_vector(s->d_Kp, n);
  g_resid_init<<<gn, kBlock, 0, st>>>(d_args, run, s->d_gred, cap, off, 0);
  count_launch("g_resid_init");
  exchange_residual_scalars(run, st);

  PcgRun h;
  bool ok = true;
  const int nm = n > m ? n : m;
  const int gu = ew_grid(nm) < cap ? ew_grid(nm) : cap;
  for (;;) {
    ok &= B200_CHECK(cudaMemcpyAsync(&h, run, sizeof(PcgRun), cudaMemcpyDeviceToHost, st));
    ok &= B200_CHECK(cudaStreamSynchronize(st));
    if (ctx().trace_on) trace_point("host read of the loop condition");
    if (!ok || !(h.rnorm > h.eps && h.it < a.max_iter)) break;
    if (m > 0) {
      if (s->lean) g_lean_pass<0><<<lean_grid(*s->A, false), kLeanBlock, 0, st>>>(d_args, run, s->d_gred, cap);
      else g_pass_A<1><<<pass_grid(*s->A, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args);
      count_launch("g_resid_init");
    }
    if (s->lean) g_lean_pass<5><<<lean_grid(s->K2, false), kLeanBlock, 0, st>>>(d_args, run, s->d_gred, cap);
    else g_pass_K<3><<<pass_grid(s->K2, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, run, s->d_gred, cap);
    count_launch("g_resid_init");
    exchange_vector(s->d_Kp, n);
    g_dots3<<<gn, kBlock, 0, st>>>(d_args, run, s->d_gred, cap, off);
    count_launch("g_dots3");
    exchange_scalars(run->dots, 3, false);
    g_step_scalars<<<1, 32, 0, st>>>(run);
    count_launch("g_step_scalars");
    g_update_fused<<<gu, kBlock, 0, st>>>(d_args, run, s->d_gred, cap, none, off, 0);
    count_launch("g_update_fused");
    exchange_residual_scalars(run, st);
  }
  g_epilogue<<<ew_grid(nm), kBlock, 0, st>>>(d_args, run);
  count_launch("g_epilogue");
  return ok ? 0 : 1;
}

// ------------------------------------------------------------------ phase profile (development aid)
// Times the kernels of one CG iteration of the most recently created solver one by one (plain
// launches, CUDA events on the library stream).  The iterate is left in an arbitrary state:
// call it on a solver that is thrown away afterwards.
namespace {
thread_local b200_pcg* g_profile_target = nullptr;
__global__ void g_nop() {}
}
void b200_pcg_profile_register(b200_pcg* s, bool alive) {
  if (alive) g_profile_target = s;
  else if (g_profile_target == s) g_profile_target = nullptr;
}

extern "C" int b200_pcg_profile_last(int reps, double* out_us, int nout) {
  b200_pcg* s = g_profile_target;
  if (!s || !s->d_args || !s->lean || nout < 12) return -1;   // sharded: the passes of this rank, no exchange
  Context& c = ctx();
  cudaStream_t st = c.stream;
  const int cap = s->gred_stride;
  const int d_args = ctx().slot;
  PcgRun* run = s->d_run;
  double* red = s->d_gred;
  const int n = s->n, m = s->m, nm = n > m ? n : m;
  const int gu = ew_grid(nm) < cap ? ew_grid(nm) : cap;
  cudaGraphConditionalHandle none = 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](auto&& body) {
    body();
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps; i++) body();
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return (double)ms * 1e3 / reps;
  };
  auto passA  = [&] { if (m > 0) g_lean_pass<0><<<lean_grid(*s->A, false), kLeanBlock, 0, st>>>(d_args, run, red, cap); };
  auto passK  = [&] { g_lean_pass<1><<<lean_grid(s->K2), kLeanBlock, 0, st>>>(d_args, run, red, cap); };
  auto upd    = [&] { g_update_fused<<<gu, kBlock, 0, st>>>(d_args, run, red, cap, none, 0, 0); };
  out_us[0]  = timeit(passA);
  out_us[1]  = timeit(passK);
  out_us[2]  = timeit(upd);
  out_us[3]  = timeit([&] { passA(); passK(); upd(); });
  out_us[4]  = timeit([&] { g_pass_K<0><<<pass_grid(s->K2, cap), kSpmvBlock, kSpmvSmemBytes, st>>>(d_args, run, red, cap); });
  out_us[5]  = timeit([&] { g_lean_pass<2><<<lean_grid(s->K2), kLeanBlock, 0, st>>>(d_args, run, red, cap); });
  out_us[6]  = timeit([&] { if (m > 0) g_p1_carried<<<ew_grid(m), kBlock, 0, st>>>(d_args); });
  out_us[7]  = timeit([&] { g_epilogue<<<ew_grid(nm), kBlock, 0, st>>>(d_args, run); });
  out_us[8]  = 0.0;   // (slots of the retired two-kernel L3 / L4 update)
  out_us[9]  = 0.0;
  out_us[10] = timeit([&] { g_nop<<<1, 32, 0, st>>>(); });   // launch floor
  out_us[11] = timeit([&] { if (m > 0) g_lean_pass<3><<<lean_grid(*s->A, false), kLeanBlock, 0, st>>>(d_args, run, red, cap); });
  if (nout >= 14) {
    out_us[12] = timeit([&] { if (m > 0) g_lean_pass<4><<<lean_grid(*s->At), kLeanBlock, 0, st>>>(d_args, run, red, cap); });
    out_us[13] = timeit([&] { g_lean_pass<5><<<lean_grid(s->K2), kLeanBlock, 0, st>>>(d_args, run, red, cap); });
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}
