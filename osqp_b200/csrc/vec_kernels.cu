// vec_kernels.cu -- elementwise and reduction kernels behind OSQPVectorf_* .
//
// One grid-stride kernel template instantiated with device lambdas replaces the 19
// hand-written elementwise kernels + cuBLAS axpy/scal/copy/dot/amax/asum calls of the
// reference CUDA backend (algebra/cuda/src/cuda_lin_alg.cu:38-358, :543-780).  Numerics
// follow the reference CPU backend (algebra/builtin/vector.c), which is the parity oracle.
//
// HBM roofline: every kernel here streams each operand exactly once; algorithmic bytes are
// (#operands read + #written) * n * sizeof(T).
#include "common.cuh"

using namespace b200;

namespace {

template <class F>
__global__ void __launch_bounds__(kBlock) ew_kernel(int n, F f) {
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  // 4-way unrolled grid-stride loop: 4 independent loads in flight per thread
  for (; i + 3 * stride < n; i += 4 * stride) {
    f(i);
    f(i + stride);
    f(i + 2 * stride);
    f(i + 3 * stride);
  }
  for (; i < n; i += stride) f(i);
}

template <class F>
inline void launch_ew(int n, F f) {
  if (n <= 0) return;
  ew_kernel<<<ew_grid(n), kBlock, 0, ctx().stream>>>(n, f);
  count_launch();
}

enum RedOp { RED_SUM = 0, RED_MAX = 1 };

// Two-stage deterministic reduction: per-thread accumulate -> warp shuffle -> CTA partial;
// the last CTA to arrive (ticket) folds the partials in index order and writes the scalar.
template <int OP, class F>
__global__ void __launch_bounds__(kBlock) reduce_kernel(int n, F f, double* partials,
                                                        unsigned* ticket, double* out, double* mail,
                                                        unsigned long long seq) {
  __shared__ double sh[33];
  __shared__ bool   is_last;
  const int stride = gridDim.x * blockDim.x;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v = f(i);
    if (OP == RED_SUM) acc += v;
    else acc = (v > acc) ? v : acc;     // NaN-ignoring, like the CPU loop's `if (a > max)`
  }
  acc = (OP == RED_SUM) ? block_sum(acc, sh) : block_max(acc, sh);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = acc;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double a = 0.0;
    for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
      double v = __ldcg(&partials[b]);
      if (OP == RED_SUM) a += v;
      else a = (v > a) ? v : a;
    }
    a = (OP == RED_SUM) ? block_sum(a, sh) : block_max(a, sh);
    if (threadIdx.x == 0) {
      *out    = a;
      *ticket = 0;
      if (mail) mail_post(mail, &a, 1, seq);
    }
  }
}

template <int OP, class F>
inline double run_reduce(int n, F f) {
  if (n <= 0 && !dist_scope()) return 0.0;
  Context& c = ctx();
  int grid = ew_grid(n);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  if (!dist_scope()) {
    // result comes back through the mapped mailbox: no copy, no stream synchronisation
    const unsigned long long seq = ++c.mail_seq;
    reduce_kernel<OP><<<grid, kBlock, 0, c.stream>>>(n, f, c.d_partials, c.d_ticket, c.d_scalar, c.d_mail, seq);
    count_launch();
    if (mail_wait(seq)) return c.h_mail[0];
    return 0.0;
  }
  reduce_kernel<OP><<<grid, kBlock, 0, c.stream>>>(n, f, c.d_partials, c.d_ticket, c.d_scalar, nullptr, 0ull);
  count_launch();
  {
    // row-sharded operand: combine over the ranks -- through peer memory (one tiny kernel that also
    // posts the host mailbox) when available, else NCCL + copy + stream synchronisation
    const unsigned long long seq = c.mail_seq + 1;
    if (dist_p2p_small(c.d_scalar, 1, OP == RED_MAX ? 1u : 0u, 1u, c.d_mail, seq)) {
      c.mail_seq = seq;
      return mail_wait(seq) ? c.h_mail[0] : 0.0;
    }
  }
  dist_allreduce_f64(c.d_scalar, 1, OP == RED_MAX);   // row-sharded operand
  B200_CHECK(cudaMemcpyAsync(c.h_scalar, c.d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  B200_CHECK(cudaStreamSynchronize(c.stream));
  return c.h_scalar[0];
}

__device__ __forceinline__ double dabs(double v) { return v < 0.0 ? -v : v; }

}  // namespace

extern "C" {

// OSQPVectorf_set_scalar (algebra/builtin/vector.c; cuda: vec_set_sc_kernel cuda_lin_alg.cu:52)
void b200_vec_set_scalar(T* a, T sc, int n) {
  launch_ew(n, [=] __device__(int i) { a[i] = sc; });
}

// OSQPVectorf_set_scalar_conditional (cuda: vec_set_sc_cond_kernel cuda_lin_alg.cu:64)
void b200_vec_set_scalar_cond(T* a, const int* test, T neg, T zero, T pos, int n) {
  launch_ew(n, [=] __device__(int i) {
    int t = test[i];
    a[i] = (t == 0) ? zero : (t > 0 ? pos : neg);
  });
}

// OSQPVectorf_round_to_zero (cuda: vec_round_kernel cuda_lin_alg.cu:38)
void b200_vec_round_to_zero(T* a, T tol, int n) {
  launch_ew(n, [=] __device__(int i) {
    T v = a[i];
    if ((v < (T)0 ? -v : v) < tol) a[i] = (T)0;   // strict, like algebra/builtin/vector.c:345-356
  });
}

void b200_vec_mult_scalar(T* a, T sc, int n) {
  launch_ew(n, [=] __device__(int i) { a[i] *= sc; });
}

// OSQPVectorf_add_scaled / plus / minus / copy (algebra/builtin/vector.c; cuda_lin_alg.cu:616-643)
void b200_vec_add_scaled(T* x, T sca, const T* a, T scb, const T* b, int n) {
  launch_ew(n, [=] __device__(int i) { x[i] = sca * a[i] + scb * b[i]; });
}

void b200_vec_add_scaled3(T* x, T sca, const T* a, T scb, const T* b, T scc, const T* c, int n) {
  launch_ew(n, [=] __device__(int i) { x[i] = sca * a[i] + scb * b[i] + scc * c[i]; });
}

void b200_vec_ew_prod(T* c, const T* a, const T* b, int n) {
  launch_ew(n, [=] __device__(int i) { c[i] = a[i] * b[i]; });
}

// OSQPVectorf_ew_bound_vec: x = min(max(z, l), u)  (vector.c:667-681; cuda: vec_bound_kernel :180)
void b200_vec_ew_bound(T* x, const T* z, const T* l, const T* u, int n) {
  launch_ew(n, [=] __device__(int i) {
    T v = z[i], lo = l[i], hi = u[i];
    v = (v > lo) ? v : lo;
    v = (v < hi) ? v : hi;
    x[i] = v;
  });
}

// OSQPVectorf_project_polar_reccone (vector.c:683-708)
void b200_vec_project_polar_reccone(T* y, const T* l, const T* u, T infval, int n) {
  launch_ew(n, [=] __device__(int i) {
    T yi = y[i];
    if (u[i] > +infval) {
      if (l[i] < -infval) yi = (T)0;
      else yi = (yi < (T)0) ? yi : (T)0;
    } else if (l[i] < -infval) {
      yi = (yi > (T)0) ? yi : (T)0;
    }
    y[i] = yi;
  });
}

void b200_vec_ew_reciprocal(T* b, const T* a, int n) {
  launch_ew(n, [=] __device__(int i) { b[i] = (T)1.0 / a[i]; });
}

void b200_vec_ew_sqrt(T* a, int n) {
  launch_ew(n, [=] __device__(int i) { a[i] = sqrt(a[i]); });
}

void b200_vec_ew_max(T* c, const T* a, const T* b, int n) {
  launch_ew(n, [=] __device__(int i) { T x = a[i], y = b[i]; c[i] = (x > y) ? x : y; });
}

void b200_vec_ew_min(T* c, const T* a, const T* b, int n) {
  launch_ew(n, [=] __device__(int i) { T x = a[i], y = b[i]; c[i] = (x < y) ? x : y; });
}

void b200_vec_set_scalar_if_lt(T* x, const T* z, T testval, T newval, int n) {
  launch_ew(n, [=] __device__(int i) { T v = z[i]; x[i] = (v < testval) ? newval : v; });
}

void b200_vec_set_scalar_if_gt(T* x, const T* z, T testval, T newval, int n) {
  launch_ew(n, [=] __device__(int i) { T v = z[i]; x[i] = (v > testval) ? newval : v; });
}

void b200_vec_scatter(T* dst, const T* src, const int* idx, int n) {
  launch_ew(n, [=] __device__(int i) { dst[idx[i]] = src[i]; });
}

void b200_vec_gather(T* dst, const T* src, const int* idx, int n) {
  launch_ew(n, [=] __device__(int i) { dst[i] = src[idx[i]]; });
}

// ------------------------------------------------------------------ reductions
T b200_vec_norm_inf(const T* v, int n) {
  return (T)run_reduce<RED_MAX>(n, [=] __device__(int i) { return dabs((double)v[i]); });
}

// ||S v||_inf  (vector.c:497-513)
T b200_vec_scaled_norm_inf(const T* s, const T* v, int n) {
  return (T)run_reduce<RED_MAX>(n, [=] __device__(int i) { return dabs((double)(s[i] * v[i])); });
}

T b200_vec_norm_inf_diff(const T* a, const T* b, int n) {
  return (T)run_reduce<RED_MAX>(n, [=] __device__(int i) { return dabs((double)(a[i] - b[i])); });
}

T b200_vec_norm_1(const T* v, int n) {
  return (T)run_reduce<RED_SUM>(n, [=] __device__(int i) { return dabs((double)v[i]); });
}

T b200_vec_norm_2(const T* v, int n) {
  double s = run_reduce<RED_SUM>(n, [=] __device__(int i) { double x = v[i]; return x * x; });
  return (T)sqrt(s);
}

T b200_vec_dot(const T* a, const T* b, int n) {
  return (T)run_reduce<RED_SUM>(n, [=] __device__(int i) { return (double)a[i] * (double)b[i]; });
}

// a' max(b,0) / a' min(b,0)  (vector.c:591-617); no per-thread atomics, no cudaMalloc
// (reference: vec_prod_pos/neg_kernel cuda_lin_alg.cu:81-111 + :756-780)
T b200_vec_dot_signed(const T* a, const T* b, int sign, int n) {
  if (sign == 1)
    return (T)run_reduce<RED_SUM>(n, [=] __device__(int i) {
      double bv = b[i];
      return (double)a[i] * (bv > 0.0 ? bv : 0.0);
    });
  if (sign == -1)
    return (T)run_reduce<RED_SUM>(n, [=] __device__(int i) {
      double bv = b[i];
      return (double)a[i] * (bv < 0.0 ? bv : 0.0);
    });
  return b200_vec_dot(a, b, n);
}

// flags are reduced as max over a 0/1 "violation" indicator
int b200_vec_all_leq(const T* l, const T* u, int n) {
  double viol = run_reduce<RED_MAX>(n, [=] __device__(int i) { return (l[i] > u[i]) ? 1.0 : 0.0; });
  return viol > 0.0 ? 0 : 1;
}

// OSQPVectorf_in_reccone (vector.c:710-733)
int b200_vec_in_reccone(const T* y, const T* l, const T* u, T infval, T tol, int n) {
  double viol = run_reduce<RED_MAX>(n, [=] __device__(int i) {
    T yi = y[i];
    bool bad = ((u[i] < +infval) && (yi > +tol)) || ((l[i] > -infval) && (yi < -tol));
    return bad ? 1.0 : 0.0;
  });
  return viol > 0.0 ? 0 : 1;
}

int b200_vec_is_eq(const T* a, const T* b, T tol, int n) {
  double viol = run_reduce<RED_MAX>(n, [=] __device__(int i) {
    double d = (double)a[i] - (double)b[i];
    return (dabs(d) > (double)tol) ? 1.0 : 0.0;
  });
  return viol > 0.0 ? 0 : 1;
}

int b200_veci_is_eq(const int* a, const int* b, int n) {
  double viol = run_reduce<RED_MAX>(n, [=] __device__(int i) { return (a[i] != b[i]) ? 1.0 : 0.0; });
  return viol > 0.0 ? 0 : 1;
}

// OSQPVectorf_ew_bounds_type (vector.c:888-922): loose (-1) tested first, then equality (1)
int b200_vec_bounds_type(int* iseq, const T* l, const T* u, T tol, T infval, int n) {
  double changed = run_reduce<RED_MAX>(n, [=] __device__(int i) {
    int old = iseq[i], nv;
    T lo = l[i], hi = u[i];
    if ((lo < -infval) && (hi > infval)) nv = -1;
    else if (hi - lo < tol) nv = 1;
    else nv = 0;
    iseq[i] = nv;
    return (nv != old) ? 1.0 : 0.0;
  });
  return changed > 0.0 ? 1 : 0;
}

// ------------------------------------------------------------- fused termination-check reductions
namespace {

struct ResArgs {
  const T *x, *y, *z, *Ax, *Px, *Aty, *q, *l, *u, *Einv, *Dinv;
  T infval, deadzone;
  int n, m;
  int n_off;     // column-split layout: first n-entry this rank counts (shared slice: rank 0 only)
};

__device__ __forceinline__ bool res_is_sum(int s) {
  return s == B200_RES_SC || s == B200_RES_XPX || s == B200_RES_QX;
}

__global__ void __launch_bounds__(kBlock) residuals_kernel(ResArgs a, double* partials, unsigned* ticket,
                                                           double* out, double* mail, unsigned long long seq) {
  __shared__ double shw[B200_RES_COUNT][kBlock / 32];
  __shared__ bool is_last;
  double v[B200_RES_COUNT];
#pragma unroll
  for (int s = 0; s < B200_RES_COUNT; s++) v[s] = 0.0;
  const int stride = gridDim.x * blockDim.x;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = gtid; j < a.m; j += stride) {
    const double Ax = a.Ax[j], z = a.z[j], e = a.Einv ? (double)a.Einv[j] : 1.0;
    const double d = Ax - z;
    v[B200_RES_PRIM_S] = fmax(v[B200_RES_PRIM_S], dabs(d));
    v[B200_RES_PRIM_U] = fmax(v[B200_RES_PRIM_U], dabs(e * d));
    v[B200_RES_Z_S]    = fmax(v[B200_RES_Z_S], dabs(z));
    v[B200_RES_Z_U]    = fmax(v[B200_RES_Z_U], dabs(e * z));
    v[B200_RES_AX_S]   = fmax(v[B200_RES_AX_S], dabs(Ax));
    v[B200_RES_AX_U]   = fmax(v[B200_RES_AX_U], dabs(e * Ax));
    // support function of [l, u] at y projected on the polar recession cone and dead-zoned
    // (compute_obj_val_dual_gap, auxil.c:245-259)
    const double lo = a.l[j], hi = a.u[j];
    double yp = a.y[j];
    if (hi > +(double)a.infval) {
      if (lo < -(double)a.infval) yp = 0.0;
      else yp = (yp < 0.0) ? yp : 0.0;
    } else if (lo < -(double)a.infval) {
      yp = (yp > 0.0) ? yp : 0.0;
    }
    if (dabs(yp) < (double)a.deadzone) yp = 0.0;
    v[B200_RES_SC] += hi * (yp > 0.0 ? yp : 0.0) + lo * (yp < 0.0 ? yp : 0.0);
  }
  for (int i = a.n_off + gtid; i < a.n; i += stride) {
    const double q = a.q[i], Px = a.Px[i], Aty = a.m > 0 ? (double)a.Aty[i] : 0.0;
    const double dv = a.Dinv ? (double)a.Dinv[i] : 1.0, x = a.x[i];
    double r = q + Px;
    if (a.m > 0) r += Aty;
    v[B200_RES_DUAL_S] = fmax(v[B200_RES_DUAL_S], dabs(r));
    v[B200_RES_DUAL_U] = fmax(v[B200_RES_DUAL_U], dabs(dv * r));
    v[B200_RES_Q_S]    = fmax(v[B200_RES_Q_S], dabs(q));
    v[B200_RES_Q_U]    = fmax(v[B200_RES_Q_U], dabs(dv * q));
    v[B200_RES_PX_S]   = fmax(v[B200_RES_PX_S], dabs(Px));
    v[B200_RES_PX_U]   = fmax(v[B200_RES_PX_U], dabs(dv * Px));
    v[B200_RES_ATY_S]  = fmax(v[B200_RES_ATY_S], dabs(Aty));
    v[B200_RES_ATY_U]  = fmax(v[B200_RES_ATY_U], dabs(dv * Aty));
    v[B200_RES_XPX] += Px * x;
    v[B200_RES_QX]  += q * x;
  }
  // CTA partials with ONE barrier: warp shuffles, per-warp results in shared memory, then thread s
  // folds slot s over the warps
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < B200_RES_COUNT; s++) {
    const double r = res_is_sum(s) ? warp_sum(v[s]) : warp_max(v[s]);
    if (lane == 0) shw[s][w] = r;
  }
  __syncthreads();
  if (threadIdx.x < B200_RES_COUNT) {
    const int s = threadIdx.x;
    double acc = 0.0;
    for (int k = 0; k < kBlock / 32; k++) acc = res_is_sum(s) ? acc + shw[s][k] : fmax(acc, shw[s][k]);
    partials[s * gridDim.x + blockIdx.x] = acc;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // warp w folds slots w, w + 8, ... in index order; no CTA barriers
    for (int s = w; s < B200_RES_COUNT; s += kBlock / 32) {
      double acc = 0.0;
      for (int b = lane; b < gridDim.x; b += 32) {
        const double p = __ldcg(&partials[s * gridDim.x + b]);
        acc = res_is_sum(s) ? acc + p : fmax(acc, p);
      }
      acc = res_is_sum(s) ? warp_sum(acc) : warp_max(acc);
      if (lane == 0) { out[s] = acc; shw[s][0] = acc; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *ticket = 0;
      if (mail) {
        double vals[B200_RES_COUNT];
        for (int s = 0; s < B200_RES_COUNT; s++) vals[s] = shw[s][0];
        mail_post(mail, vals, B200_RES_COUNT, seq);
      }
    }
  }
}

}  // namespace

extern "C" void b200_admm_residuals(const T* x, const T* y, const T* z, const T* Ax, const T* Px,
                                    const T* Aty, const T* q, const T* l, const T* u, const T* Einv,
                                    const T* Dinv, T infval, T deadzone, int n, int m, double* h_out) {
  Context& c = ctx();
  ResArgs a{x, y, z, Ax, Px, Aty, q, l, u, Einv, Dinv, infval, deadzone, n, m, dist_col_off()};
  const int nm = n > m ? n : m;
  int grid = ew_grid(nm);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  if (!dist_active()) {
    const unsigned long long seq = ++c.mail_seq;
    residuals_kernel<<<grid, kBlock, 0, c.stream>>>(a, c.d_partials, c.d_ticket, c.d_scalar, c.d_mail, seq);
    count_launch();
    if (mail_wait(seq)) {
      for (int s = 0; s < B200_RES_COUNT; s++) h_out[s] = c.h_mail[s];
      return;
    }
  }
  residuals_kernel<<<grid, kBlock, 0, c.stream>>>(a, c.d_partials, c.d_ticket, c.d_scalar, nullptr, 0ull);
  count_launch();
  if (dist_active()) {
    // slots 0..5 are maxima over the row-sharded m-vectors, slot 6 (support function) a sum over
    // them; the n-vector slots are replicated and identical on every rank (plain row layout) or partial
    // too (column-split layout).  One peer-memory exchange of the 17 scalars when available ...
    unsigned max_mask = 0, active = 0;
    for (int s = 0; s < B200_RES_COUNT; s++) {
      if (!(s == B200_RES_SC || s == B200_RES_XPX || s == B200_RES_QX)) max_mask |= 1u << s;
      if (s <= B200_RES_SC || dist_split()) active |= 1u << s;
    }
    const unsigned long long seq = c.mail_seq + 1;
    if (dist_p2p_small(c.d_scalar, B200_RES_COUNT, max_mask, active, c.d_mail, seq)) {
      c.mail_seq = seq;
      if (mail_wait(seq)) {
        for (int s = 0; s < B200_RES_COUNT; s++) h_out[s] = c.h_mail[s];
        return;
      }
    }
    // ... else four NCCL all-reduces
    dist_allreduce_f64(c.d_scalar + B200_RES_PRIM_S, B200_RES_SC - B200_RES_PRIM_S, true);
    dist_allreduce_f64(c.d_scalar + B200_RES_SC, 1, false);
    if (dist_split()) {
      // column-split layout: the n-vector slots are partial too (maxima, then the two sums)
      dist_allreduce_f64(c.d_scalar + B200_RES_DUAL_S, B200_RES_XPX - B200_RES_DUAL_S, true);
      dist_allreduce_f64(c.d_scalar + B200_RES_XPX, 2, false);
    }
  }
  B200_CHECK(cudaMemcpyAsync(c.h_scalar, c.d_scalar, sizeof(double) * B200_RES_COUNT,
                             cudaMemcpyDeviceToHost, c.stream));
  B200_CHECK(cudaStreamSynchronize(c.stream));
  for (int s = 0; s < B200_RES_COUNT; s++) h_out[s] = c.h_scalar[s];
}

// ------------------------------------------------------------- infeasibility pre-tests
namespace {
struct InfArgs {
  T* dy; const T *l, *u, *E, *dx, *D, *q;
  T infval;
  int n, m, n_off, do_primal, do_dual;
};
// 0: ||E .* proj(dy)||_inf   1: u' max(proj(dy), 0)   2: l' min(proj(dy), 0)   3: ||D .* dx||_inf   4: q' dx
__global__ void __launch_bounds__(kBlock) infeas_kernel(InfArgs a, double* partials, unsigned* ticket, double* out,
                                                        double* mail, unsigned long long seq) {
  __shared__ double shw[5][kBlock / 32];
  __shared__ bool is_last;
  double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int stride = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  if (a.do_primal) {
    for (int j = gtid; j < a.m; j += stride) {
      // project_polar_reccone in place, exactly as b200_vec_project_polar_reccone
      T yi = a.dy[j];
      const T lo = a.l[j], hi = a.u[j];
      if (hi > +a.infval) {
        if (lo < -a.infval) yi = (T)0;
        else yi = (yi < (T)0) ? yi : (T)0;
      } else if (lo < -a.infval) {
        yi = (yi > (T)0) ? yi : (T)0;
      }
      a.dy[j] = yi;
      const double y = yi, e = a.E ? (double)a.E[j] : 1.0;
      v[0] = fmax(v[0], dabs(e * y));
      v[1] += (double)hi * (y > 0.0 ? y : 0.0);
      v[2] += (double)lo * (y < 0.0 ? y : 0.0);
    }
  }
  if (a.do_dual) {
    for (int i = a.n_off + gtid; i < a.n; i += stride) {
      const double x = a.dx[i], d = a.D ? (double)a.D[i] : 1.0;
      v[3] = fmax(v[3], dabs(d * x));
      v[4] += (double)a.q[i] * x;
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < 5; s++) {
    const bool is_max = (s == 0 || s == 3);
    const double r = is_max ? warp_max(v[s]) : warp_sum(v[s]);
    if (lane == 0) shw[s][w] = r;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    const int s = threadIdx.x;
    const bool is_max = (s == 0 || s == 3);
    double acc = 0.0;
    for (int k = 0; k < kBlock / 32; k++) acc = is_max ? fmax(acc, shw[s][k]) : acc + shw[s][k];
    partials[s * gridDim.x + blockIdx.x] = acc;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (w < 5) {
      const int s = w;
      const bool is_max = (s == 0 || s == 3);
      double acc = 0.0;
      for (int b = lane; b < gridDim.x; b += 32) {
        const double p = __ldcg(&partials[s * gridDim.x + b]);
        acc = is_max ? fmax(acc, p) : acc + p;
      }
      acc = is_max ? warp_max(acc) : warp_sum(acc);
      if (lane == 0) { out[s] = acc; shw[s][0] = acc; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *ticket = 0;
      if (mail) {
        double vals[5];
        for (int s = 0; s < 5; s++) vals[s] = shw[s][0];
        mail_post(mail, vals, 5, seq);
      }
    }
  }
}
}  // namespace

// The scalars that decide whether is_primal_infeasible / is_dual_infeasible (src/auxil.c:460-585) can
// return non-zero at all, in ONE kernel and ONE host round trip instead of five blocking reductions:
// h_out = { ||E .* dy||_inf, u' max(dy, 0), l' min(dy, 0), ||D .* dx||_inf, q' dx } with dy projected on the
// polar of the recession cone IN PLACE (as the reference does first).  E / D may be NULL (no scaling or
// scaled termination).  Row-sharded: combined over the ranks (rows: 0-2; columns: 3-4 when column-split).
extern "C" void b200_admm_infeas_scalars(T* dy, const T* l, const T* u, const T* E, const T* dx, const T* D,
                                         const T* q, T infval, int n, int m, int do_primal, int do_dual,
                                         double* h_out) {
  Context& c = ctx();
  InfArgs a{dy, l, u, E, dx, D, q, infval, n, m, dist_col_off(), do_primal, do_dual};
  const int nm = n > m ? n : m;
  int grid = ew_grid(nm);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  if (!dist_active()) {
    const unsigned long long seq = ++c.mail_seq;
    infeas_kernel<<<grid, kBlock, 0, c.stream>>>(a, c.d_partials, c.d_ticket, c.d_scalar, c.d_mail, seq);
    count_launch();
    if (mail_wait(seq)) {
      for (int s = 0; s < 5; s++) h_out[s] = c.h_mail[s];
      return;
    }
    for (int s = 0; s < 5; s++) h_out[s] = 0.0;
    return;
  }
  infeas_kernel<<<grid, kBlock, 0, c.stream>>>(a, c.d_partials, c.d_ticket, c.d_scalar, nullptr, 0ull);
  count_launch();
  const unsigned active = dist_split() ? 0x1Fu : 0x07u;
  const unsigned long long seq = c.mail_seq + 1;
  if (dist_p2p_small(c.d_scalar, 5, 0x09u, active, c.d_mail, seq)) {
    c.mail_seq = seq;
    if (mail_wait(seq)) {
      for (int s = 0; s < 5; s++) h_out[s] = c.h_mail[s];
      return;
    }
  }
  for (int s = 0; s < 5; s++)
    if ((active >> s) & 1u) dist_allreduce_f64(c.d_scalar + s, 1, s == 0 || s == 3);
  B200_CHECK(cudaMemcpyAsync(c.h_scalar, c.d_scalar, sizeof(double) * 5, cudaMemcpyDeviceToHost, c.stream));
  B200_CHECK(cudaStreamSynchronize(c.stream));
  for (int s = 0; s < 5; s++) h_out[s] = c.h_scalar[s];
}

// ------------------------------------------------------------- fused ADMM steps
// compute_rhs (src/auxil.c:136-158): x~ = sigma x_prev - q ; z~ = z_prev - rho^-1 y
void b200_admm_compute_rhs(T* xt, T* zt, const T* x_prev, const T* q, const T* z_prev, const T* y,
                           const T* rho_inv_vec, T rho_inv, T sigma, int n, int m) {
  const int tot = n + m;
  launch_ew(tot, [=] __device__(int i) {
    if (i < n) {
      xt[i] = sigma * x_prev[i] - q[i];
    } else {
      int j = i - n;
      T ri = rho_inv_vec ? rho_inv_vec[j] : rho_inv;
      zt[j] = z_prev[j] - ri * y[j];
    }
  });
}

// update_x, update_z, update_y (src/auxil.c:172-229) in one pass:
//   x  = alpha x~ + (1-alpha) x_prev ;              dx = x - x_prev
//   z  = clip(alpha z~ + (1-alpha) z_prev + y/rho, l, u)
//   dy = rho (alpha z~ + (1-alpha) z_prev - z) ;    y += dy
// expression order follows the reference's add_scaled / add_scaled3 calls.
void b200_admm_update_xzy_carry(T* x, T* dx, T* z, T* y, T* dy, const T* xt, const T* zt, const T* x_prev,
                                const T* z_prev, const T* l, const T* u, const T* rho_vec,
                                const T* rho_inv_vec, T rho, T rho_inv, T alpha, int n, int m, T* Ax);

void b200_admm_update_xzy(T* x, T* dx, T* z, T* y, T* dy, const T* xt, const T* zt, const T* x_prev,
                          const T* z_prev, const T* l, const T* u, const T* rho_vec,
                          const T* rho_inv_vec, T rho, T rho_inv, T alpha, int n, int m) {
  b200_admm_update_xzy_carry(x, dx, z, y, dy, xt, zt, x_prev, z_prev, l, u, rho_vec, rho_inv_vec, rho, rho_inv,
                             alpha, n, m, nullptr);
}

// Same, and additionally carries the product A x through the relaxation step when Ax != NULL:
// x+ = alpha x~ + (1 - alpha) x  and  z~ = A x~ (what the linear solve returns), so
// A x+ = alpha z~ + (1 - alpha) A x  -- the termination check then needs no SpMV for A x
// (SURVEY.md section 8f.1).  Ax must hold A x_prev on entry.
void b200_admm_update_xzy_carry(T* x, T* dx, T* z, T* y, T* dy, const T* xt, const T* zt, const T* x_prev,
                                const T* z_prev, const T* l, const T* u, const T* rho_vec,
                                const T* rho_inv_vec, T rho, T rho_inv, T alpha, int n, int m, T* Ax) {
  const int tot = n + m;
  const T oma = (T)1.0 - alpha;
  launch_ew(tot, [=] __device__(int i) {
    if (i < n) {
      T xp = x_prev[i];
      T xn = alpha * xt[i] + oma * xp;
      x[i]  = xn;
      dx[i] = xn - xp;
    } else {
      int j = i - n;
      T zti = zt[j], zp = z_prev[j], yj = y[j];
      T zn, d;
      if (rho_vec) {
        // rho_is_vec branch of update_z / update_y (auxil.c:192-198, 220-222)
        zn = (T)1.0 * (rho_inv_vec[j] * yj) + alpha * zti + oma * zp;
      } else {
        zn = alpha * zti + oma * zp + rho_inv * yj;
      }
      T lo = l[j], hi = u[j];
      zn = (zn > lo) ? zn : lo;
      zn = (zn < hi) ? zn : hi;
      d = alpha * zti + oma * zp + (T)(-1.0) * zn;
      d = rho_vec ? d * rho_vec[j] : d * rho;
      z[j]  = zn;
      dy[j] = d;
      y[j]  = yj + d;
      if (Ax) Ax[j] = alpha * zti + oma * Ax[j];
    }
  });
}

}  // extern "C"
