// batch.cu -- batches of small independent QPs that share P and A (BASELINE configs[4]: 4096 MPC QPs with
// n = 204, m = 360, differing in their bounds): ONE CTA PER QP runs the WHOLE osqp_solve loop -- compute_rhs,
// the reduced-KKT Jacobi PCG, update_x / update_z / update_y, update_info, check_termination, adaptive rho --
// with every iterate in shared memory.  The per-op private interface cannot get there: through launches a
// 204-variable QP costs ~90 us per ADMM iteration of pure launch latency (profiles/r01_batch_mpc_threads.txt:
// 505 QPs/s on a whole B200); here an iteration is a few dozen CTA barriers and hundreds of QPs are resident.
//
// What is reproduced, with the reference line it follows (all on the SCALED problem of the template solver:
// the batch shares P, A, hence the Ruiz scaling D, E and -- q being shared too unless given -- the cost
// scaling c):
//   ADMM loop, check/adapt cadence        src/osqp_api.c:696-903
//   compute_rhs, update_x/z/y             src/auxil.c:136-229
//   PCG + tolerance schedule              algebra/cuda/lin_sys/indirect/cuda_pcg.cu:113-208,
//                                         cuda_pcg_interface.cu:32-92,229-273 (as osqp_b200/csrc/pcg.cu)
//   update_info, residuals, objective     src/auxil.c:231-410,676-762
//   check_termination, tolerances         src/auxil.c:334-458,808-945 (solved / solved inaccurate / max iter /
//                                         non-convex; the infeasibility certificates are NOT evaluated: a QP that
//                                         ends in "maximum iterations reached" here is re-solved through the
//                                         ordinary API by the caller)
//   compute_rho_estimate, adapt_rho       src/auxil.c:14-74, osqp_update_rho src/osqp_api.c:1412-1462
//   constraint classification, rho_vec    src/auxil.c:76-105 (per QP, from its own bounds)
//   unscale_solution                      src/scaling.c:194-208
#include "csr.cuh"

#include <cstring>

using namespace b200;

namespace {

constexpr int kBB = 128;               // threads per CTA (= per QP)
#ifndef B200_BATCH_MINBLOCKS
#define B200_BATCH_MINBLOCKS 4         /* resident CTAs per SM the register allocation must allow (128 regs at 4) */
#endif
constexpr int kBW = kBB / 32;
#ifdef B200_USE_FLOAT
constexpr double kInfty = 1e17;        // OSQP_INFTY of a float build (osqp_api_constants.h:196-203)
#else
constexpr double kInfty = 1e30;        // OSQP_INFTY
#endif
constexpr double kMinScaling = 1e-4;   // OSQP_MIN_SCALING
constexpr double kDivTol = 1.0 / kInfty;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoTol = 1e-4, kRhoEqOverIneq = 1e3;
constexpr double kCgTolMinB = 1e-7;
#ifdef B200_USE_FLOAT
constexpr double kDeadzone = 1e-10;    // OSQP_ZERO_DEADZONE (float)
#else
constexpr double kDeadzone = 1e-15;
#endif

struct BatchArgs {
  int n, m, nb;
  const int *Prp, *Pci; const T* Pv;       // full symmetric P (scaled), structurally full diagonal
  const int *Arp, *Aci; const T* Av;       // A (scaled), CSR
  const int *Atrp, *Atci; const T* Atv;    // A' (scaled), CSR
  const T* q;                               // scaled q shared by the batch (q_batch == nullptr)
  const T* q_batch;                         // nb x n, user (unscaled) values, or nullptr
  const T* l_batch; const T* u_batch;       // nb x m, user (unscaled) values
  const T *D, *Dinv, *E, *Einv;             // scaling vectors, nullptr when scaling is off
  T c, cinv;
  T rho0, sigma, alpha, eps_abs, eps_rel, adaptive_rho_tolerance;
  int rho_is_vec, max_iter, check_termination, adaptive_rho, adaptive_rho_interval, check_dualgap, scaled_termination;
  int cg_max_iter, cg_tol_reduction;
  double cg_tol_fraction;
  T* x_out; T* y_out;                       // nb x n, nb x m (unscaled)
  int* iters; int* status; int* cg_iters; int* rho_updates;
  T* obj; T* prim_res; T* dual_res;
};

// K values reduced over the CTA at once (sum or max per slot); result valid in ALL threads
template <int K>
__device__ __forceinline__ void block_reduce(double (&v)[K], unsigned max_mask, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) v[k] = ((max_mask >> k) & 1u) ? warp_max(v[k]) : warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) sh[k * kBW + w] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    double a = sh[k * kBW];
#pragma unroll
    for (int ww = 1; ww < kBW; ww++) a = ((max_mask >> k) & 1u) ? fmax(a, sh[k * kBW + ww]) : a + sh[k * kBW + ww];
    v[k] = a;
  }
  __syncthreads();
}

__device__ __forceinline__ T row_dot(const int* __restrict__ rp, const int* __restrict__ ci, const T* __restrict__ v,
                                     const T* src, int row) {
  T s = 0;
  const int e = __ldg(rp + row + 1);
  for (int k = __ldg(rp + row); k < e; k++) s += __ldg(v + k) * src[__ldg(ci + k)];
  return s;
}

__global__ void __launch_bounds__(kBB, B200_BATCH_MINBLOCKS) batch_admm_kernel(BatchArgs a) {
  extern __shared__ __align__(16) unsigned char dsm_raw[];
  const int n = a.n, m = a.m, tid = threadIdx.x;
  T* base = reinterpret_cast<T*>(dsm_raw);
  T* xa = base;        T* xb = xa + n;   T* xcg = xb + n;  T* r = xcg + n;  T* p = r + n;
  T* Kp = p + n;       T* minv = Kp + n; T* qs = minv + n;
  T* za = qs + n;      T* zb = za + m;   T* y = zb + m;    T* t = y + m;    T* w = t + m;
  T* l = w + m;        T* u = l + m;     T* rhov = u + m;
  double* sh = reinterpret_cast<double*>(rhov + m + (((size_t)(8 * n + 8 * m)) & 1));   // 8-byte aligned scratch
  const T infval = (T)(kInfty * kMinScaling);
  const bool scaling = a.D != nullptr;
  const bool unscale = scaling && !a.scaled_termination;

  for (int qp = blockIdx.x; qp < a.nb; qp += gridDim.x) {
    // ---------------------------------------------------------------- data of this QP, scaled
    for (int j = tid; j < m; j += kBB) {
      const T e = scaling ? a.E[j] : (T)1;
      T lo = a.l_batch[(size_t)qp * m + j], hi = a.u_batch[(size_t)qp * m + j];
      lo = fmax(fmin(lo, (T)kInfty), (T)-kInfty) * e;
      hi = fmax(fmin(hi, (T)kInfty), (T)-kInfty) * e;
      l[j] = lo; u[j] = hi;
      za[j] = 0; zb[j] = 0; y[j] = 0;
    }
    for (int i = tid; i < n; i += kBB) {
      qs[i] = a.q_batch ? a.c * (scaling ? a.D[i] : (T)1) * a.q_batch[(size_t)qp * n + i] : a.q[i];
      xa[i] = 0; xb[i] = 0; xcg[i] = 0;
    }
    __syncthreads();
    T rho = a.rho0;
    // rho_vec from this QP's own constraint types (set_rho_vec: loose tested before equality)
    auto set_rho = [&]() {
      for (int j = tid; j < m; j += kBB) {
        T rj = rho;
        if (a.rho_is_vec) {
          if (l[j] < -infval && u[j] > infval) rj = (T)kRhoMin;
          else if (u[j] - l[j] < (T)kRhoTol) rj = (T)kRhoEqOverIneq * rho;
        }
        rhov[j] = rj;
      }
      __syncthreads();
      for (int i = tid; i < n; i += kBB) {
        T s = 0, pd = 0;
        const int e1 = __ldg(a.Atrp + i + 1);
        for (int k = __ldg(a.Atrp + i); k < e1; k++) { const T v = __ldg(a.Atv + k); s += v * v * rhov[__ldg(a.Atci + k)]; }
        const int e0 = __ldg(a.Prp + i + 1);
        for (int k = __ldg(a.Prp + i); k < e0; k++) if (__ldg(a.Pci + k) == i) pd = __ldg(a.Pv + k);
        minv[i] = (T)1 / (a.sigma + pd + s);
      }
      __syncthreads();
    };
    set_rho();

    T* x = xa; T* xp = xb; T* z = za; T* zp = zb;
    double rf = a.cg_tol_fraction, eps_prev = 1.0;
    int zero_iters = 0;
    double spr = 0.0, sdr = 0.0;       // scaled primal / dual residual of the last update_info
    int status = 0, iter = 0, cg_total = 0, n_rho = 0;
    bool fresh = false;                // info of the current iterate available
    double R[17];
    double prim_res = 0, dual_res = 0, obj = 0, sdg = 0;

    // update_info (auxil.c:676-762): A x -> w, P x -> Kp, A'y -> r (the CG work vectors are free here)
    auto info = [&]() {
      for (int j = tid; j < m; j += kBB) w[j] = row_dot(a.Arp, a.Aci, a.Av, x, j);
      for (int i = tid; i < n; i += kBB) {
        Kp[i] = row_dot(a.Prp, a.Pci, a.Pv, x, i);
        r[i]  = row_dot(a.Atrp, a.Atci, a.Atv, y, i);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 17; k++) R[k] = 0.0;
      for (int j = tid; j < m; j += kBB) {
        const double ei = scaling ? (double)a.Einv[j] : 1.0;
        const double ax = w[j], zz = z[j], d = fabs(ax - zz);
        R[0] = fmax(R[0], d);            R[1] = fmax(R[1], ei * d);
        R[2] = fmax(R[2], fabs(zz));     R[3] = fmax(R[3], ei * fabs(zz));
        R[4] = fmax(R[4], fabs(ax));     R[5] = fmax(R[5], ei * fabs(ax));
        // support function of [l, u] at y projected on the polar of the recession cone (auxil.c:246-259)
        double yj = y[j];
        if (u[j] > infval) { if (l[j] < -infval) yj = 0.0; else yj = fmin(yj, 0.0); }
        else if (l[j] < -infval) yj = fmax(yj, 0.0);
        if (fabs(yj) < kDeadzone) yj = 0.0;
        R[16] += (double)u[j] * fmax(yj, 0.0) + (double)l[j] * fmin(yj, 0.0);
      }
      for (int i = tid; i < n; i += kBB) {
        const double di = scaling ? (double)a.Dinv[i] : 1.0;
        const double px = Kp[i], aty = r[i], qq = qs[i], d = fabs(qq + px + aty);
        R[6]  = fmax(R[6], d);           R[7]  = fmax(R[7], di * d);
        R[8]  = fmax(R[8], fabs(qq));    R[9]  = fmax(R[9], di * fabs(qq));
        R[10] = fmax(R[10], fabs(px));   R[11] = fmax(R[11], di * fabs(px));
        R[12] = fmax(R[12], fabs(aty));  R[13] = fmax(R[13], di * fabs(aty));
        R[14] += (double)x[i] * px;      R[15] += qq * (double)x[i];
      }
      block_reduce<17>(R, 0x3FFFu, sh);
      spr = R[0]; sdr = R[6];
      prim_res = (m == 0) ? 0.0 : (unscale ? R[1] : R[0]);
      dual_res = unscale ? (double)a.cinv * R[7] : R[6];
      sdg = R[14] + R[15] + R[16];
      obj = (0.5 * R[14] + R[15]) * (scaling ? (double)a.cinv : 1.0);
      fresh = true;
    };
    // check_termination (auxil.c:808-945) without the infeasibility branches; returns the status or 0
    auto check = [&](double mult) -> int {
      if (prim_res > kInfty || dual_res > kInfty) return 9;    // OSQP_NON_CVX
      const double ea = mult * (double)a.eps_abs, er = mult * (double)a.eps_rel;
      bool pc = true;
      if (m > 0) pc = prim_res < ea + er * (unscale ? fmax(R[3], R[5]) : fmax(R[2], R[4]));
      const double dn = unscale ? (double)a.cinv * fmax(fmax(R[9], R[13]), R[11]) : fmax(fmax(R[8], R[12]), R[10]);
      const bool dc = dual_res < ea + er * dn;
      bool gc = true;
      if (a.check_dualgap) {
        const double mx = fmax(fmax(fabs(R[14]), fabs(R[15])), fabs(R[16])) * (unscale ? (double)a.cinv : 1.0);
        const double gap = unscale ? fabs((double)a.cinv * sdg) : fabs(sdg);
        gc = gap < ea + er * mx;
      }
      return (pc && dc && gc) ? (mult > 1.0 ? 2 : 1) : 0;
    };

    for (iter = 1; iter <= a.max_iter; iter++) {
      { T* s_ = x; x = xp; xp = s_; s_ = z; z = zp; zp = s_; }
      fresh = false;
      // ---- compute_rhs + reduced right-hand side: rhs = sigma x_prev - q + A'(rho .* (z_prev - y / rho))
      for (int j = tid; j < m; j += kBB) t[j] = rhov[j] * (zp[j] - ((T)1 / rhov[j]) * y[j]);
      __syncthreads();
      double rhs_max = 0.0;
      for (int i = tid; i < n; i += kBB) {
        const T b = (a.sigma * xp[i] - qs[i]) + row_dot(a.Atrp, a.Atci, a.Atv, t, i);
        Kp[i] = b;                                    // rhs parked in Kp until the first K p
        rhs_max = fmax(rhs_max, fabs((double)b));
      }
      __syncthreads();
      // ---- tolerance schedule (cuda_pcg_interface.cu:32-64)
      double eps;
      if (iter == 1) {
        double v1[1] = {rhs_max};
        block_reduce<1>(v1, 1u, sh);
        rf = a.cg_tol_fraction;
        eps_prev = (v1[0] < kCgTolMinB) ? 1.0 : v1[0] * rf;
        eps = eps_prev;
      } else {
        if (zero_iters >= a.cg_tol_reduction) { rf *= 0.5; zero_iters = 0; }
        eps = rf * sqrt(spr * sdr);
        eps = fmax(fmin(eps, eps_prev), kCgTolMinB);
        eps_prev = eps;
      }
      // ---- initial residual r = K xcg - rhs, p = -M^-1 r
      for (int j = tid; j < m; j += kBB) t[j] = rhov[j] * row_dot(a.Arp, a.Aci, a.Av, xcg, j);
      __syncthreads();
      double v2[2] = {0.0, 0.0};
      for (int i = tid; i < n; i += kBB) {
        const T kx = row_dot(a.Prp, a.Pci, a.Pv, xcg, i) + a.sigma * xcg[i] + row_dot(a.Atrp, a.Atci, a.Atv, t, i);
        const T rr = kx - Kp[i];
        const T yy = minv[i] * rr;
        r[i] = rr; p[i] = -yy;
        v2[0] += (double)rr * (double)yy;
        v2[1] = fmax(v2[1], fabs((double)rr));
      }
      block_reduce<2>(v2, 2u, sh);
      double rTy = v2[0], rnorm = v2[1];
      int it = 0;
      while (rnorm > eps && it < a.cg_max_iter) {
        for (int j = tid; j < m; j += kBB) t[j] = rhov[j] * row_dot(a.Arp, a.Aci, a.Av, p, j);
        __syncthreads();
        double v1[1] = {0.0};
        for (int i = tid; i < n; i += kBB) {
          const T kp = row_dot(a.Prp, a.Pci, a.Pv, p, i) + a.sigma * p[i] + row_dot(a.Atrp, a.Atci, a.Atv, t, i);
          Kp[i] = kp;
          v1[0] += (double)p[i] * (double)kp;
        }
        block_reduce<1>(v1, 0u, sh);
        const T al = (T)(rTy / v1[0]);
        v2[0] = 0.0; v2[1] = 0.0;
        for (int i = tid; i < n; i += kBB) {
          xcg[i] += al * p[i];
          const T rr = r[i] + al * Kp[i];
          r[i] = rr;
          const T yy = minv[i] * rr;
          v2[0] += (double)rr * (double)yy;
          v2[1] = fmax(v2[1], fabs((double)rr));
        }
        block_reduce<2>(v2, 2u, sh);
        const T be = (T)(v2[0] / rTy);
        rTy = v2[0]; rnorm = v2[1];
        for (int i = tid; i < n; i += kBB) p[i] = be * p[i] - minv[i] * r[i];
        __syncthreads();
        it++;
      }
      zero_iters = (it == 0) ? zero_iters + 1 : 0;
      cg_total += it;
      // ---- z~ = A x~ ; update_x, update_z, update_y
      const T oma = (T)1 - a.alpha;
      for (int j = tid; j < m; j += kBB) {
        const T zt = row_dot(a.Arp, a.Aci, a.Av, xcg, j);
        const T rj = rhov[j], yj = y[j], zpj = zp[j];
        T zn = a.rho_is_vec ? (T)1 * (((T)1 / rj) * yj) + a.alpha * zt + oma * zpj
                            : a.alpha * zt + oma * zpj + ((T)1 / rj) * yj;
        zn = (zn > l[j]) ? zn : l[j];
        zn = (zn < u[j]) ? zn : u[j];
        const T d = (a.alpha * zt + oma * zpj + (T)(-1) * zn) * rj;
        z[j] = zn;
        y[j] = yj + d;
      }
      for (int i = tid; i < n; i += kBB) x[i] = a.alpha * xcg[i] + oma * xp[i];
      __syncthreads();
      // ---- info / termination / rho (osqp_api.c:767-889)
      const bool can_check = a.check_termination && (iter % a.check_termination == 0);
      const bool can_adapt = a.adaptive_rho == 1 && a.adaptive_rho_interval && (iter % a.adaptive_rho_interval == 0);
      if (can_check || can_adapt || iter == 1) info();
      if (can_check) {
        status = check(1.0);
        if (status) break;
      }
      if (can_adapt) {
        const double pn = spr / (fmax(R[2], R[4]) + kDivTol);
        const double dn = sdr / (fmax(fmax(R[8], R[12]), R[10]) + kDivTol);
        double rn = (double)rho * sqrt(pn / dn);
        rn = fmin(fmax(rn, kRhoMin), kRhoMax);
        if (rn > (double)rho * (double)a.adaptive_rho_tolerance || rn < (double)rho / (double)a.adaptive_rho_tolerance) {
          rho = (T)rn;
          n_rho++;
          set_rho();
        }
      }
    }
    if (iter > a.max_iter) iter = a.max_iter;
    if (!status) {                       // osqp_api.c:905-951
      if (!fresh) info();
      status = check(1.0);
      if (!status) status = check(10.0);
      if (!status) status = 7;           // OSQP_MAX_ITER_REACHED
    }
    // ---- solution (unscale_solution, scaling.c:194-208) and info of this QP
    for (int i = tid; i < n; i += kBB) a.x_out[(size_t)qp * n + i] = scaling ? x[i] * a.D[i] : x[i];
    for (int j = tid; j < m; j += kBB) a.y_out[(size_t)qp * m + j] = scaling ? y[j] * a.E[j] * a.cinv : y[j];
    if (tid == 0) {
      a.iters[qp] = iter; a.status[qp] = status; a.cg_iters[qp] = cg_total; a.rho_updates[qp] = n_rho;
      a.obj[qp] = (T)obj; a.prim_res[qp] = (T)prim_res; a.dual_res[qp] = (T)dual_res;
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" {

// see include/osqp_b200.h
int b200_batch_solve(const b200_csr* P, const b200_csr* A, const b200_csr* At, int n, int m, int nb,
                     const b200_float* d_q, const b200_float* d_q_batch, const b200_float* d_l_batch,
                     const b200_float* d_u_batch, const b200_float* d_D, const b200_float* d_Dinv,
                     const b200_float* d_E, const b200_float* d_Einv, b200_float c, b200_float cinv,
                     const b200_batch_settings* st, b200_float* d_x, b200_float* d_y, int* d_iters, int* d_status,
                     int* d_cg_iters, int* d_rho_updates, b200_float* d_obj, b200_float* d_prim_res,
                     b200_float* d_dual_res) {
  if (nb <= 0 || n <= 0) return 0;
  Context& cx = ctx();
  BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.m = m; a.nb = nb;
  a.Prp = P->d_row_ptr; a.Pci = P->d_col_ind; a.Pv = P->d_val;
  if (m > 0) {
    a.Arp = A->d_row_ptr; a.Aci = A->d_col_ind; a.Av = A->d_val;
    a.Atrp = At->d_row_ptr; a.Atci = At->d_col_ind; a.Atv = At->d_val;
  }
  a.q = d_q; a.q_batch = d_q_batch; a.l_batch = d_l_batch; a.u_batch = d_u_batch;
  a.D = d_D; a.Dinv = d_Dinv; a.E = d_E; a.Einv = d_Einv; a.c = c; a.cinv = cinv;
  a.rho0 = st->rho; a.sigma = st->sigma; a.alpha = st->alpha; a.eps_abs = st->eps_abs; a.eps_rel = st->eps_rel;
  a.adaptive_rho_tolerance = st->adaptive_rho_tolerance;
  a.rho_is_vec = st->rho_is_vec; a.max_iter = st->max_iter; a.check_termination = st->check_termination;
  a.adaptive_rho = st->adaptive_rho; a.adaptive_rho_interval = st->adaptive_rho_interval;
  a.check_dualgap = st->check_dualgap; a.scaled_termination = st->scaled_termination;
  a.cg_max_iter = st->cg_max_iter; a.cg_tol_reduction = st->cg_tol_reduction; a.cg_tol_fraction = st->cg_tol_fraction;
  a.x_out = d_x; a.y_out = d_y; a.iters = d_iters; a.status = d_status; a.cg_iters = d_cg_iters;
  a.rho_updates = d_rho_updates; a.obj = d_obj; a.prim_res = d_prim_res; a.dual_res = d_dual_res;
  if (m == 0) return 2;                  // unconstrained batches: use the ordinary API (closed-form after one solve)
  const size_t smem = sizeof(T) * ((size_t)8 * n + 8 * m + 1) + sizeof(double) * (17 * kBW + 8);
  if (smem > 227u * 1024u) return 3;     // iterates do not fit one SM's shared memory: not a "small QP"
  static thread_local size_t configured = 0;
  if (smem > configured) {
    if (!B200_CHECK(cudaFuncSetAttribute(batch_admm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return 1;
    // NOT the maximum shared-memory carve-out: 5-6 CTAs per SM then fit, but L1 shrinks below the 30 KB of
    // matrix entries every CTA re-reads through the read-only path -- measured 46 ms (4 CTAs), 45 ms (5),
    // 43 ms (6) against 31 ms with the default carve-out and 4 CTAs per SM (gpurun_out/r2c15_batch_variants.log)
    configured = smem;
  }
  int per_sm = 0;
  if (!B200_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, batch_admm_kernel, kBB, smem))) return 1;
  if (per_sm < 1) per_sm = 1;
  int grid = per_sm * cx.sm_count;
  if (grid > nb) grid = nb;
  batch_admm_kernel<<<grid, kBB, smem, cx.stream>>>(a);
  count_launch("batch_admm_kernel");
  return 0;
}

}  // extern "C"
