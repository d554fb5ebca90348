// Micro-benchmark harness for SpMV kernel design on a Lasso-like matrix (development aid).
// Builds A = [Ad -I 0; I 0 -I; I 0 I] with Ad = sprand(1e6, 1e5, 1e-4) in CSR and times variants.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#include <cuda_runtime.h>

typedef double T;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ int lds_i(const int* p) { int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ double lds_d(const double* p) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

struct Desc { int row0, nrows, nnz0, cnt, lg; };

// ---------------- variant A: CTA tile, stage products in smem, group reduce (the shipped design)
template <int BLOCK, int TILE, int KU, int MODE>   // MODE 0 full, 1 no gather, 2 no reduce
__global__ void __launch_bounds__(BLOCK) k_tile(const int* __restrict__ rp, const int* __restrict__ ci,
                                                const T* __restrict__ va, const Desc* __restrict__ desc, int nblocks,
                                                const T* __restrict__ x, T* __restrict__ y) {
  __shared__ T sm[TILE];
  __shared__ int srp[TILE / 2 + 1];
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const Desc d = desc[b];
    for (int i = tid; i <= d.nrows; i += BLOCK) srp[i] = lds_i(rp + d.row0 + i) - d.nnz0;
    T keep = 0;
    for (int k0 = 0; k0 < d.cnt; k0 += KU * BLOCK) {
      int c[KU]; T v[KU];
#pragma unroll
      for (int u = 0; u < KU; u++) { int k = k0 + u * BLOCK + tid; if (k < d.cnt) { c[u] = lds_i(ci + d.nnz0 + k); v[u] = lds_d(va + d.nnz0 + k); } }
#pragma unroll
      for (int u = 0; u < KU; u++) {
        int k = k0 + u * BLOCK + tid;
        if (k < d.cnt) {
          T xv = (MODE == 1) ? (T)(c[u] & 7) : __ldg(x + c[u]);
          if (MODE == 2) keep += v[u] * xv; else sm[k] = v[u] * xv;
        }
      }
    }
    if (MODE == 2) { if (keep == 1234.5) y[d.row0] = keep; continue; }
    __syncthreads();
    const int g = 1 << d.lg, gid = tid >> d.lg, lig = tid & (g - 1), ng = BLOCK >> d.lg;
    for (int base = 0; base < d.nrows; base += ng) {
      int r = base + gid; T acc = 0;
      if (r < d.nrows) { int e = srp[r + 1]; for (int k = srp[r] + lig; k < e; k += g) acc += sm[k]; }
      for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (r < d.nrows && lig == 0) y[d.row0 + r] = acc;
    }
    __syncthreads();
  }
}

// ---------------- variant B: CSR-vector, G lanes per row straight from global memory
template <int G>
__global__ void __launch_bounds__(256) k_vector(const int* __restrict__ rp, const int* __restrict__ ci,
                                                const T* __restrict__ va, int nrows, const T* __restrict__ x, T* __restrict__ y) {
  const int lane = threadIdx.x & (G - 1);
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const long long ngr = (long long)gridDim.x * blockDim.x / G;
  for (long long r = gid; r < nrows; r += ngr) {
    int s = __ldg(rp + r), e = __ldg(rp + r + 1);
    T acc = 0;
    for (int k = s + lane; k < e; k += G) acc += lds_d(va + k) * __ldg(x + lds_i(ci + k));
    for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[r] = acc;
  }
}

// ---------------- variant C: warp tile (WT entries per warp), products in registers, smem per warp, no CTA sync
template <int WT>
__global__ void __launch_bounds__(256) k_warptile(const int* __restrict__ rp, const int* __restrict__ ci,
                                                  const T* __restrict__ va, const Desc* __restrict__ desc, int nblocks,
                                                  const T* __restrict__ x, T* __restrict__ y) {
  __shared__ T sm[8][WT];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gw = blockIdx.x * 8 + w, GW = gridDim.x * 8;
  T* s = sm[w];
  for (int b = gw; b < nblocks; b += GW) {
    const Desc d = desc[b];
    int c[WT / 32]; T v[WT / 32];
#pragma unroll
    for (int u = 0; u < WT / 32; u++) { int k = lane + 32 * u; if (k < d.cnt) { c[u] = lds_i(ci + d.nnz0 + k); v[u] = lds_d(va + d.nnz0 + k); } }
#pragma unroll
    for (int u = 0; u < WT / 32; u++) { int k = lane + 32 * u; if (k < d.cnt) s[k] = v[u] * __ldg(x + c[u]); }
    __syncwarp();
    for (int base = 0; base < d.nrows; base += 32) {
      int r = base + lane;
      if (r < d.nrows) {
        int st = __ldg(rp + d.row0 + r) - d.nnz0, en = __ldg(rp + d.row0 + r + 1) - d.nnz0;
        T acc = 0;
        for (int k = st; k < en; k++) acc += s[k];
        y[d.row0 + r] = acc;
      }
    }
    __syncwarp();
  }
}


// ---------------- variant D: persistent CTA-tile with the NEXT tile's (col, val, row_ptr) prefetched into
// registers while the current tile is gathered and reduced (hides the HBM latency of the stream)
template <int BLOCK, int KU, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_tile_pf(const int* __restrict__ rp, const int* __restrict__ ci,
                                                         const T* __restrict__ va, const Desc* __restrict__ desc, int nblocks,
                                                         const T* __restrict__ x, T* __restrict__ y) {
  constexpr int TILE = BLOCK * KU;
  constexpr int NRP = (TILE / 2 + BLOCK) / BLOCK;      // row pointers per thread (<= TILE/2 + 1 rows)
  __shared__ T sm[TILE];
  __shared__ int srp[TILE / 2 + 1];
  const int tid = threadIdx.x;
  int b = blockIdx.x;
  if (b >= nblocks) return;
  Desc d = desc[b];
  int c[KU]; T v[KU]; int rpr[NRP];
#pragma unroll
  for (int u = 0; u < KU; u++) { int k = u * BLOCK + tid; if (k < d.cnt) { c[u] = lds_i(ci + d.nnz0 + k); v[u] = lds_d(va + d.nnz0 + k); } }
#pragma unroll
  for (int u = 0; u < NRP; u++) { int i = u * BLOCK + tid; if (i <= d.nrows) rpr[u] = lds_i(rp + d.row0 + i) - d.nnz0; }
  for (;;) {
    // gathers of the current tile
    T g[KU];
#pragma unroll
    for (int u = 0; u < KU; u++) { int k = u * BLOCK + tid; g[u] = (k < d.cnt) ? __ldg(x + c[u]) : (T)0; }
#pragma unroll
    for (int u = 0; u < NRP; u++) { int i = u * BLOCK + tid; if (i <= d.nrows) srp[i] = rpr[u]; }
    // prefetch of the next tile (registers)
    const int bn = b + gridDim.x;
    Desc dn = d;
    int cn[KU]; T vn[KU];
    if (bn < nblocks) {
      dn = desc[bn];
#pragma unroll
      for (int u = 0; u < KU; u++) { int k = u * BLOCK + tid; if (k < dn.cnt) { cn[u] = lds_i(ci + dn.nnz0 + k); vn[u] = lds_d(va + dn.nnz0 + k); } }
#pragma unroll
      for (int u = 0; u < NRP; u++) { int i = u * BLOCK + tid; if (i <= dn.nrows) rpr[u] = lds_i(rp + dn.row0 + i) - dn.nnz0; }
    }
#pragma unroll
    for (int u = 0; u < KU; u++) { int k = u * BLOCK + tid; if (k < d.cnt) sm[k] = v[u] * g[u]; }
    __syncthreads();
    const int gg = 1 << d.lg, gid = tid >> d.lg, lig = tid & (gg - 1), ng = BLOCK >> d.lg;
    for (int base = 0; base < d.nrows; base += ng) {
      int r = base + gid; T acc = 0;
      if (r < d.nrows) { int e = srp[r + 1]; for (int k = srp[r] + lig; k < e; k += gg) acc += sm[k]; }
      for (int o = gg >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (r < d.nrows && lig == 0) y[d.row0 + r] = acc;
    }
    __syncthreads();
    if (bn >= nblocks) break;
    b = bn; d = dn;
#pragma unroll
    for (int u = 0; u < KU; u++) { c[u] = cn[u]; v[u] = vn[u]; }
  }
}


// ---------------- variant E: warp-autonomous tiles (no CTA barrier): each warp owns WT-entry tiles of whole
// rows, stages products in its private shared-memory segment, __syncwarp, group-reduces its rows
template <int WT, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_warp(const int* __restrict__ rp, const int* __restrict__ ci,
                                                          const T* __restrict__ va, const Desc* __restrict__ desc, int nblocks,
                                                          const T* __restrict__ x, T* __restrict__ y) {
  constexpr int KU = WT / 32;
  __shared__ T sm_all[WARPS][WT];
  __shared__ int srp_all[WARPS][WT / 2 + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T* sm = sm_all[w]; int* srp = srp_all[w];
  for (int b = blockIdx.x * WARPS + w; b < nblocks; b += gridDim.x * WARPS) {
    const Desc d = desc[b];
    for (int i = lane; i <= d.nrows; i += 32) srp[i] = lds_i(rp + d.row0 + i) - d.nnz0;
    int c[KU]; T v[KU];
#pragma unroll
    for (int u = 0; u < KU; u++) { int k = u * 32 + lane; if (k < d.cnt) { c[u] = lds_i(ci + d.nnz0 + k); v[u] = lds_d(va + d.nnz0 + k); } }
#pragma unroll
    for (int u = 0; u < KU; u++) { int k = u * 32 + lane; if (k < d.cnt) sm[k] = v[u] * __ldg(x + c[u]); }
    __syncwarp();
    const int g = 1 << d.lg, gid = lane >> d.lg, lig = lane & (g - 1), ng = 32 >> d.lg;
    for (int base = 0; base < d.nrows; base += ng) {
      int r = base + gid; T acc = 0;
      if (r < d.nrows) { int e = srp[r + 1]; for (int k = srp[r] + lig; k < e; k += g) acc += sm[k]; }
      for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (r < d.nrows && lig == 0) y[d.row0 + r] = acc;
    }
    __syncwarp();
  }
}

template <class F> float timeit(F f, int reps = 20) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) f();
  cudaEventRecord(e0); for (int i = 0; i < reps; i++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); CK(cudaGetLastError()); return ms / reps * 1e3f;
}

std::vector<Desc> schedule(const std::vector<int>& rp, int tile, int maxrows, int block) {
  std::vector<Desc> d; int nrows = (int)rp.size() - 1, r = 0;
  while (r < nrows) {
    int r1 = r, cnt = 0;
    while (r1 < nrows && r1 - r < maxrows) { int l = rp[r1 + 1] - rp[r1]; if (cnt + l > tile) break; cnt += l; r1++; }
    if (r1 == r) { printf("row too long\n"); exit(1); }
    int nr = r1 - r, lg = 0, mean = (cnt + nr - 1) / nr;
    while (lg < 5 && (2 << lg) * nr <= block && (1 << lg) < mean) lg++;
    d.push_back({r, nr, rp[r], cnt, lg}); r = r1;
  }
  return d;
}

int main(int argc, char** argv) {
  const int nf = 100000, ms_ = 1000000;
  std::mt19937 rng(1);
  std::vector<int> rp{0}, ci; std::vector<T> va;
  std::poisson_distribution<int> pois(10.0);
  std::uniform_int_distribution<int> col(0, nf - 1);
  std::uniform_real_distribution<double> uni(0, 1);
  for (int i = 0; i < ms_; i++) {
    int k = pois(rng); std::vector<int> cs(k); for (auto& c : cs) c = col(rng);
    std::sort(cs.begin(), cs.end()); cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
    for (int c : cs) { ci.push_back(c); va.push_back(uni(rng)); }
    ci.push_back(nf + i); va.push_back(-1.0); rp.push_back((int)ci.size());
  }
  for (int s = 0; s < 2; s++) for (int j = 0; j < nf; j++) { ci.push_back(j); va.push_back(1.0); ci.push_back(nf + ms_ + j); va.push_back(s ? 1.0 : -1.0); rp.push_back((int)ci.size()); }
  const int m = (int)rp.size() - 1, n = nf + ms_ + nf, nnz = (int)ci.size();
  const double bytes = nnz * 12.0 + (m + 1) * 4.0 + n * 8.0 + m * 8.0;
  printf("A: %d x %d nnz %d  algorithmic %.1f MB\n", m, n, nnz, bytes / 1e6);
  std::vector<T> hx(n); for (auto& v : hx) v = uni(rng);
  std::vector<double> yref(m, 0.0);
  for (int r = 0; r < m; r++) { double a = 0; for (int k = rp[r]; k < rp[r + 1]; k++) a += va[k] * hx[ci[k]]; yref[r] = a; }
  int *drp, *dci; T *dva, *dx, *dy;
  CK(cudaMalloc(&drp, (m + 16) * 4)); CK(cudaMalloc(&dci, (nnz + 16) * 4)); CK(cudaMalloc(&dva, (nnz + 16) * 8));
  CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&dy, m * 8));
  CK(cudaMemcpy(drp, rp.data(), (m + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dci, ci.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dva, va.data(), nnz * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dx, hx.data(), n * 8, cudaMemcpyHostToDevice));
  auto check = [&](const char* name, float us) {
    std::vector<T> hy(m); CK(cudaMemcpy(hy.data(), dy, m * 8, cudaMemcpyDeviceToHost));
    double err = 0; for (int r = 0; r < m; r++) err = std::max(err, std::abs(hy[r] - yref[r]));
    printf("%-44s %7.1f us  %6.0f GB/s  (%.0f%% of 6541)  maxerr %.1e\n", name, us, bytes / us / 1e3, bytes / us / 1e3 / 65.41, err);
    CK(cudaMemset(dy, 0, m * 8));
  };
  auto upload = [&](const std::vector<Desc>& d) { Desc* p; CK(cudaMalloc(&p, d.size() * sizeof(Desc))); CK(cudaMemcpy(p, d.data(), d.size() * sizeof(Desc), cudaMemcpyHostToDevice)); return p; };
  // sweep: single-batch tiles (TILE = BLOCK * KU), persistent grids of 148 * c CTAs vs one CTA per tile
#define SWEEP(BLOCK, TILE, KU)                                                                          \
  {                                                                                                     \
    auto d = schedule(rp, TILE, TILE / 2, BLOCK); Desc* dd = upload(d); int nb = (int)d.size();       \
    for (int c : {0, 1024 / BLOCK, 2048 / BLOCK}) {                                                     \
      int g = c ? std::min(nb, 148 * c) : nb; char nm[96];                                              \
      snprintf(nm, 96, "tile%d b%d ku%d grid=%s(%d)", TILE, BLOCK, KU, c ? "148x" : "nb", c ? c : nb); \
      check(nm, timeit([&] { k_tile<BLOCK, TILE, KU, 0><<<g, BLOCK>>>(drp, dci, dva, dd, nb, dx, dy); })); \
    }                                                                                                   \
  }
  SWEEP(256, 1024, 4) SWEEP(512, 2048, 4)
#define PF(BLOCK, KU, MINB)                                                                             \
  {                                                                                                     \
    auto d = schedule(rp, BLOCK * KU, BLOCK * KU / 2, BLOCK); Desc* dd = upload(d); int nb = (int)d.size(); \
    int g = std::min(nb, 148 * MINB); char nm[96];                                                      \
    snprintf(nm, 96, "prefetch tile%d b%d ku%d grid=148x%d", BLOCK * KU, BLOCK, KU, MINB);             \
    check(nm, timeit([&] { k_tile_pf<BLOCK, KU, MINB><<<g, BLOCK>>>(drp, dci, dva, dd, nb, dx, dy); })); \
  }
  PF(512, 4, 3)
#define WV(WT, WARPS, MINB, PERCTA)                                                                      \
  {                                                                                                     \
    auto d = schedule(rp, WT, WT / 2, 32); Desc* dd = upload(d); int nb = (int)d.size();                \
    int g = PERCTA ? (nb + WARPS * PERCTA - 1) / (WARPS * PERCTA) : 148 * MINB; char nm[96];            \
    snprintf(nm, 96, "warp tile%d warps%d minb%d grid=%d", WT, WARPS, MINB, g);                        \
    check(nm, timeit([&] { k_warp<WT, WARPS, MINB><<<g, WARPS * 32>>>(drp, dci, dva, dd, nb, dx, dy); })); \
  }
  WV(128, 8, 8, 0) WV(128, 8, 8, 1) WV(128, 8, 8, 4) WV(256, 8, 6, 0) WV(256, 8, 6, 1) WV(256, 8, 6, 4)
  WV(128, 4, 16, 1) WV(128, 4, 16, 4) WV(64, 8, 8, 1) WV(64, 8, 8, 8) WV(256, 4, 12, 2) WV(512, 8, 3, 1)
  return 0;
}
