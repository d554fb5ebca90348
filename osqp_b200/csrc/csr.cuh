// csr.cuh -- device CSR view, row-block schedule and the "CSR-stream" pass shared by the
// stand-alone SpMV kernels (csr.cu) and the persistent PCG kernel (pcg.cu).
//
// Schedule (built once on the host in b200_csr_create):
//   * a normal block covers whole consecutive rows with <= kTile nonzeros and <= kMaxRows rows;
//     one CTA streams val/col of the block with fully coalesced, L1-bypassing loads (kU
//     independent (col, val, gather) chains in flight per thread), stages the per-entry terms in
//     shared memory, then sub-warp groups of g lanes (g = 1..32 from the descriptor, chosen so
//     that every row of the block gets its own group in one trip) reduce one row each.
//   * a row with more than kTile nonzeros is cut into kTile-sized chunks, one block each; each
//     chunk publishes a partial and the LAST CTA to arrive (atomic ticket on an integer counter)
//     folds the partials in chunk order -> deterministic, no floating-point atomics.
//
// Measured design notes (Lasso A, 1.14e7 nnz, f64; profiles/ and DESIGN.md have the numbers):
//   * B200 retires ONE divergent 8-byte gather per clock per SM (tools/micro/gather_bench.cu:
//     1.14e7 random gathers take 41-45 us whatever the unroll / CTA size / precision), so a
//     random-sparse SpMV is bound by its gathers (~36 us here), not by the 21 us HBM stream.
//   * three TMA-fed variants (cp.async.bulk + mbarrier rings: CTA ring of 3 x 28 KB stages,
//     per-warp rings of 128-entry tiles, CTA ring + one dense gather burst) all measured SLOWER
//     (76 / 135 / 125 us vs 65 us): hiding the latency of the HBM stream does not help when the
//     L1 gather queue is the scarce resource, and flooding it stalls the other CTAs' shared-
//     memory traffic.  They are kept in the git history of this file.
//
// Algorithmic bytes of one pass over an r x c matrix with nnz entries:
//   nnz (sizeof(T)+4) + (r+1) 4 + c sizeof(T) [gather, once] + r sizeof(T) [store]
// (SURVEY.md section 8d).
#pragma once

#include "common.cuh"

namespace b200 {

#ifndef B200_SPMV_BLOCK
#define B200_SPMV_BLOCK 256
#endif
#ifndef B200_SPMV_TILE
#define B200_SPMV_TILE 2048
#endif
#ifndef B200_SPMV_MINBLOCKS
#define B200_SPMV_MINBLOCKS 4   /* caps the pass kernels at 64 registers (ptxas takes 94 unbounded) */
#endif
#ifndef B200_PCG_MINBLOCKS
#define B200_PCG_MINBLOCKS 4
#endif
constexpr int kSpmvBlock = B200_SPMV_BLOCK;  // threads per CTA of every kernel that runs spmv_pass
constexpr int kTile      = B200_SPMV_TILE;   // staged nonzeros per CTA pass
constexpr int kMaxRows   = kTile / 2;        // rows per normal block
constexpr int kLongRow   = kTile / 2;        // rows with more entries get CTA-wide chunk(s) of their own
constexpr int kPad       = 8;      // slack elements at the end of every matrix array

constexpr int kSmemElems     = kTile + 40;                  // staged terms + reduction scratch
constexpr int kSpmvSmemBytes = kSmemElems * (int)sizeof(T) + (kMaxRows + 8) * (int)sizeof(int) + 16;

struct CsrView {
  const int*  row_ptr;
  const int*  col_ind;
  T*          val;
  const int4* desc;          // per block: {row0, nrows | log2g<<24  or  -(lr+1), nnz0, cnt}
  const int4* long_rows;     // per long row: {row, first_block, nchunks, 0}
  double*     long_partials; // indexed by block id
  unsigned*   long_counters; // indexed by long-row id
  int nrows, ncols, nnz, nblocks, nlong;
};

struct SumOp {
  __device__ __forceinline__ static T identity() { return (T)0; }
  __device__ __forceinline__ static T apply(T a, T b) { return a + b; }
};
struct MaxOp {
  __device__ __forceinline__ static T identity() { return (T)0; }   // used on |.| only
  __device__ __forceinline__ static T apply(T a, T b) { return (b > a) ? b : a; }
};

template <class CB>
__device__ __forceinline__ T group_reduce(T v, int g) {
  // g is uniform over the CTA; lanes of one group are contiguous inside a warp
  for (int o = g >> 1; o > 0; o >>= 1) v = CB::apply(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <class CB>
__device__ __forceinline__ T block_reduce_T(T v, T* sh /* >= 33 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = group_reduce<CB>(v, 32);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    T t = (lane < (blockDim.x >> 5)) ? sh[lane] : CB::identity();
    t = group_reduce<CB>(t, 32);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// streaming (read-once) loads of the matrix arrays: non-coherent path, do not allocate in L1 so
// that L1 stays available for the gathered vector entries
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// CTA shared memory of a pass: kTile staged terms, 40 scratch elements, the block's row
// pointers (rebased to the block), a flag.
struct Pipe {
  T*   sm;
  int* srp;
  int* flag;
};

__device__ __forceinline__ Pipe pipe_init(unsigned char* dsm) {
  Pipe P;
  P.sm   = reinterpret_cast<T*>(dsm);
  P.srp  = reinterpret_cast<int*>(dsm + kSmemElems * sizeof(T));
  P.flag = P.srp + kMaxRows + 4;
  return P;
}

// One pass over the blocks cta, cta+G, ... of M.
//   ef(k, col, val) -> T   term contributed by stored entry k (k = global entry index)
//   CB                     combine op over the terms of one row (SumOp / MaxOp)
//   ep(row, value)         called exactly once per row, by one thread
// M.val / M.col_ind / M.row_ptr must not be written while the calling kernel runs (they are read
// through the non-coherent path).  All threads of the CTA must call this together.
template <class CB, class EF, class EP>
__device__ __forceinline__ void spmv_pass(const CsrView& M, int cta, int G, Pipe& P, EF ef, EP ep) {
  const int tid = threadIdx.x;
  // independent (col, val, gather) chains in flight per thread.  Measured on the Lasso matrix
  // (tools/micro/spmv_variants.cu, profiles/r01_spmv_ncu.md): lean stand-alone kernels prefer one
  // batch per tile (56 us), the register-bound persistent PCG kernel prefers 2048-entry tiles in
  // two batches of 4 (204 us / CG iteration vs 212-243 for the other tilings).
  constexpr int kU = 4;
  T* const sm = P.sm;
  for (int b = cta; b < M.nblocks; b += G) {
    const int4 d   = __ldg(M.desc + b);
    const int nnz0 = d.z, cnt = d.w;
    if (d.y >= 0) {
      // ---- normal block: stage terms, then per-row group reduction
      const int nrows = d.y & 0xffffff;
      const int lg    = d.y >> 24;
      const int g     = 1 << lg;
      int* const srp  = P.srp;
      for (int i = tid; i <= nrows; i += kSpmvBlock) srp[i] = ld_stream(M.row_ptr + d.x + i) - nnz0;
      for (int k0 = 0; k0 < cnt; k0 += kU * kSpmvBlock) {
        int c[kU];
        T   v[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
          const int k = k0 + u * kSpmvBlock + tid;
          if (k < cnt) {
            c[u] = ld_stream(M.col_ind + nnz0 + k);
            v[u] = ld_stream(M.val + nnz0 + k);
          }
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
          const int k = k0 + u * kSpmvBlock + tid;
          if (k < cnt) sm[k] = ef(nnz0 + k, c[u], v[u]);
        }
      }
      __syncthreads();
      const int gid    = tid >> lg;
      const int lig    = tid & (g - 1);
      const int ngroup = kSpmvBlock >> lg;
      for (int base = 0; base < nrows; base += ngroup) {
        const int r = base + gid;
        T acc = CB::identity();
        if (r < nrows) {
          const int e = srp[r + 1];
          for (int k = srp[r] + lig; k < e; k += g) acc = CB::apply(acc, sm[k]);
        }
        acc = group_reduce<CB>(acc, g);
        if (r < nrows && lig == 0) ep(d.x + r, acc);
      }
      __syncthreads();   // sm is reused by the next block
    } else {
      // ---- chunk of a long row
      const int  lr   = -d.y - 1;
      const int4 info = __ldg(M.long_rows + lr);
      T acc = CB::identity();
      for (int k0 = 0; k0 < cnt; k0 += kU * kSpmvBlock) {
        int c[kU];
        T   v[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
          const int k = k0 + u * kSpmvBlock + tid;
          if (k < cnt) {
            c[u] = ld_stream(M.col_ind + nnz0 + k);
            v[u] = ld_stream(M.val + nnz0 + k);
          }
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
          const int k = k0 + u * kSpmvBlock + tid;
          if (k < cnt) acc = CB::apply(acc, ef(nnz0 + k, c[u], v[u]));
        }
      }
      T* sh = sm + kTile;
      acc = block_reduce_T<CB>(acc, sh);
      if (tid == 0) {
        M.long_partials[b] = (double)acc;
        __threadfence();
        const unsigned t = atomicAdd(&M.long_counters[lr], 1u);
        *P.flag = (t == (unsigned)(info.z - 1));
      }
      __syncthreads();
      if (*P.flag) {
        __threadfence();
        T a = CB::identity();
        for (int c = tid; c < info.z; c += kSpmvBlock)
          a = CB::apply(a, (T)__ldcg(&M.long_partials[info.y + c]));
        a = block_reduce_T<CB>(a, sh);
        if (tid == 0) {
          M.long_counters[lr] = 0;
          ep(info.x, a);
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace b200

// host-side object behind the opaque C handle
struct b200_csr {
  int nrows = 0, ncols = 0, nnz = 0, nblocks = 0, nlong = 0;
  int*      d_row_ptr = nullptr;
  int*      d_col_ind = nullptr;
  T*        d_val     = nullptr;
  int4*     d_desc    = nullptr;
  int4*     d_long    = nullptr;
  double*   d_long_partials = nullptr;
  unsigned* d_long_counters = nullptr;
  b200::CsrView view() const {
    b200::CsrView v;
    v.row_ptr = d_row_ptr; v.col_ind = d_col_ind; v.val = d_val; v.desc = d_desc;
    v.long_rows = d_long; v.long_partials = d_long_partials; v.long_counters = d_long_counters;
    v.nrows = nrows; v.ncols = ncols; v.nnz = nnz; v.nblocks = nblocks; v.nlong = nlong;
    return v;
  }
};

// build the row-block schedule for a host CSR pattern (shared by csr.cu and pcg.cu)
int b200_build_schedule(b200_csr* M, const int* h_row_ptr);

// every kernel using spmv_pass needs kSpmvSmemBytes of dynamic shared memory
template <class K>
inline void b200_enable_spmv_smem(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, b200::kSpmvSmemBytes);
}
