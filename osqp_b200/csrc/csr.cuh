// csr.cuh -- device CSR view, row-block schedule and the "CSR-stream" row-block primitive
// shared by the stand-alone SpMV kernel (csr.cu) and the persistent PCG kernel (pcg.cu).
//
// Schedule (built once on the host in b200_csr_create):
//   * a normal block covers whole consecutive rows with <= kTile nonzeros and <= kMaxRows
//     rows; one CTA streams val/col of the block with fully coalesced loads, stages the
//     per-entry terms in shared memory, then sub-warp groups of g lanes (g = 1..32, chosen
//     from the mean row length of the block) reduce one row each with shuffles.
//   * a row with more than kTile nonzeros is cut into kTile-sized chunks, one CTA each; each
//     chunk CTA publishes a partial, and the LAST CTA to arrive (atomic ticket on an integer
//     counter) folds the partials in chunk order -> deterministic, no floating-point atomics.
// Algorithmic bytes of one pass over an r x c matrix with nnz entries:
//   nnz (sizeof(T)+4) + (r+1) 4 + c sizeof(T) [gather, once] + r sizeof(T) [store]
// (SURVEY.md section 8d).
#pragma once

#include "common.cuh"

namespace b200 {

constexpr int kTile    = 2048;   // staged nonzeros per CTA pass (16 KB of doubles)
constexpr int kMaxRows = 1024;   // rows per normal block

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

// Process row block `b` of matrix M.
//   ef(k, col, val) -> T   term contributed by stored entry k
//   CB                     combine op over the terms of one row (SumOp / MaxOp)
//   ep(row, value)         called exactly once per row of the block, by one thread
// `sm` is CTA shared memory with at least kTile + 40 elements of T.
template <class CB, class EF, class EP>
__device__ __forceinline__ void rowblock_apply(const CsrView& M, int b, T* sm, EF ef, EP ep) {
  const int4 d   = M.desc[b];
  const int nnz0 = d.z, cnt = d.w;
  const int tid  = threadIdx.x;

  if (d.y >= 0) {
    // ---- normal block: stage terms, then per-row group reduction
    const int nrows = d.y & 0xffffff;
    const int lg    = d.y >> 24;
    const int g     = 1 << lg;
#pragma unroll 4
    for (int k = tid; k < cnt; k += kBlock) {
      const int kk = nnz0 + k;
      sm[k] = ef(kk, __ldg(M.col_ind + kk), M.val[kk]);
    }
    __syncthreads();
    const int gid    = tid >> lg;
    const int lig    = tid & (g - 1);
    const int ngroup = kBlock >> lg;
    for (int base = 0; base < nrows; base += ngroup) {
      const int r = base + gid;
      T acc = CB::identity();
      if (r < nrows) {
        const int s = __ldg(M.row_ptr + d.x + r) - nnz0;
        const int e = __ldg(M.row_ptr + d.x + r + 1) - nnz0;
        for (int k = s + lig; k < e; k += g) acc = CB::apply(acc, sm[k]);
      }
      acc = group_reduce<CB>(acc, g);
      if (r < nrows && lig == 0) ep(d.x + r, acc);
    }
    __syncthreads();   // sm is reused by the next block
  } else {
    // ---- chunk of a long row
    const int  lr   = -d.y - 1;
    const int4 info = M.long_rows[lr];
    T acc = CB::identity();
#pragma unroll 4
    for (int k = tid; k < cnt; k += kBlock) {
      const int kk = nnz0 + k;
      acc = CB::apply(acc, ef(kk, __ldg(M.col_ind + kk), M.val[kk]));
    }
    T* sh = sm + kTile;
    acc = block_reduce_T<CB>(acc, sh);
    __shared__ int s_last;
    if (tid == 0) {
      M.long_partials[b] = (double)acc;
      __threadfence();
      unsigned t = atomicAdd(&M.long_counters[lr], 1u);
      s_last = (t == (unsigned)(info.z - 1));
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      T a = CB::identity();
      for (int c = tid; c < info.z; c += kBlock)
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
