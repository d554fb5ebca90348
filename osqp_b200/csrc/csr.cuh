// csr.cuh -- device CSR view, row-block schedule and the TMA-fed "CSR-stream" pass shared by the
// stand-alone SpMV kernels (csr.cu) and the persistent PCG kernel (pcg.cu).
//
// Schedule (built once on the host in b200_csr_create):
//   * a normal block covers whole consecutive rows with <= kTile nonzeros and <= kMaxRows rows;
//   * a row with more than kTile nonzeros is cut into kTile-sized chunks, one block each; each
//     chunk publishes a partial and the LAST CTA to arrive (atomic ticket on an integer counter)
//     folds the partials in chunk order -> deterministic, no floating-point atomics.
//
// Execution (spmv_pass): a CTA walks its blocks b = cta, cta+G, ... through a kStages-deep
// shared-memory ring.  One elected thread issues, per block, three 1-D TMA bulk copies
// (cp.async.bulk.shared.global + mbarrier complete_tx): the block's column indices, values and
// row pointers, kStages-1 blocks ahead of the arithmetic, so the HBM stream costs no registers,
// no LSU issue slots and no exposed latency.  When a block has landed, ALL threads take its
// entries round-robin and issue their gathers of the vector in one dense burst (the gather is
// the scarce resource: measured 1 divergent 8-byte gather per clock per SM on B200, see
// tools/micro/gather_bench.cu), stage the products in shared memory, and groups of g lanes
// (g = 1..32 from the descriptor) then sum one row each and call the epilogue.  Two CTAs per SM
// alternate so that one gathers while the other reduces.
//
// Algorithmic bytes of one pass over an r x c matrix with nnz entries:
//   nnz (sizeof(T)+4) + (r+1) 4 + c sizeof(T) [gather, once] + r sizeof(T) [store]
// (SURVEY.md section 8d).
#pragma once

#include "common.cuh"

namespace b200 {

constexpr int kSpmvBlock = 512;    // threads per CTA of every kernel that runs spmv_pass
constexpr int kTile      = 2048;   // nonzeros per block / stage
constexpr int kMaxRows   = 1024;   // rows per normal block
constexpr int kStages    = 3;      // depth of the TMA ring
constexpr int kPad       = 8;      // slack elements: aligned-down starts + 16-byte rounding
constexpr int kGU        = kTile / kSpmvBlock;   // gathers in flight per thread

constexpr int kColsBytes  = (kTile + kPad) * 4;
constexpr int kValsBytes  = (kTile + kPad) * (int)sizeof(T);
constexpr int kRpBytes    = (kMaxRows + kPad) * 4;
constexpr int kStageBytes = kColsBytes + kValsBytes + kRpBytes;
constexpr int kProdBytes  = kTile * (int)sizeof(T);
constexpr int kScratchElems = 40;
constexpr int kSpmvSmemBytes = kStages * kStageBytes + kProdBytes +
                               kScratchElems * (int)sizeof(double) + kStages * 8 + 16;
static_assert(kColsBytes % 16 == 0 && kValsBytes % 16 == 0 && kRpBytes % 16 == 0 &&
              kProdBytes % 16 == 0, "TMA alignment");

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

// ------------------------------------------------------------------ TMA / mbarrier PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Shared-memory ring of one CTA.  `it` counts blocks consumed since pipe_init: block number `it`
// lives in stage it % kStages and is the (it / kStages)-th use of that stage's barrier.
struct Pipe {
  unsigned char* base;
  T*             prod;      // kTile staged products
  T*             scratch;   // reduction scratch
  uint64_t*      bars;
  int*           flag;      // last-arriver broadcast
  unsigned       it;
};

__device__ __forceinline__ Pipe pipe_init(unsigned char* dsm) {
  Pipe P;
  P.base    = dsm;
  P.prod    = reinterpret_cast<T*>(dsm + kStages * kStageBytes);
  P.scratch = reinterpret_cast<T*>(dsm + kStages * kStageBytes + kProdBytes);
  P.bars    = reinterpret_cast<uint64_t*>(dsm + kStages * kStageBytes + kProdBytes +
                                          kScratchElems * sizeof(double));
  P.flag    = reinterpret_cast<int*>(P.bars + kStages);
  P.it      = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&P.bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  return P;
}

constexpr int kValAlign = 16 / (int)sizeof(T);   // elements of T per 16 bytes

// elected thread: start the three bulk copies of block b into stage s
__device__ __forceinline__ void issue_block(const CsrView& M, int b, const Pipe& P, int s) {
  const int4 d = __ldg(M.desc + b);
  unsigned char* st = P.base + s * kStageBytes;
  uint64_t* bar = &P.bars[s];
  const int nnz0 = d.z, cnt = d.w;
  uint32_t bc = 0, bv = 0, br = 0;
  const int c0 = nnz0 & ~3, v0 = nnz0 & ~(kValAlign - 1), r0 = d.x & ~3;
  if (cnt > 0) {
    bc = (uint32_t)(((nnz0 - c0 + cnt) * 4 + 15) & ~15);
    bv = (uint32_t)(((nnz0 - v0 + cnt) * (int)sizeof(T) + 15) & ~15);
  }
  if (d.y >= 0) br = (uint32_t)(((d.x - r0 + (d.y & 0xffffff) + 1) * 4 + 15) & ~15);
  mbar_expect_tx(bar, bc + bv + br);
  if (bc) {
    tma_load_1d(st, M.col_ind + c0, bc, bar);
    tma_load_1d(st + kColsBytes, M.val + v0, bv, bar);
  }
  if (br) tma_load_1d(st + kColsBytes + kValsBytes, M.row_ptr + r0, br, bar);
}

// One pass over the blocks cta, cta+G, ... of M.
//   ef(k, col, val) -> T   term contributed by stored entry k (k = global entry index)
//   CB                     combine op over the terms of one row (SumOp / MaxOp)
//   ep(row, value)         called exactly once per row, by one thread
// M.val / M.col_ind / M.row_ptr must not be written while the calling kernel runs (they are read
// through the async proxy).  All threads of the CTA must call this together.
template <class CB, class EF, class EP>
__device__ __forceinline__ void spmv_pass(const CsrView& M, int cta, int G, Pipe& P, EF ef, EP ep) {
  const int tid   = threadIdx.x;
  const int nmine = (cta < M.nblocks) ? (M.nblocks - cta + G - 1) / G : 0;
  if (tid == 0) {
    const int pre = nmine < kStages - 1 ? nmine : kStages - 1;
    for (int i = 0; i < pre; i++) issue_block(M, cta + i * G, P, (P.it + i) % kStages);
  }
  T* const prod = P.prod;
  for (int i = 0; i < nmine; i++) {
    const int b = cta + i * G;
    const int s = P.it % kStages;
    if (tid == 0 && i + kStages - 1 < nmine)
      issue_block(M, b + (kStages - 1) * G, P, (P.it + kStages - 1) % kStages);
    const int4 d = __ldg(M.desc + b);
    const int nnz0 = d.z, cnt = d.w;
    unsigned char* st = P.base + s * kStageBytes;
    const int* cols = reinterpret_cast<const int*>(st) + (nnz0 & 3);
    const T*   vals = reinterpret_cast<const T*>(st + kColsBytes) + (nnz0 & (kValAlign - 1));
    const int* rp   = reinterpret_cast<const int*>(st + kColsBytes + kValsBytes) + (d.x & 3);
    mbar_wait(&P.bars[s], (P.it / kStages) & 1);

    // ---- dense gather burst: kGU independent gathers per thread
    T pr[kGU];
#pragma unroll
    for (int u = 0; u < kGU; u++) {
      const int k = tid + u * kSpmvBlock;
      pr[u] = (k < cnt) ? ef(nnz0 + k, cols[k], vals[k]) : CB::identity();
    }

    if (d.y >= 0) {
      // ---- normal block: stage products, then g lanes per row
#pragma unroll
      for (int u = 0; u < kGU; u++) {
        const int k = tid + u * kSpmvBlock;
        if (k < cnt) prod[k] = pr[u];
      }
      __syncthreads();
      const int nrows  = d.y & 0xffffff;
      const int lg     = d.y >> 24;
      const int g      = 1 << lg;
      const int gid    = tid >> lg;
      const int lig    = tid & (g - 1);
      const int ngroup = kSpmvBlock >> lg;
      for (int base = 0; base < nrows; base += ngroup) {
        const int r = base + gid;
        T acc = CB::identity();
        if (r < nrows) {
          int k = rp[r] - nnz0 + lig;
          const int e = rp[r + 1] - nnz0;
          for (; k + 3 * g < e; k += 4 * g) {
            const T t0 = prod[k], t1 = prod[k + g], t2 = prod[k + 2 * g], t3 = prod[k + 3 * g];
            acc = CB::apply(CB::apply(acc, t0), t1);
            acc = CB::apply(CB::apply(acc, t2), t3);
          }
          for (; k < e; k += g) acc = CB::apply(acc, prod[k]);
        }
        acc = group_reduce<CB>(acc, g);
        if (r < nrows && lig == 0) ep(d.x + r, acc);
      }
    } else {
      // ---- chunk of a long row
      const int  lr   = -d.y - 1;
      const int4 info = __ldg(M.long_rows + lr);
      T acc = pr[0];
#pragma unroll
      for (int u = 1; u < kGU; u++) acc = CB::apply(acc, pr[u]);
      acc = block_reduce_T<CB>(acc, P.scratch);
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
        a = block_reduce_T<CB>(a, P.scratch);
        if (tid == 0) {
          M.long_counters[lr] = 0;
          ep(info.x, a);
        }
      }
    }
    __syncthreads();   // everyone is done with stage s and prod before they are refilled
    P.it++;
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
