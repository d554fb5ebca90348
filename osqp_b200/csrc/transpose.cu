// transpose.cu -- CSR of the transpose, built on the device.
//
// The user's CSC arrays of A are the CSR of A' (uploaded once); the solver also needs CSR(A) for
// A x.  Building it on the host (a parallel stable counting sort, algebra/b200/matrix.c) and
// uploading a second 12-byte-per-entry copy cost 45 + 14 ms of the 97 ms osqp_setup on the
// 1.14e7-nnz Lasso; here it is five kernels over data that is already in HBM:
//   count     row lengths of the transpose (integer atomics: order-independent result)
//   scan      three-kernel inclusive prefix sum -> row_ptr, and the longest row
//   scatter   every entry takes the next free slot of its destination row (atomic cursor)
//   sort      one warp per destination row puts its entries in ascending column order by RANK
//             (rank = number of entries of the row with a smaller (column, source position)), so
//             the result does not depend on the order in which the atomics were served:
//             bit-identical to the host sort, run-to-run deterministic
//   gather    values through the recorded source positions; map[k] = final position of entry k
//             (device-resident index map for OSQPMatrix_update_values)
// Role in the reference: csr_transpose / csr_from_csc of algebra/cuda/src/cuda_csr.cu:489-628,
// which sorts with thrust and calls cusparseCsr2cscEx2.
#include "csr.cuh"

#include <vector>

using namespace b200;

namespace {

constexpr int kSortMaxRow  = 4096;   // longer destination rows: the caller keeps the host path
constexpr int kScanBlock   = 1024;
constexpr int kScanPerThr  = 4;
constexpr int kScanChunk   = kScanBlock * kScanPerThr;

__global__ void __launch_bounds__(kBlock) tr_count(int nnz, const int* __restrict__ col, int* cnt) {
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) atomicAdd(&cnt[col[k] + 1], 1);
}

// CTA-wide inclusive scan of one value per thread; returns the inclusive prefix, total in `tot`
__device__ __forceinline__ int block_scan_incl(int v, int* sh /* >= 33 */, int& tot) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();
  if (lane == 31) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    int s = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    sh[lane] = s;
  }
  __syncthreads();
  tot = sh[(blockDim.x >> 5) - 1];
  return v + (w > 0 ? sh[w - 1] : 0);
}

// phase 1: totals of the kScanChunk-sized chunks (+ the largest single count)
__global__ void __launch_bounds__(kScanBlock) scan_totals(int n, const int* __restrict__ in, int* totals, int* maxval) {
  __shared__ int sh[33];
  const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanPerThr;
  int s = 0, mx = 0;
#pragma unroll
  for (int u = 0; u < kScanPerThr; u++)
    if (base + u < n) { const int v = in[base + u]; s += v; mx = v > mx ? v : mx; }
  int tot;
  block_scan_incl(s, sh, tot);
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(maxval, mx);
  if (threadIdx.x == 0) totals[blockIdx.x] = tot;
}
// phase 2: exclusive scan of the chunk totals by one CTA
__global__ void __launch_bounds__(kScanBlock) scan_offsets(int nb, int* totals) {
  __shared__ int sh[33];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += kScanBlock) {
    const int i = b0 + threadIdx.x;
    const int v = i < nb ? totals[i] : 0;
    int tot;
    const int inc = block_scan_incl(v, sh, tot);
    if (i < nb) totals[i] = carry + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
}
// phase 3: inclusive scan inside every chunk + its offset, in place
__global__ void __launch_bounds__(kScanBlock) scan_apply(int n, int* data, const int* __restrict__ offsets) {
  __shared__ int sh[33];
  const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanPerThr;
  int v[kScanPerThr], s = 0;
#pragma unroll
  for (int u = 0; u < kScanPerThr; u++) { v[u] = (base + u < n) ? data[base + u] : 0; s += v[u]; }
  int tot;
  int run = block_scan_incl(s, sh, tot) - s + offsets[blockIdx.x];
#pragma unroll
  for (int u = 0; u < kScanPerThr; u++) {
    run += v[u];
    if (base + u < n) data[base + u] = run;
  }
}

// one warp per SOURCE row j: entry k of column i goes to the next free slot of destination row i
__global__ void __launch_bounds__(kBlock) tr_scatter(int nrows_src, const int* __restrict__ rp_src,
                                                     const int* __restrict__ col_src, int* cursor,
                                                     int* tmp_col, int* tmp_src) {
  const int lane = threadIdx.x & 31, wpb = kBlock >> 5;
  for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < nrows_src; j += gridDim.x * wpb) {
    const int s = rp_src[j], e = rp_src[j + 1];
    for (int k = s + lane; k < e; k += 32) {
      const int pos = atomicAdd(&cursor[col_src[k]], 1);
      tmp_col[pos] = j;
      tmp_src[pos] = k;
    }
  }
}

// one warp per DESTINATION row: rank-sort by (column, source position); gather the values
__global__ void __launch_bounds__(kBlock) tr_sort_rows(int nrows, const int* __restrict__ rp,
                                                       const int* __restrict__ tmp_col, const int* __restrict__ tmp_src,
                                                       const T* __restrict__ val_src, int* col, T* val, int* map) {
  const int lane = threadIdx.x & 31, wpb = kBlock >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nrows; i += gridDim.x * wpb) {
    const int base = rp[i], len = rp[i + 1] - base;
    if (len <= 32) {
      const int c = lane < len ? tmp_col[base + lane] : 0x7fffffff;
      const int s = lane < len ? tmp_src[base + lane] : 0x7fffffff;
      int rank = 0;
      for (int t = 0; t < len; t++) {
        const int ct = __shfl_sync(0xffffffffu, c, t), st = __shfl_sync(0xffffffffu, s, t);
        rank += (ct < c) || (ct == c && st < s);
      }
      if (lane < len) {
        col[base + rank] = c;
        val[base + rank] = val_src[s];
        if (map) map[s] = base + rank;
      }
    } else {
      for (int e = lane; e < len; e += 32) {
        const int c = tmp_col[base + e], s = tmp_src[base + e];
        int rank = 0;
        for (int t = 0; t < len; t++) {
          const int ct = tmp_col[base + t], st = tmp_src[base + t];
          rank += (ct < c) || (ct == c && st < s);
        }
        col[base + rank] = c;
        val[base + rank] = val_src[s];
        if (map) map[s] = base + rank;
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock) veci_gather_kernel(int* dst, const int* __restrict__ src,
                                                             const int* __restrict__ idx, int n) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(kBlock) scatter_nonneg_kernel(T* dst, const T* __restrict__ src,
                                                                const int* __restrict__ idx, int n) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int p = idx[i];
    if (p >= 0) dst[p] = src[i];
  }
}

// ---------------------------------------------------------------- symmetric expansion kernels
// CSC of the upper triangle == CSR of L = triu', row j holding the entries (i <= j) of column j.
// Full row j = [strictly-lower mirrors (L row j without its diagonal), in source order] ++
//              [explicit zero diagonal if the column has none] ++ [row j of U = L', ascending
//              columns: the upper triangle itself, diagonal first] -- the layout of
// algebra/b200/matrix.c full_from_triu, entry for entry.

// one warp per row: off[j] = off-diagonal entries of column j, nodiag[j] = 1 if no diagonal entry,
// len[j + 1] = length of the full row
__global__ void __launch_bounds__(kBlock) sym_count(int n, const int* __restrict__ rpL, const int* __restrict__ ciL,
                                                    const int* __restrict__ rpU, int* off, int* nodiag, int* len) {
  const int lane = threadIdx.x & 31, wpb = kBlock >> 5;
  for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n; j += gridDim.x * wpb) {
    int d = 0;
    const int s = rpL[j], e = rpL[j + 1];
    for (int k = s + lane; k < e; k += 32) d += (ciL[k] == j);
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) {
      off[j]     = (e - s) - d;
      nodiag[j]  = (d == 0);
      len[j + 1] = (e - s) - d + (d == 0) + (rpU[j + 1] - rpU[j]);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) len[0] = 0;
}

__global__ void __launch_bounds__(kBlock) sym_fill(int n, const int* __restrict__ rpL, const int* __restrict__ ciL,
                                                   const T* __restrict__ vL, const int* __restrict__ rpU,
                                                   const int* __restrict__ ciU, const T* __restrict__ vU,
                                                   const int* __restrict__ rpF, const int* __restrict__ off,
                                                   const int* __restrict__ nodiag, int* ciF, T* vF, int* map_l) {
  const int lane = threadIdx.x & 31, wpb = kBlock >> 5;
  for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n; j += gridDim.x * wpb) {
    const int base = rpF[j];
    // strictly-lower mirrors, in source order (warp-wide prefix count of the off-diagonal entries)
    int run = 0;
    const int s = rpL[j], e = rpL[j + 1];
    for (int k0 = s; k0 < e; k0 += 32) {
      const int k = k0 + lane;
      const int i = k < e ? ciL[k] : j;
      const bool isoff = (k < e) && (i != j);
      const unsigned bal = __ballot_sync(0xffffffffu, isoff);
      if (k < e) {
        if (isoff) {
          const int pos = base + run + __popc(bal & ((1u << lane) - 1u));
          ciF[pos] = i;
          vF[pos]  = vL[k];
          map_l[k] = pos;
        } else {
          map_l[k] = -1;
        }
      }
      run += __popc(bal);
    }
    const int o = off[j], nd = nodiag[j];
    if (nd && lane == 0) {
      ciF[base + o] = j;
      vF[base + o]  = (T)0;
    }
    const int us = rpU[j], ue = rpU[j + 1];
    for (int t = us + lane; t < ue; t += 32) {
      ciF[base + o + nd + (t - us)] = ciU[t];
      vF[base + o + nd + (t - us)]  = vU[t];
    }
  }
}

// map_u[k] = position of source entry k = (i, j) itself: in the upper part of full row i
__global__ void __launch_bounds__(kBlock) sym_map_u(int nnz, const int* __restrict__ ciL, const int* __restrict__ tmap,
                                                    const int* __restrict__ rpU, const int* __restrict__ rpF,
                                                    const int* __restrict__ off, const int* __restrict__ nodiag,
                                                    int* map_u) {
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const int i = ciL[k];
    map_u[k] = rpF[i] + off[i] + nodiag[i] + (tmap[k] - rpU[i]);
  }
}

}  // namespace

namespace {

// ---- row selection (OSQPMatrix_submatrix_byrows): flags -> 0/1, lengths of the kept rows, row copy
__global__ void __launch_bounds__(kBlock) sel_mark(int n, const int* __restrict__ flags, int* sel) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) sel[i] = flags[i] != 0;
}
// pos[i] = inclusive scan of sel: a kept row i becomes row pos[i] - 1; lens[new + 1] = its length
__global__ void __launch_bounds__(kBlock) sel_lens(int n, const int* __restrict__ flags, const int* __restrict__ pos,
                                                   const int* __restrict__ rp, int* lens) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (flags[i] != 0) lens[pos[i]] = rp[i + 1] - rp[i];
}
// one warp per source row
__global__ void __launch_bounds__(kBlock) sel_copy(int n, const int* __restrict__ flags, const int* __restrict__ pos,
                                                   const int* __restrict__ rp, const int* __restrict__ ci,
                                                   const T* __restrict__ v, const int* __restrict__ rp_red, int* ci_red,
                                                   T* v_red) {
  const int lane = threadIdx.x & 31, wpb = kBlock >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    if (flags[i] == 0) continue;
    const int s0 = rp[i], e0 = rp[i + 1], d0 = rp_red[pos[i] - 1];
    for (int k = s0 + lane; k < e0; k += 32) { ci_red[d0 + k - s0] = ci[k]; v_red[d0 + k - s0] = v[k]; }
  }
}

}  // namespace

namespace {

// prefix sum over data[0..n) in place (inclusive); optionally the largest input value
bool device_scan(int* d_data, int n, int* d_max, cudaStream_t st) {
  const int nchunks = (n + kScanChunk - 1) / kScanChunk;
  int* d_tot = nullptr;
  int* d_dummy = nullptr;
  bool ok = B200_CHECK(dev_malloc(&d_tot, sizeof(int) * ((size_t)nchunks + 1)));
  if (!d_max) {
    ok &= B200_CHECK(dev_malloc(&d_dummy, sizeof(int)));
    if (ok) ok &= B200_CHECK(cudaMemsetAsync(d_dummy, 0, sizeof(int), st));   // atomicMax target (initcheck)
  }
  if (ok) {
    scan_totals<<<nchunks, kScanBlock, 0, st>>>(n, d_data, d_tot, d_max ? d_max : d_dummy);
    count_launch();
    scan_offsets<<<1, kScanBlock, 0, st>>>(nchunks, d_tot);
    count_launch();
    scan_apply<<<nchunks, kScanBlock, 0, st>>>(n, d_data, d_tot);
    count_launch();
  }
  dev_free(d_tot);
  dev_free(d_dummy);
  return ok;
}

// Transpose of a device CSR given by raw arrays (nrows_src x ncols_src, nnz entries).  The result
// gets a row-block schedule only if `schedule`; *h_rp_out (optional) receives its row pointers.
b200_csr* transpose_impl(const int* d_rp_src, const int* d_col_src, const T* d_val_src, int nrows_src,
                         int ncols_src, int nnz, bool schedule, int** d_map_out) {
  Context& c = ctx();
  cudaStream_t st = c.stream;
  if (d_map_out) *d_map_out = nullptr;
  if (nnz <= 0 || ncols_src <= 0) return nullptr;
  const int nr = ncols_src, nc = nrows_src;
  b200_csr* M = new b200_csr();
  M->nrows = nr; M->ncols = nc; M->nnz = nnz;
  int *d_max = nullptr, *d_cursor = nullptr, *d_tcol = nullptr, *d_tsrc = nullptr, *d_map = nullptr;
  bool ok = true;
  ok &= B200_CHECK(dev_malloc(&M->d_row_ptr, sizeof(int) * ((size_t)nr + 2 * kPad)));
  ok &= B200_CHECK(dev_malloc(&M->d_col_ind, sizeof(int) * ((size_t)nnz + 2 * kPad)));
  ok &= B200_CHECK(dev_malloc(&M->d_val, sizeof(T) * ((size_t)nnz + 2 * kPad)));
  ok &= B200_CHECK(dev_malloc(&d_max, sizeof(int)));
  ok &= B200_CHECK(dev_malloc(&d_cursor, sizeof(int) * ((size_t)nr + 1)));
  ok &= B200_CHECK(dev_malloc(&d_tcol, sizeof(int) * ((size_t)nnz + 1)));
  ok &= B200_CHECK(dev_malloc(&d_tsrc, sizeof(int) * ((size_t)nnz + 1)));
  if (d_map_out) ok &= B200_CHECK(dev_malloc(&d_map, sizeof(int) * ((size_t)nnz + 1)));
  int h_max = 0, h_total = -1;
  std::vector<int> h_rp;
  if (ok) {
    ok &= B200_CHECK(cudaMemsetAsync(M->d_row_ptr, 0, sizeof(int) * ((size_t)nr + 1), st));
    ok &= B200_CHECK(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    const int cap = c.sm_count * 8;
    int g = (nnz + kBlock - 1) / kBlock;
    tr_count<<<g < cap ? g : cap, kBlock, 0, st>>>(nnz, d_col_src, M->d_row_ptr);
    count_launch();
    ok &= device_scan(M->d_row_ptr, nr + 1, d_max, st);
    ok &= B200_CHECK(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (schedule) {
      h_rp.resize((size_t)nr + 1);
      ok &= B200_CHECK(cudaMemcpyAsync(h_rp.data(), M->d_row_ptr, sizeof(int) * ((size_t)nr + 1), cudaMemcpyDeviceToHost, st));
    }
    ok &= B200_CHECK(cudaMemcpyAsync(&h_total, M->d_row_ptr + nr, sizeof(int), cudaMemcpyDeviceToHost, st));
    ok &= B200_CHECK(cudaMemcpyAsync(d_cursor, M->d_row_ptr, sizeof(int) * (size_t)nr, cudaMemcpyDeviceToDevice, st));
    ok &= B200_CHECK(cudaStreamSynchronize(st));
  }
  if (ok && h_max <= kSortMaxRow && h_total == nnz) {
    const int wpb = kBlock >> 5, cap = c.sm_count * 8;
    int g1 = (nc + wpb - 1) / wpb, g2 = (nr + wpb - 1) / wpb;
    tr_scatter<<<g1 < cap ? g1 : cap, kBlock, 0, st>>>(nc, d_rp_src, d_col_src, d_cursor, d_tcol, d_tsrc);
    count_launch();
    tr_sort_rows<<<g2 < cap ? g2 : cap, kBlock, 0, st>>>(nr, M->d_row_ptr, d_tcol, d_tsrc, d_val_src, M->d_col_ind,
                                                      M->d_val, d_map);
    count_launch();
    if (schedule) ok = b200_build_schedule(M, h_rp.data()) == 0;    // synchronises the stream
  } else {
    ok = false;
  }
  dev_free(d_max); dev_free(d_cursor); dev_free(d_tcol); dev_free(d_tsrc);
  if (!ok) {
    dev_free(d_map);
    b200_csr_destroy(M);
    return nullptr;
  }
  if (d_map_out) *d_map_out = d_map;
  return M;
}

}  // namespace

extern "C" {

// CSR of the transpose of Mt (Mt: r x c  ->  result: c x r), columns ascending inside every row.
// *d_map_out (optional) receives a device array with, for every stored entry k of Mt, its position
// in the result; the caller frees it with b200_free.  Returns NULL when the result has a row longer
// than kSortMaxRow or on failure -- the caller then keeps its host path.
b200_csr* b200_csr_transpose(const b200_csr* Mt, int** d_map_out) {
  if (d_map_out) *d_map_out = nullptr;
  if (!Mt) return nullptr;
  return transpose_impl(Mt->d_row_ptr, Mt->d_col_ind, Mt->d_val, Mt->nrows, Mt->ncols, Mt->nnz, true, d_map_out);
}

// The rows of M whose flag is non-zero, order preserved, as a new CSR with its row-block schedule -- on the
// device: flag scan, row-length scan, one warp per kept row (replaces the host filter of round 1; the
// reference's role: csr_submatrix_byrows, algebra/cuda/src/cuda_csr.cu:763-843).  NULL on failure or when
// nothing is kept (the caller keeps its host path).
b200_csr* b200_csr_select_rows(const b200_csr* M, const int* d_flags, int* nrows_out) {
  Context& c = ctx();
  cudaStream_t st = c.stream;
  if (nrows_out) *nrows_out = 0;
  if (!M || M->nrows <= 0 || M->nnz <= 0) return nullptr;
  const int n = M->nrows, cap = c.sm_count * 8;
  int* d_pos = nullptr;
  int h_kept = 0, h_nnz = 0;
  bool ok = B200_CHECK(dev_malloc(&d_pos, sizeof(int) * ((size_t)n + 1)));
  b200_csr* R = nullptr;
  int* d_rp = nullptr;
  if (ok) {
    int g = (n + kBlock - 1) / kBlock;
    sel_mark<<<g < cap ? g : cap, kBlock, 0, st>>>(n, d_flags, d_pos);
    count_launch();
    ok &= device_scan(d_pos, n, nullptr, st);
    ok &= B200_CHECK(cudaMemcpyAsync(&h_kept, d_pos + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    ok &= B200_CHECK(cudaStreamSynchronize(st));
  }
  if (ok && h_kept > 0) {
    ok &= B200_CHECK(dev_malloc(&d_rp, sizeof(int) * ((size_t)h_kept + 2 * kPad)));
    if (ok) {
      ok &= B200_CHECK(cudaMemsetAsync(d_rp, 0, sizeof(int) * ((size_t)h_kept + 1), st));
      int g = (n + kBlock - 1) / kBlock;
      sel_lens<<<g < cap ? g : cap, kBlock, 0, st>>>(n, d_flags, d_pos, M->d_row_ptr, d_rp);
      count_launch();
      ok &= device_scan(d_rp, h_kept + 1, nullptr, st);
      ok &= B200_CHECK(cudaMemcpyAsync(&h_nnz, d_rp + h_kept, sizeof(int), cudaMemcpyDeviceToHost, st));
      ok &= B200_CHECK(cudaStreamSynchronize(st));
    }
    if (ok && h_nnz > 0) {
      R = new b200_csr();
      R->nrows = h_kept; R->ncols = M->ncols; R->nnz = h_nnz;
      R->d_row_ptr = d_rp; d_rp = nullptr;
      ok &= B200_CHECK(dev_malloc(&R->d_col_ind, sizeof(int) * ((size_t)h_nnz + 2 * kPad)));
      ok &= B200_CHECK(dev_malloc(&R->d_val, sizeof(T) * ((size_t)h_nnz + 2 * kPad)));
      if (ok) {
        const int wpb = kBlock >> 5;
        int g = (n + wpb - 1) / wpb;
        sel_copy<<<g < cap ? g : cap, kBlock, 0, st>>>(n, d_flags, d_pos, M->d_row_ptr, M->d_col_ind, M->d_val,
                                                        R->d_row_ptr, R->d_col_ind, R->d_val);
        count_launch();
        std::vector<int> h_rp((size_t)h_kept + 1);
        ok &= B200_CHECK(cudaMemcpyAsync(h_rp.data(), R->d_row_ptr, sizeof(int) * ((size_t)h_kept + 1), cudaMemcpyDeviceToHost, st));
        ok &= B200_CHECK(cudaStreamSynchronize(st));
        if (ok) ok = b200_build_schedule(R, h_rp.data()) == 0;
      }
      if (!ok) { b200_csr_destroy(R); R = nullptr; }
    }
  }
  dev_free(d_pos);
  dev_free(d_rp);
  if (R && nrows_out) *nrows_out = h_kept;
  return R;
}

// Full symmetric CSR with a structurally full diagonal from the upper-triangular CSC arrays of P
// (host pointers), expanded on the device: upload the triangle once, transpose it, merge.
// *d_map_u / *d_map_l receive device arrays (nnz ints): position of every user entry itself and of
// its mirror (-1 for diagonal entries).  NULL when a row of the triangle's transpose is longer than
// the rank-sort limit (the caller keeps its host path) or on failure.
b200_csr* b200_csr_symmetric_from_triu(int n, const int* h_p, const int* h_i, const T* h_x, int** d_map_u,
                                       int** d_map_l) {
  Context& c = ctx();
  cudaStream_t st = c.stream;
  *d_map_u = *d_map_l = nullptr;
  if (n <= 0) return nullptr;
  const int nnz = h_p[n];
  if (nnz <= 0) return nullptr;
  int *rpL = nullptr, *ciL = nullptr, *off = nullptr, *nodiag = nullptr, *tmap = nullptr, *mu = nullptr, *ml = nullptr;
  T* vL = nullptr;
  b200_csr* U = nullptr;
  b200_csr* F = nullptr;
  bool ok = true;
  ok &= B200_CHECK(dev_malloc(&rpL, sizeof(int) * ((size_t)n + 2)));
  ok &= B200_CHECK(dev_malloc(&ciL, sizeof(int) * ((size_t)nnz + 1)));
  ok &= B200_CHECK(dev_malloc(&vL, sizeof(T) * ((size_t)nnz + 1)));
  ok &= B200_CHECK(dev_malloc(&off, sizeof(int) * ((size_t)n + 1)));
  ok &= B200_CHECK(dev_malloc(&nodiag, sizeof(int) * ((size_t)n + 1)));
  ok &= B200_CHECK(dev_malloc(&mu, sizeof(int) * ((size_t)nnz + 1)));
  ok &= B200_CHECK(dev_malloc(&ml, sizeof(int) * ((size_t)nnz + 1)));
  if (ok) {
    ok &= upload(rpL, h_p, sizeof(int) * ((size_t)n + 1));
    ok &= upload(ciL, h_i, sizeof(int) * (size_t)nnz);
    ok &= upload(vL, h_x, sizeof(T) * (size_t)nnz);
  }
  if (ok) U = transpose_impl(rpL, ciL, vL, n, n, nnz, false, &tmap);
  if (ok && U) {
    F = new b200_csr();
    F->nrows = n; F->ncols = n;
    ok &= B200_CHECK(dev_malloc(&F->d_row_ptr, sizeof(int) * ((size_t)n + 2 * kPad)));
    const int wpb = kBlock >> 5, cap = c.sm_count * 8;
    const int gr = (n + wpb - 1) / wpb < cap ? (n + wpb - 1) / wpb : cap;
    std::vector<int> h_rp((size_t)n + 1);
    if (ok) {
      sym_count<<<gr, kBlock, 0, st>>>(n, rpL, ciL, U->d_row_ptr, off, nodiag, F->d_row_ptr);
      count_launch();
      ok &= device_scan(F->d_row_ptr, n + 1, nullptr, st);
      ok &= B200_CHECK(cudaMemcpyAsync(h_rp.data(), F->d_row_ptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
      ok &= B200_CHECK(cudaStreamSynchronize(st));
    }
    if (ok) {
      F->nnz = h_rp[n];
      ok &= B200_CHECK(dev_malloc(&F->d_col_ind, sizeof(int) * ((size_t)F->nnz + 2 * kPad)));
      ok &= B200_CHECK(dev_malloc(&F->d_val, sizeof(T) * ((size_t)F->nnz + 2 * kPad)));
    }
    if (ok) {
      sym_fill<<<gr, kBlock, 0, st>>>(n, rpL, ciL, vL, U->d_row_ptr, U->d_col_ind, U->d_val, F->d_row_ptr, off, nodiag,
                                      F->d_col_ind, F->d_val, ml);
      count_launch();
      int g = (nnz + kBlock - 1) / kBlock;
      sym_map_u<<<g < cap ? g : cap, kBlock, 0, st>>>(nnz, ciL, tmap, U->d_row_ptr, F->d_row_ptr, off, nodiag, mu);
      count_launch();
      ok = b200_build_schedule(F, h_rp.data()) == 0;     // synchronises the stream
    }
  } else {
    ok = false;
  }
  dev_free(rpL); dev_free(ciL); dev_free(vL); dev_free(off); dev_free(nodiag); dev_free(tmap);
  b200_csr_destroy(U);
  if (!ok) {
    dev_free(mu); dev_free(ml);
    b200_csr_destroy(F);
    return nullptr;
  }
  *d_map_u = mu;
  *d_map_l = ml;
  return F;
}

// dst[i] = src[idx[i]] for integer arrays (index maps of OSQPMatrix_update_values)
void b200_veci_gather(int* dst, const int* src, const int* idx, int n) {
  if (n <= 0) return;
  const int cap = ctx().sm_count * 8;
  const int g = (n + kBlock - 1) / kBlock;
  veci_gather_kernel<<<g < cap ? g : cap, kBlock, 0, ctx().stream>>>(dst, src, idx, n);
  count_launch();
}

// dst[idx[i]] = src[i] where idx[i] >= 0 (mirror positions: -1 marks "no mirror")
void b200_vec_scatter_nonneg(T* dst, const T* src, const int* idx, int n) {
  if (n <= 0) return;
  const int cap = ctx().sm_count * 8;
  const int g = (n + kBlock - 1) / kBlock;
  scatter_nonneg_kernel<<<g < cap ? g : cap, kBlock, 0, ctx().stream>>>(dst, src, idx, n);
  count_launch();
}

}  // extern "C"
