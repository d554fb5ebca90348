// Micro-benchmark v2 (TMA-fed): "slice-blocked" SpMV (SB) -- the gathered source vector is staged slice by slice in
// SHARED memory (bulk async copies, double buffered), so that the random gathers hit the shared-memory
// crossbar (~6 wavefronts per warp for random f64) instead of the L1 tag stage (32 wavefronts per warp:
// the "one divergent gather per clock per SM" wall of profiles/r01b_leanpass_ncu.md).
//
//   panel  = (block of <= Rmax consecutive rows, range of column blocks); the row sums of a panel are
//            accumulated in shared memory and written once (final, or a partial when the row block's
//            columns are split over several panels -- folded in panel order by a combine kernel)
//   step   = one column block of a panel: either STAGED (the touched column range, <= W elements, is
//            copied to shared memory; 16-bit local column indices) or DIRECT (too few entries for the
//            range: 32-bit global column indices, gathered through L1 as before)
//   slab   = <= 1024 row segments of a step, sorted by length (stable) and stored as jagged diagonals:
//            thread i owns segment i, every load of the matrix stream is coalesced, no row pointers,
//            no shuffles, and no two threads of a step touch the same accumulator
//
// Built on the host here (development aid); matrices: A and K2 = [P + sigma I | A'] of the Lasso
// workload (BASELINE configs[1]: 1e5 features x 1e6 samples, density 1e-4).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <random>
#include <algorithm>
#include <numeric>
#include <cuda_runtime.h>

typedef double T;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)


// v2: v1 (spmv_sb.cu) was latency bound -- three dependent global loads per slab (descriptor, diagonal
// lengths, entries), one slab in flight per thread: 105 us for the A pass against 67 us for the shipped
// kernel.  Here EVERYTHING a step needs -- its slice of the source vector and one contiguous "blob" holding
// the values, 16-bit column indices, segment rows and slab headers -- is brought into a shared-memory ring
// by two bulk async copies (cp.async.bulk + mbarrier complete_tx) issued one step ahead; the 1024 consumer
// threads only ever touch shared memory.
struct SbPanel { int row0, nrows, step0, step1, part_off; };    // part_off < 0: final output
struct SbStep  { int col0, len, src_sel, direct; long long blob_off; int blob_bytes, off_meta, off_col, off_seg; };
struct SbComb  { int row0, nrows, part_off, nparts; };
struct SbView { const SbPanel* panels; const SbStep* steps; const unsigned char* blobs; int npanels, csplit; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double ld_stream(const double* p) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ int ld_stream(const int* p) { int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ int ld_stream(const uint16_t* p) { unsigned short v; asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p)); return (int)v; }


constexpr int kSbBlock = 1024;

// dynamic shared memory: NST stages of [W doubles slice | SMAX bytes blob], Rmax accumulators, mbarriers
template <int NST>
__global__ void __launch_bounds__(kSbBlock, 1) sb_pass(SbView M, const T* __restrict__ s0, const T* __restrict__ s1,
                                                       T* __restrict__ y, T* __restrict__ part, int W, int SMAX, int Rmax,
                                                       int* __restrict__ counter) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const int stage_bytes = W * (int)sizeof(T) + SMAX;
  T* acc = reinterpret_cast<T*>(dsm + (size_t)NST * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(acc + Rmax);
  int* s_next = reinterpret_cast<int*>(bars + NST);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < NST; i++) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned k_use = 0;     // steps consumed so far by this CTA: stage = k % NST, parity = (k / NST) & 1
  auto issue = [&](const SbStep& st, unsigned k) {
    unsigned char* stage = dsm + (size_t)(k % NST) * stage_bytes;
    uint64_t* bar = &bars[k % NST];
    const uint32_t sb = (uint32_t)st.len * sizeof(T);
    mbar_expect_tx(bar, sb + (uint32_t)st.blob_bytes);
    if (sb) bulk_load(stage, st.src_sel ? (const void*)(s1 + (st.col0 - M.csplit)) : (const void*)(s0 + st.col0), sb, bar);
    bulk_load(stage + W * sizeof(T), M.blobs + st.blob_off, (uint32_t)st.blob_bytes, bar);
  };
  for (;;) {
    if (tid == 0) *s_next = atomicAdd(counter, 1);
    __syncthreads();
    const int p = *s_next;
    if (p >= M.npanels) break;
    const SbPanel pn = M.panels[p];
    for (int i = tid; i < pn.nrows; i += kSbBlock) acc[i] = 0;
    const int nsteps = pn.step1 - pn.step0;
    if (tid == 0)
      for (int j = 0; j < NST - 1 && j < nsteps; j++) issue(M.steps[pn.step0 + j], k_use + j);
    __syncthreads();
    for (int s = 0; s < nsteps; s++, k_use++) {
      const SbStep st = M.steps[pn.step0 + s];
      if (tid == 0 && s + NST - 1 < nsteps) issue(M.steps[pn.step0 + s + NST - 1], k_use + NST - 1);
      const unsigned char* stage = dsm + (size_t)(k_use % NST) * stage_bytes;
      mbar_wait(&bars[k_use % NST], (k_use / NST) & 1);
      const T* xs = reinterpret_cast<const T*>(stage);
      const unsigned char* blob = stage + W * sizeof(T);
      const T* sval = reinterpret_cast<const T*>(blob);
      const int* meta = reinterpret_cast<const int*>(blob + st.off_meta);
      const uint16_t* scol16 = reinterpret_cast<const uint16_t*>(blob + st.off_col);
      const int* scol32 = reinterpret_cast<const int*>(blob + st.off_col);
      const uint16_t* sseg = reinterpret_cast<const uint16_t*>(blob + st.off_seg);
      const int nslab = meta[0];
      int mp = 1;
      for (int sl = 0; sl < nslab; sl++) {
        const int ent_off = meta[mp], nseg = meta[mp + 1], seg_off = meta[mp + 2], maxlen = meta[mp + 3];
        const int* dl = meta + mp + 4;
        mp += 4 + maxlen;
        if (tid < nseg) {
          T sum = 0;
          int off = ent_off + tid;
          if (!st.direct) {
            for (int j = 0; j < maxlen; j++) {
              const int d = dl[j];
              if (tid >= d) break;
              sum += sval[off] * xs[scol16[off]];
              off += d;
            }
          } else {
            for (int j = 0; j < maxlen; j++) {
              const int d = dl[j];
              if (tid >= d) break;
              const int c = scol32[off];
              sum += sval[off] * (c < M.csplit ? __ldg(s0 + c) : __ldg(s1 + (c - M.csplit)));
              off += d;
            }
          }
          acc[sseg[seg_off + tid]] += sum;
        }
      }
      __syncthreads();   // slice / blob of this stage fully consumed, accumulator updates of this step done
    }
    if (pn.part_off < 0) {
      for (int i = tid; i < pn.nrows; i += kSbBlock) y[pn.row0 + i] = acc[i];
    } else {
      for (int i = tid; i < pn.nrows; i += kSbBlock) part[pn.part_off + i] = acc[i];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sb_combine(const SbComb* __restrict__ cb, int ncomb, const T* __restrict__ part,
                                                  T* __restrict__ y) {
  for (int b = blockIdx.y; b < ncomb; b += gridDim.y) {
    const SbComb c = cb[b];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.nrows; i += gridDim.x * blockDim.x) {
      T s = 0;
      for (int k = 0; k < c.nparts; k++) s += ld_stream(part + c.part_off + (size_t)k * c.nrows + i);
      y[c.row0 + i] = s;
    }
  }
}

// ------------------------------------------------------------------ baseline: the shipped flat CSR-stream tile
struct Desc { int row0, nrows, nnz0, cnt, lg; };
template <int BLOCK, int TILE, int KU>
__global__ void __launch_bounds__(BLOCK) k_tile(const int* __restrict__ rp, const int* __restrict__ ci,
                                                const T* __restrict__ va, const Desc* __restrict__ desc, int nblocks,
                                                const T* __restrict__ s0, const T* __restrict__ s1, int csplit, T* __restrict__ y) {
  __shared__ T sm[TILE];
  __shared__ int srp[TILE / 2 + 1];
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const Desc d = desc[b];
    for (int i = tid; i <= d.nrows; i += BLOCK) srp[i] = ld_stream(rp + d.row0 + i) - d.nnz0;
    int c[KU]; T v[KU];
#pragma unroll
    for (int u = 0; u < KU; u++) { int k = u * BLOCK + tid; if (k < d.cnt) { c[u] = ld_stream(ci + d.nnz0 + k); v[u] = ld_stream(va + d.nnz0 + k); } }
#pragma unroll
    for (int u = 0; u < KU; u++) { int k = u * BLOCK + tid; if (k < d.cnt) sm[k] = v[u] * (c[u] < csplit ? __ldg(s0 + c[u]) : __ldg(s1 + (c[u] - csplit))); }
    __syncthreads();
    const int g = 1 << d.lg, gid = tid >> d.lg, lig = tid & (g - 1), ng = BLOCK >> d.lg;
    for (int base = 0; base < d.nrows; base += ng) {
      int r = base + gid; T a = 0;
      if (r < d.nrows) { int e = srp[r + 1]; for (int k = srp[r] + lig; k < e; k += g) a += sm[k]; }
      for (int o = g >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (r < d.nrows && lig == 0) y[d.row0 + r] = a;
    }
    __syncthreads();
  }
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


// ------------------------------------------------------------------ host builder
struct SbHost {
  std::vector<SbPanel> panels; std::vector<SbStep> steps; std::vector<SbComb> combs; std::vector<unsigned char> blobs;
  size_t part_elems = 0; int csplit = 0;
  long long staged_entries = 0, direct_entries = 0, slice_elems = 0, nseg = 0, nslab = 0;
};
struct Ent { int cb; int rl; int c; T v; };

// one step from the entries [b, en) (row order): JDS slabs packed into a blob; returns false if it does
// not fit SMAX (the caller then splits the entries)
static bool emit_step(SbHost& H, const Ent* b, const Ent* en, bool staged, int col0, int len, int sel, int SMAX) {
  struct Seg { int rl, first, n; };
  std::vector<Seg> segs;
  for (const Ent* q = b; q < en;) { const Ent* q2 = q; while (q2 < en && q2->rl == q->rl) q2++; segs.push_back({q->rl, (int)(q - b), (int)(q2 - q)}); q = q2; }
  std::stable_sort(segs.begin(), segs.end(), [](const Seg& a, const Seg& c) { return a.n > c.n; });
  const int nent = (int)(en - b);
  std::vector<T> val; std::vector<uint16_t> c16; std::vector<int> c32; std::vector<uint16_t> segrow; std::vector<int> meta{0};
  int nslab = 0;
  for (size_t s0 = 0; s0 < segs.size(); s0 += kSbBlock) {
    const int ns = (int)std::min<size_t>(kSbBlock, segs.size() - s0);
    const int maxlen = segs[s0].n;
    meta.push_back((int)val.size()); meta.push_back(ns); meta.push_back((int)segrow.size()); meta.push_back(maxlen);
    for (int i = 0; i < ns; i++) segrow.push_back((uint16_t)segs[s0 + i].rl);
    for (int j = 0; j < maxlen; j++) {
      int dl = 0;
      for (int i = 0; i < ns && segs[s0 + i].n > j; i++) {
        const Ent& q = b[segs[s0 + i].first + j];
        val.push_back(q.v); c16.push_back((uint16_t)(q.c - col0)); c32.push_back(q.c); dl++;
      }
      meta.push_back(dl);
    }
    nslab++;
  }
  meta[0] = nslab;
  auto al = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
  const size_t off_meta = (size_t)nent * 8;
  const size_t off_col = al(off_meta + meta.size() * 4, 8);
  const size_t off_seg = al(off_col + (staged ? (size_t)nent * 2 : (size_t)nent * 4), 4);
  const size_t bytes = al(off_seg + segrow.size() * 2, 16);
  if ((int)bytes > SMAX) return false;
  SbStep st; st.col0 = col0; st.len = staged ? len : 0; st.src_sel = sel; st.direct = staged ? 0 : 1;
  st.blob_off = (long long)H.blobs.size(); st.blob_bytes = (int)bytes; st.off_meta = (int)off_meta; st.off_col = (int)off_col; st.off_seg = (int)off_seg;
  H.blobs.resize(H.blobs.size() + bytes, 0);
  unsigned char* o = H.blobs.data() + st.blob_off;
  memcpy(o, val.data(), val.size() * 8);
  memcpy(o + off_meta, meta.data(), meta.size() * 4);
  if (staged) memcpy(o + off_col, c16.data(), c16.size() * 2); else memcpy(o + off_col, c32.data(), c32.size() * 4);
  memcpy(o + off_seg, segrow.data(), segrow.size() * 2);
  H.steps.push_back(st);
  H.nseg += (long long)segs.size(); H.nslab += nslab;
  if (staged) { H.staged_entries += nent; H.slice_elems += len; } else H.direct_entries += nent;
  return true;
}

// entries [b, en) of one column block (row order): staged if the touched range is short enough for its
// entry count, split in two column halves while the blob does not fit; the rest is returned as direct
static void emit_block(SbHost& H, std::vector<Ent> v, int csplit, int SMAX, std::vector<Ent>& direct) {
  if (v.empty()) return;
  int cmin = v[0].c, cmax = v[0].c;
  for (const Ent& q : v) { cmin = std::min(cmin, q.c); cmax = std::max(cmax, q.c); }
  const int sel = cmin >= csplit;
  const int col0 = sel ? csplit + ((cmin - csplit) & ~1) : (cmin & ~1);
  const int len = (cmax - col0 + 1 + 1) & ~1;
  if ((long long)len > 16 * (long long)v.size()) { direct.insert(direct.end(), v.begin(), v.end()); return; }
  if (emit_step(H, v.data(), v.data() + v.size(), true, col0, len, sel, SMAX)) return;
  const int mid = (cmin + cmax + 1) / 2;
  std::vector<Ent> lo, hi;
  for (const Ent& q : v) (q.c < mid ? lo : hi).push_back(q);
  if (lo.empty() || hi.empty()) {   // one very dense column: split by rows instead
    lo.assign(v.begin(), v.begin() + v.size() / 2); hi.assign(v.begin() + v.size() / 2, v.end());
  }
  emit_block(H, std::move(lo), csplit, SMAX, direct);
  emit_block(H, std::move(hi), csplit, SMAX, direct);
}
static void emit_direct(SbHost& H, const Ent* b, const Ent* en, int SMAX) {
  if (b >= en) return;
  if (emit_step(H, b, en, false, 0, 0, 0, SMAX)) return;
  const Ent* mid = b + (en - b) / 2;
  while (mid < en && mid > b && mid->rl == (mid - 1)->rl) mid++;    // keep a row's entries together
  if (mid == en) { mid = b + (en - b) / 2; }
  emit_direct(H, b, mid, SMAX);
  emit_direct(H, mid, en, SMAX);
}

SbHost sb_build(const std::vector<int>& rp, const std::vector<int>& ci, const std::vector<T>& va, int ncols, int csplit,
                int Rmax, int W, int SMAX, long long target_nnz) {
  SbHost H; H.csplit = csplit;
  const int nrows = (int)rp.size() - 1;
  const int nb0 = (csplit + W - 1) / W;
  auto cblock = [&](int c) { return c < csplit ? c / W : nb0 + (c - csplit) / W; };
  const int nb = nb0 + (ncols - csplit + W - 1) / W + 1;
  std::vector<long long> cnt(nb);
  std::vector<Ent> ents;
  for (int r0 = 0; r0 < nrows; r0 += Rmax) {
    const int nr = std::min(Rmax, nrows - r0);
    std::fill(cnt.begin(), cnt.end(), 0);
    ents.clear();
    for (int r = r0; r < r0 + nr; r++)
      for (int k = rp[r]; k < rp[r + 1]; k++) { int cb = cblock(ci[k]); cnt[cb]++; ents.push_back({cb, r - r0, ci[k], va[k]}); }
    const long long nnz_rb = (long long)ents.size();
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.cb < b.cb; });
    int nsplit = (int)std::max(1LL, (nnz_rb + target_nnz / 2) / target_nnz);
    std::vector<int> cuts{0};
    { long long accn = 0; int part = 1;
      for (int cb = 0; cb < nb; cb++) { accn += cnt[cb]; if (part < nsplit && accn >= nnz_rb * part / nsplit) { cuts.push_back(cb + 1); part++; } }
      cuts.push_back(nb); }
    const int nparts = (int)cuts.size() - 1;
    size_t comb_off = H.part_elems;
    if (nparts > 1) { H.combs.push_back({r0, nr, (int)comb_off, nparts}); H.part_elems += (size_t)nparts * nr; }
    size_t e = 0;
    for (int pi = 0; pi < nparts; pi++) {
      SbPanel pn; pn.row0 = r0; pn.nrows = nr; pn.step0 = (int)H.steps.size();
      pn.part_off = nparts > 1 ? (int)(comb_off + (size_t)pi * nr) : -1;
      std::vector<Ent> direct;
      for (int cb = cuts[pi]; cb < cuts[pi + 1]; cb++) {
        if (!cnt[cb]) continue;
        emit_block(H, std::vector<Ent>(ents.begin() + e, ents.begin() + e + cnt[cb]), csplit, SMAX, direct);
        e += cnt[cb];
      }
      if (!direct.empty()) {
        std::stable_sort(direct.begin(), direct.end(), [](const Ent& a, const Ent& c) { return a.rl < c.rl; });
        emit_direct(H, direct.data(), direct.data() + direct.size(), SMAX);
      }
      pn.step1 = (int)H.steps.size();
      H.panels.push_back(pn);
    }
  }
  std::vector<long long> w(H.panels.size());
  for (size_t i = 0; i < H.panels.size(); i++) {
    long long c = 0;
    for (int s = H.panels[i].step0; s < H.panels[i].step1; s++) c += H.steps[s].blob_bytes / 10 + 300;
    w[i] = c + H.panels[i].nrows / 4;
  }
  std::vector<int> ord(H.panels.size()); std::iota(ord.begin(), ord.end(), 0);
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return w[a] > w[b]; });
  std::vector<SbPanel> p2; for (int i : ord) p2.push_back(H.panels[i]); H.panels.swap(p2);
  return H;
}

// host emulation of sb_pass + sb_combine: validates the builder without a GPU (./spmv_sb2 cpu)
double sb_emulate(const SbHost& H, const std::vector<T>& src, const std::vector<double>& yref, int W, int SMAX, int Rmax) {
  std::vector<double> y(yref.size(), 0.0), part(H.part_elems + 1, 0.0), acc(Rmax);
  for (const SbPanel& pn : H.panels) {
    std::fill(acc.begin(), acc.end(), 0.0);
    for (int s = pn.step0; s < pn.step1; s++) {
      const SbStep& st = H.steps[s];
      if (st.len > W || (st.len & 1) || (st.col0 & 1) || st.blob_bytes > SMAX || (st.blob_bytes & 15) || (st.blob_off & 15)) { printf("bad step\n"); exit(1); }
      const unsigned char* blob = H.blobs.data() + st.blob_off;
      const T* sval = (const T*)blob; const int* meta = (const int*)(blob + st.off_meta);
      const uint16_t* c16 = (const uint16_t*)(blob + st.off_col); const int* c32 = (const int*)(blob + st.off_col);
      const uint16_t* sseg = (const uint16_t*)(blob + st.off_seg);
      std::vector<char> seen(pn.nrows, 0);
      int mp = 1;
      for (int sl = 0; sl < meta[0]; sl++) {
        const int ent_off = meta[mp], nseg = meta[mp + 1], seg_off = meta[mp + 2], maxlen = meta[mp + 3];
        const int* dl = meta + mp + 4; mp += 4 + maxlen;
        for (int i = 0; i < nseg; i++) {
          double sum = 0; int off = ent_off + i;
          for (int j = 0; j < maxlen; j++) {
            if (i >= dl[j]) break;
            if (!st.direct) { if (c16[off] >= st.len) { printf("col16 out of slice\n"); exit(1); } sum += sval[off] * src[st.col0 + c16[off]]; }
            else sum += sval[off] * src[c32[off]];
            off += dl[j];
          }
          const int r = sseg[seg_off + i];
          if (r >= pn.nrows || seen[r]) { printf("row %d twice in one step\n", r); exit(1); }
          seen[r] = 1; acc[r] += sum;
        }
      }
    }
    for (int i = 0; i < pn.nrows; i++) { if (pn.part_off < 0) y[pn.row0 + i] = acc[i]; else part[pn.part_off + i] = acc[i]; }
  }
  for (const SbComb& c : H.combs)
    for (int i = 0; i < c.nrows; i++) { double s2 = 0; for (int k = 0; k < c.nparts; k++) s2 += part[c.part_off + (size_t)k * c.nrows + i]; y[c.row0 + i] = s2; }
  double err = 0;
  for (size_t r = 0; r < y.size(); r++) err = std::max(err, std::abs(y[r] - yref[r]) / (1.0 + std::abs(yref[r])));
  return err;
}

bool g_cpu_only = false;
template <class V> typename V::value_type* up(const V& v) {
  typename V::value_type* p; CK(cudaMalloc(&p, (v.size() + 64) * sizeof(typename V::value_type)));
  CK(cudaMemcpy(p, v.data(), v.size() * sizeof(typename V::value_type), cudaMemcpyHostToDevice)); return p;
}

char* g_flush = nullptr; const size_t kFlushBytes = 512u << 20;
template <class F> float timeit(F f, int reps, bool flush) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) f();
  float tot = 0;
  for (int i = 0; i < reps; i++) {
    if (flush) CK(cudaMemsetAsync(g_flush, i, kFlushBytes));
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms;
  }
  CK(cudaGetLastError());
  return tot / reps * 1e3f;
}

struct Csr { std::vector<int> rp, ci; std::vector<T> va; int nrows, ncols; };
struct Cfg { int Rmax, W, SMAX, NST; double tf; };

void run_matrix(const char* name, const Csr& M, int csplit, const std::vector<T>& src, double alg_bytes) {
  const int R = M.nrows, nnz = (int)M.ci.size();
  std::vector<double> yref(R);
  for (int r = 0; r < R; r++) { double a = 0; for (int k = M.rp[r]; k < M.rp[r + 1]; k++) a += M.va[k] * src[M.ci[k]]; yref[r] = a; }
  const long long target = (long long)nnz / 148;
  const Cfg cfgs[] = {{8192, 4096, 40960, 2, 1.0}, {8192, 2048, 20480, 4, 1.0}, {8192, 1536, 12288, 6, 1.0}, {8192, 3072, 24576, 3, 1.0},
                      {6144, 2048, 24576, 4, 1.0}, {8192, 1024, 16384, 6, 1.0}, {10240, 2048, 16384, 4, 1.0}, {8192, 4096, 40960, 2, 0.5},
                      {8192, 2048, 20480, 4, 0.5}, {6144, 3072, 32768, 3, 1.0}};
  if (g_cpu_only) {
    for (const Cfg& c : {cfgs[0], cfgs[1], cfgs[2]}) {
      SbHost H = sb_build(M.rp, M.ci, M.va, M.ncols, csplit, c.Rmax, c.W, c.SMAX, target);
      printf("%s cpu emulation Rmax %d W %d SMAX %d: %zu panels %zu steps %lld slabs %zu combs, staged %lld direct %lld nseg %lld slices %.1f MB blobs %.1f MB, relerr %.2e\n",
             name, c.Rmax, c.W, c.SMAX, H.panels.size(), H.steps.size(), H.nslab, H.combs.size(), H.staged_entries, H.direct_entries, H.nseg,
             H.slice_elems * 8e-6, H.blobs.size() * 1e-6, sb_emulate(H, src, yref, c.W, c.SMAX, c.Rmax));
      { // per-panel statistics
        int worst = 0; long long worst_bytes = 0; int nd = 0;
        for (size_t i = 0; i < H.panels.size(); i++) {
          long long by = 0; int dsteps = 0; long long dbytes = 0;
          for (int s = H.panels[i].step0; s < H.panels[i].step1; s++) { by += H.steps[s].blob_bytes; if (H.steps[s].direct) { dsteps++; dbytes += H.steps[s].blob_bytes; } }
          if (i < 6 || dsteps > 0 && nd++ < 6) printf("   panel %zu: rows %d..+%d steps %d blob bytes %lld direct steps %d (%lld bytes) part %d\n", i, H.panels[i].row0, H.panels[i].nrows, H.panels[i].step1 - H.panels[i].step0, by, dsteps, dbytes, H.panels[i].part_off);
          if (by > worst_bytes) { worst_bytes = by; worst = (int)i; }
        }
        printf("   heaviest panel %d: %lld blob bytes, %d steps\n", worst, worst_bytes, H.panels[worst].step1 - H.panels[worst].step0);
      }
    }
    return;
  }
  T *dsrc = up(src), *dy; CK(cudaMalloc(&dy, (R + 64) * sizeof(T)));
  const T* s0 = dsrc; const T* s1 = dsrc + csplit;
  auto check = [&](const char* what, float us, float us_flush) {
    std::vector<T> hy(R); CK(cudaMemcpy(hy.data(), dy, R * sizeof(T), cudaMemcpyDeviceToHost));
    double err = 0; for (int r = 0; r < R; r++) err = std::max(err, std::abs(hy[r] - yref[r]) / (1.0 + std::abs(yref[r])));
    printf("%-3s %-96s %7.1f us (L2 flushed %7.1f us)  %5.0f GB/s alg = %4.1f%% of 6541   relerr %.1e\n", name, what, us, us_flush,
           alg_bytes / us_flush / 1e3, alg_bytes / us_flush / 1e3 / 65.41, err);
    fflush(stdout);
    CK(cudaMemset(dy, 0, R * sizeof(T)));
  };
  {
    int *drp = up(M.rp), *dci = up(M.ci); T* dva = up(M.va);
    auto d = schedule(M.rp, 2048, 1024, 512); Desc* dd = up(d); int nbk = (int)d.size();
    auto f = [&] { k_tile<512, 2048, 4><<<nbk, 512>>>(drp, dci, dva, dd, nbk, s0, s1, csplit, dy); };
    float a = timeit(f, 20, false), b = timeit(f, 20, true);
    check("baseline flat CSR-stream tile 2048 x 512 thr, CTA per tile", a, b);
    cudaFree(drp); cudaFree(dci); cudaFree(dva); cudaFree(dd);
  }
  int* counter; CK(cudaMalloc(&counter, 4));
  for (const Cfg& c : cfgs) {
    SbHost H = sb_build(M.rp, M.ci, M.va, M.ncols, csplit, c.Rmax, c.W, c.SMAX, (long long)(target * c.tf));
    SbView V; V.panels = up(H.panels); V.steps = up(H.steps); V.blobs = up(H.blobs); V.npanels = (int)H.panels.size(); V.csplit = csplit;
    SbComb* dcomb = up(H.combs); T* dpart; CK(cudaMalloc(&dpart, (H.part_elems + 64) * sizeof(T)));
    const size_t smem = (size_t)c.NST * (c.W * sizeof(T) + c.SMAX) + (size_t)c.Rmax * sizeof(T) + 64;
    if (smem > 232448) { printf("%s cfg skipped: %zu bytes of shared memory\n", name, smem); continue; }
    if (c.NST == 2) CK(cudaFuncSetAttribute(sb_pass<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (c.NST == 3) CK(cudaFuncSetAttribute(sb_pass<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (c.NST == 4) CK(cudaFuncSetAttribute(sb_pass<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CK(cudaFuncSetAttribute(sb_pass<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min(148, V.npanels), ncomb = (int)H.combs.size();
    auto f = [&] {
      cudaMemsetAsync(counter, 0, 4);
      if (c.NST == 2) sb_pass<2><<<grid, kSbBlock, smem>>>(V, s0, s1, dy, dpart, c.W, c.SMAX, c.Rmax, counter);
      else if (c.NST == 3) sb_pass<3><<<grid, kSbBlock, smem>>>(V, s0, s1, dy, dpart, c.W, c.SMAX, c.Rmax, counter);
      else if (c.NST == 4) sb_pass<4><<<grid, kSbBlock, smem>>>(V, s0, s1, dy, dpart, c.W, c.SMAX, c.Rmax, counter);
      else sb_pass<6><<<grid, kSbBlock, smem>>>(V, s0, s1, dy, dpart, c.W, c.SMAX, c.Rmax, counter);
      if (ncomb) sb_combine<<<dim3(8, std::min(ncomb, 64)), 256>>>(dcomb, ncomb, dpart, dy);
    };
    float a = timeit(f, 20, false), b = timeit(f, 20, true);
    char what[200];
    snprintf(what, sizeof what, "SB2 Rmax %5d W %4d SMAX %5d x%d tf %.1f: %3d panels %5zu steps %6lld slabs, %.0f%% staged, blobs %.0f MB + slices %.0f MB + parts %.0f MB",
             c.Rmax, c.W, c.SMAX, c.NST, c.tf, V.npanels, H.steps.size(), H.nslab, 100.0 * H.staged_entries / nnz, H.blobs.size() / 1e6,
             H.slice_elems * 8.0 / 1e6, H.part_elems * 16.0 / 1e6);
    check(what, a, b);
    cudaFree((void*)V.panels); cudaFree((void*)V.steps); cudaFree((void*)V.blobs); cudaFree(dcomb); cudaFree(dpart);
  }
  cudaFree(dsrc); cudaFree(dy); cudaFree(counter);
}

int main(int argc, char** argv) {
  const int nf = 100000, ms_ = 1000000;
  g_cpu_only = argc > 1 && strcmp(argv[argc - 1], "cpu") == 0;
  if (!g_cpu_only) CK(cudaMalloc(&g_flush, kFlushBytes));
  std::mt19937 rng(1);
  std::poisson_distribution<int> pois(10.0);
  std::uniform_int_distribution<int> col(0, nf - 1);
  std::uniform_real_distribution<double> uni(0, 1);
  // A = [Ad -I 0; I 0 -I; I 0 I]   (docs/examples/lasso.rst:57-59), variables (x, y, t)
  Csr A; A.rp.push_back(0);
  for (int i = 0; i < ms_; i++) {
    int k = pois(rng); std::vector<int> cs(k); for (auto& c : cs) c = col(rng);
    std::sort(cs.begin(), cs.end()); cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
    for (int c : cs) { A.ci.push_back(c); A.va.push_back(uni(rng)); }
    A.ci.push_back(nf + i); A.va.push_back(-1.0); A.rp.push_back((int)A.ci.size());
  }
  for (int s = 0; s < 2; s++) for (int j = 0; j < nf; j++) { A.ci.push_back(j); A.va.push_back(1.0); A.ci.push_back(nf + ms_ + j); A.va.push_back(s ? 1.0 : -1.0); A.rp.push_back((int)A.ci.size()); }
  const int m = (int)A.rp.size() - 1, n = nf + ms_ + nf, nnz = (int)A.ci.size();
  A.nrows = m; A.ncols = n;
  // K2 = [P + sigma I | A'] : n x (n + m), P = diag(0, I, 0)
  Csr K; K.nrows = n; K.ncols = n + m; K.rp.assign(n + 1, 0);
  { std::vector<int> cntc(n, 0); for (int c : A.ci) cntc[c]++;
    for (int i = 0; i < n; i++) K.rp[i + 1] = K.rp[i] + 1 + cntc[i];
    K.ci.resize(K.rp[n]); K.va.resize(K.rp[n]);
    std::vector<int> cur(n);
    for (int i = 0; i < n; i++) { K.ci[K.rp[i]] = i; K.va[K.rp[i]] = (i >= nf && i < nf + ms_ ? 1.0 : 0.0) + 1e-6; cur[i] = K.rp[i] + 1; }
    for (int r = 0; r < m; r++) for (int k = A.rp[r]; k < A.rp[r + 1]; k++) { int c = A.ci[k]; K.ci[cur[c]] = n + r; K.va[cur[c]] = A.va[k]; cur[c]++; } }
  const double bytesA = nnz * 12.0 + (m + 1) * 4.0 + n * 8.0 + m * 8.0;
  const double bytesK = K.ci.size() * 12.0 + (n + 1) * 4.0 + (n + m) * 8.0 + n * 8.0;
  printf("A: %d x %d nnz %d (%.1f MB algorithmic)   K2: %d x %d nnz %zu (%.1f MB algorithmic)\n", m, n, nnz, bytesA / 1e6, n, n + m,
         K.ci.size(), bytesK / 1e6);
  std::vector<T> x(n), pt(n + m);
  for (auto& v : x) v = uni(rng);
  for (auto& v : pt) v = uni(rng);
  const char* only = argc > 1 ? argv[1] : "";
  if (strcmp(only, "K") != 0) run_matrix("A", A, n, x, bytesA);
  if (strcmp(only, "A") != 0) run_matrix("K2", K, n, pt, bytesK);
  return 0;
}
