// common.cuh -- library context, launch helpers, warp/block reduction primitives.
// sm_100a only.  No cuBLAS / cuSPARSE / thrust.
#pragma once

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#include "osqp_b200.h"

typedef b200_float T;

namespace b200 {

constexpr int kBlock        = 256;    // threads per CTA for every kernel in the library
constexpr int kMaxRedBlocks = 1184;   // 148 SMs x 8 CTAs: upper bound for reduction grids
constexpr int kScalarSlots  = 32;
constexpr int kMaxContexts  = 64;     // library contexts (= host threads with live solvers) per process
constexpr int kMailSlots    = 32;     // result values; the sequence word lives at index kMailSlots

struct Context {
  int          refcount   = 0;
  int          device     = -1;
  int          sm_count   = 148;
  cudaStream_t stream     = nullptr;
  // reduction workspace (device): per-CTA partials, ticket counter
  double*      d_partials = nullptr;   // kMaxRedBlocks * 4 doubles
  unsigned*    d_ticket   = nullptr;
  // result scalars: device slot + pinned host mirror
  double*      d_scalar   = nullptr;
  double*      h_scalar   = nullptr;   // pinned
  // zero-copy result mailbox (mapped pinned memory): the last CTA of a reduction writes the
  // value(s) and then a sequence number straight into host memory; the host spins on the
  // sequence number instead of paying cudaMemcpyAsync + cudaStreamSynchronize (~15 us a call)
  double*      h_mail     = nullptr;   // kMailSlots doubles + 1 sequence word, host view
  double*      d_mail     = nullptr;   // device view of the same memory
  unsigned long long mail_seq = 0;
  // staging buffer for host->device index/value uploads is allocated on demand
  int          last_error = 0;
  unsigned long long launches = 0;
  unsigned long long graph_launches = 0;   // cudaGraphLaunch calls (each = 1 + 3 k kernels, k CG iterations)
  unsigned long long epoch    = 0;   // launches + device copies: anything that may change a vector
  char         name[256]  = {0};
  int          trace_on   = 0;   // B200_TRACE_FILE: one CUDA event + host timestamp per launch
  int          slot       = 0;   // index of this context (per-context __constant__ argument blocks)
  // pinned staging buffers of upload(): allocated at the first large upload, kept until shutdown
  static constexpr int kStageBufs = 4;
  void*        h_stage[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t  stage_ev[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};
  int          args_valid = 0;   // the constant block of this slot holds the last PcgArgs this thread copied
};

// ONE CONTEXT PER HOST THREAD (thread_local): its own stream, reduction workspace and result
// mailbox.  A solver must be used from the thread that created it; solvers of different threads run
// concurrently on the device (batches of small independent QPs: BASELINE configs[4]).
Context& ctx();

inline bool check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    Context& c = ctx();
    if (!c.last_error) c.last_error = (int)e;
    fprintf(stderr, "[osqp_b200] CUDA error %d (%s) in %s\n", (int)e, cudaGetErrorString(e), what);
    return false;
  }
  return true;
}

#define B200_CHECK(expr) ::b200::check((expr), #expr)

// Stream-ordered allocation from the device's default memory pool (release threshold raised to
// "never" in b200_init): setting a solver up and tearing it down again recycles the same physical
// memory instead of going through cudaMalloc/cudaFree (each a device synchronisation and, for
// the 100 MB+ matrix arrays, milliseconds of page mapping -- measured as 100-1000 ms outliers in
// the end-to-end setup time).
template <class P>
inline cudaError_t dev_malloc(P** p, size_t bytes) {
  return cudaMallocAsync((void**)p, bytes ? bytes : 8, ctx().stream);
}
inline void dev_free(void* p) {
  if (!p) return;
  // a thread without a live library context (e.g. a garbage collector finalising a solver that another
  // thread created) has no stream to order the release on: cudaFree synchronises instead
  if (ctx().stream) cudaFreeAsync(p, ctx().stream);
  else cudaFree(p);
}

// launch tracer (context.cu): development aid, active only when B200_TRACE_FILE is set
void trace_point(const char* tag);

// called after every kernel launch: counts it and records (sticky) launch-configuration errors
inline void count_launch(const char* tag = __builtin_FUNCTION()) {
  ctx().launches++;
  ctx().epoch++;
  if (ctx().trace_on) trace_point(tag);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) check(cudaGetLastError(), "kernel launch");
}

// grid for a grid-stride elementwise kernel: enough CTAs to fill the machine, 4 elements
// per thread per trip, capped at 8 resident CTAs per SM.
inline int ew_grid(long long n) {
  long long want = (n + (long long)kBlock * 4 - 1) / ((long long)kBlock * 4);
  long long cap  = (long long)ctx().sm_count * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// context.cu: host -> device copy of a (pageable) user array on the library stream.  Large arrays
// go through pinned staging buffers filled by several host threads (a pageable cudaMemcpyAsync is
// staged by the driver on one thread at ~10 GB/s: 14 ms for the 137 MB matrix of the Lasso workload);
// on return the source may be reused, as with a pageable cudaMemcpyAsync.
bool upload(void* d_dst, const void* h_src, size_t bytes);

// context.cu: wait until the device has posted sequence number `seq` in the mailbox.  Returns
// false (after synchronising the stream) if the stream finished or failed without posting it.
bool mail_wait(unsigned long long seq);

// ---- peer-memory exchange of the row-sharded CG loop (dist.cu, pcg_graph.cu)
// Every rank owns one exchange buffer; rank q's contribution to rank r lands in slot q of r's buffer.
// Two sets of slots alternate with the parity of the exchange's sequence number: a rank can run at
// most one exchange ahead of its slowest peer (every exchange waits for all peers), so the set being
// written is never the one still being read.
constexpr int    kXchgMaxWorld     = 8;
constexpr int    kXchgCap          = 1 << 17;                                  // doubles per vector slot
constexpr size_t kXchgFlagOff      = (size_t)2 * kXchgMaxWorld * kXchgCap;      // 2 x 8 vector sequence words
constexpr int    kXchgScSlot       = 32;                                        // doubles per scalar slot
constexpr size_t kXchgScOff        = kXchgFlagOff + 2 * kXchgMaxWorld;          // 2 x 8 x 32 scalar payload
constexpr size_t kXchgScFlagOff    = kXchgScOff + 2 * kXchgMaxWorld * kXchgScSlot;   // 2 x 8 scalar sequence words
constexpr size_t kXchgTotalDoubles = kXchgScFlagOff + 2 * kXchgMaxWorld + 16;
struct XchgState {            // device memory behind the buffer; identical on all ranks by construction
  unsigned long long vseq, sseq;   // vector / scalar exchanges completed so far
  unsigned ticket, ticket2;
  int err, pad;                    // 1: a wait ran into its time limit (a peer never arrived)
};
struct XchgView {
  double*    mine;
  double*    peer[kXchgMaxWorld];  // peer[r]: rank r's buffer mapped into this process (peer[rank] == mine)
  XchgState* state;
  int        world, rank;
};
__host__ __device__ inline size_t xchg_vec(int set, int r) { return ((size_t)set * kXchgMaxWorld + r) * kXchgCap; }
__host__ __device__ inline size_t xchg_vflag(int set, int r) { return kXchgFlagOff + (size_t)set * kXchgMaxWorld + r; }
__host__ __device__ inline size_t xchg_sc(int set, int r) { return kXchgScOff + ((size_t)set * kXchgMaxWorld + r) * kXchgScSlot; }
__host__ __device__ inline size_t xchg_scflag(int set, int r) { return kXchgScFlagOff + (size_t)set * kXchgMaxWorld + r; }
bool dist_p2p_ready();
const XchgView& dist_xchg_view();
// dist.cu: combine `count` (<= kXchgScSlot) device scalars over the ranks through peer memory: entry k is
// summed (bit k of max_mask clear) or maximised (set) over the ranks in rank order when bit k of
// active_mask is set, left alone otherwise; the result goes back to d_vals and, when d_mail != nullptr, to
// the mapped host mailbox with sequence number `seq`.  false: peer path unavailable (caller uses NCCL).
bool dist_p2p_small(double* d_vals, int count, unsigned max_mask, unsigned active_mask, double* d_mail,
                    unsigned long long seq);

// dist.cu
bool dist_active();
bool dist_scope();
bool dist_split();
int  dist_n_shared();
int  dist_col_off();
void dist_allreduce_f64(double* d_buf, int n, bool is_max);

// ------------------------------------------------------------------ device side
// post `cnt` results + the sequence number into the mapped host mailbox (one thread)
__device__ __forceinline__ void mail_post(double* mail, const double* vals, int cnt, unsigned long long seq) {
  for (int i = 0; i < cnt; i++) mail[i] = vals[i];
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(mail + kMailSlots) = seq;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions; result valid in every thread.  `sh` needs >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    t = warp_max(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

}  // namespace b200
