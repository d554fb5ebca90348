// dist.cu -- row-sharded multi-GPU mode: one process per GPU, NCCL over NVLink/NVSwitch for the one
// exchange step the path has (the length-n partial K p, and A'y in the residual check).
//
// The reference has no multi-GPU path at all; this is new functionality specified
// by BASELINE.json north_star / SURVEY.md section 8e.  NCCL is dlopen'ed so that the single-GPU
// library has no dependency on it; in a torch process the already-loaded libnccl.so.2 is reused.
#include "common.cuh"
#include "xchg.cuh"

#include <dlfcn.h>
#include <cstring>
#include <cstdlib>

using namespace b200;

namespace {

typedef void* nccl_comm_t;
struct nccl_uid { char internal[128]; };
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm_t*, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char* (*fn_errstr)(int);

struct Dist {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
  nccl_comm_t comm = nullptr;
  int rank = 0, world = 1, scope = 0;
  int suspended = 0;     // b200_dist_suspend: solvers created meanwhile are plain single-GPU solvers
  int n_shared = -1;     // >= 0: column-split layout, n-vectors are [shared (replicated) ; local (owned)]
  unsigned long long n_allreduce = 0, bytes_allreduce = 0;
};
Dist g;

constexpr int NCCL_FLOAT = 7, NCCL_DOUBLE = 8;   // ncclFloat32 / ncclFloat64 (nccl.h ncclDataType_t)
constexpr int NCCL_SUM = 0, NCCL_MAX = 2;        // ncclRedOp_t

bool load_nccl() {
  if (g.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g.lib) break;
  }
  if (!g.lib) {
    fprintf(stderr, "[osqp_b200] cannot dlopen libnccl.so.2: %s\n", dlerror());
    return false;
  }
  g.get_uid   = (fn_get_uid)dlsym(g.lib, "ncclGetUniqueId");
  g.init_rank = (fn_init_rank)dlsym(g.lib, "ncclCommInitRank");
  g.allreduce = (fn_allreduce)dlsym(g.lib, "ncclAllReduce");
  g.destroy   = (fn_destroy)dlsym(g.lib, "ncclCommDestroy");
  g.errstr    = (fn_errstr)dlsym(g.lib, "ncclGetErrorString");
  return g.get_uid && g.init_rank && g.allreduce && g.destroy;
}

bool nccl_ok(int rc, const char* what) {
  if (rc == 0) return true;
  fprintf(stderr, "[osqp_b200] NCCL error %d (%s) in %s\n", rc, g.errstr ? g.errstr(rc) : "?", what);
  if (!ctx().last_error) ctx().last_error = 10000 + rc;
  return false;
}

}  // namespace

namespace b200 {
// used by the reductions in vec_kernels.cu: combine a device scalar block across ranks when the
// backend declared the operand row-sharded
bool dist_active() { return g.comm != nullptr && g.world > 1 && !g.suspended; }
bool dist_scope() { return dist_active() && g.scope; }
// column-split layout: every n-vector is [shared columns (replicated on all ranks) ; local columns
// (owned by this rank)].  A reduction over such a vector counts the shared slice on rank 0 only.
bool dist_split() { return dist_active() && g.n_shared >= 0; }
int  dist_n_shared() { return dist_split() ? g.n_shared : 0; }
int  dist_col_off() { return (dist_split() && g.rank > 0) ? g.n_shared : 0; }
void dist_allreduce_f64(double* d_buf, int n, bool is_max) {
  if (!dist_active() || n <= 0) return;
  nccl_ok(g.allreduce(d_buf, d_buf, (size_t)n, NCCL_DOUBLE, is_max ? NCCL_MAX : NCCL_SUM, g.comm,
                      ctx().stream), "ncclAllReduce(f64)");
  if (ctx().trace_on) trace_point(n <= 4 ? "allreduce(scalars)" : "allreduce(f64 block)");
  g.n_allreduce++;
  g.bytes_allreduce += (unsigned long long)n * 8;
}
}  // namespace b200

// ------------------------------------------------------------------ peer-memory exchange (NVLink P2P)
// The CG loop of the row-sharded solve exchanges (a) the n_shared-long head of the partial K p plus three
// dot-product partials and (b) two scalars per iteration.  Through NCCL that is 3 collectives of 14-19 us
// each plus a host read-back (profiles/r01_sharded_scaling.md); here every rank PUSHES its contribution
// straight into a slot of every peer's exchange buffer (plain stores to cudaIpc-mapped peer memory,
// __threadfence_system, release-store of a sequence number) from inside the producing kernels, and folds
// what it received in rank order -- bit-identical on all ranks, no host involvement, so the loop can be a
// CUDA-graph WHILE node exactly as on one GPU (pcg_graph.cu).  The buffer is cudaMalloc'ed (IPC handles
// cannot be taken from the stream-ordered pool) and the 64-byte handles travel through the host program.
namespace {
struct P2P {
  double* mine = nullptr;
  double* peer[kXchgMaxWorld] = {nullptr};
  XchgState* state = nullptr;
  int ready = 0;
};
P2P p2p;
}  // namespace

namespace {
// ONE CTA: `count` scalars of this rank go into slot `rank` of every peer, the peers' come back, entry k is
// folded over the ranks in rank order (sum / max), the result is written back and posted to the host
// mailbox.  Shares the scalar slots and their sequence counter with the CG loop's scalar exchange
// (pcg_graph.cu): every rank issues the same exchanges in the same stream order.
__global__ void __launch_bounds__(64) g_xchg_small(XchgView X, double* vals, int count, unsigned max_mask,
                                                   unsigned active_mask, double* mail, unsigned long long mseq) {
  __shared__ double mine[kXchgScSlot], res[kXchgScSlot];
  XchgState* S = X.state;
  const int tid = threadIdx.x, me = X.rank, world = X.world;
  if (tid < count) mine[tid] = vals[tid];
  __syncthreads();
  const unsigned long long seq = *(volatile unsigned long long*)&S->sseq + 1;
  const int set = (int)(seq & 1ull);
  if (tid < world && tid != me) {
    double* dst = X.peer[tid] + xchg_sc(set, me);
    for (int k = 0; k < count; k++) dst[k] = mine[k];
    __threadfence_system();
    st_release_sys((unsigned long long*)(X.peer[tid] + xchg_scflag(set, me)), seq);
    if (!xchg_wait(S, (const unsigned long long*)(X.mine + xchg_scflag(set, tid)), seq)) S->err = 1;
  }
  __syncthreads();
  if (tid < count) {
    double a = mine[tid];
    if ((active_mask >> tid) & 1u) {
      const bool is_max = (max_mask >> tid) & 1u;
      a = 0.0;
      for (int r = 0; r < world; r++) {
        const double v = (r == me) ? mine[tid] : __ldcg(X.mine + xchg_sc(set, r) + tid);
        a = is_max ? fmax(a, v) : a + v;
      }
    }
    vals[tid] = a;
    res[tid] = a;
  }
  __syncthreads();
  if (tid == 0) {
    S->sseq = seq;
    if (mail) mail_post(mail, res, count, mseq);
  }
}

// the first `n` entries of `buf` summed (or maximised) over the ranks, in place, folded in rank order; same
// protocol as g_xchg_vector of the CG loop (vector slots, vseq).  The grid must be co-resident.
template <bool IS_MAX>
__global__ void __launch_bounds__(kBlock) g_xchg_plain(XchgView X, T* buf, int n) {
  __shared__ int s_last;
  XchgState* S = X.state;
  const int me = X.rank, world = X.world;
  const unsigned long long seq = *(volatile unsigned long long*)&S->vseq + 1;
  const int set = (int)(seq & 1ull);
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  for (int q = 0; q < world; q++) {
    if (q == me) continue;
    double* dst = X.peer[q] + xchg_vec(set, me);
    for (int i = gtid; i < n; i += gstride) dst[i] = (double)buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&S->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if ((int)threadIdx.x < world && (int)threadIdx.x != me)
      st_release_sys((unsigned long long*)(X.peer[threadIdx.x] + xchg_vflag(set, me)), seq);
    if (threadIdx.x == 0) S->ticket = 0;
  }
  if ((int)threadIdx.x < world && (int)threadIdx.x != me) {
    if (!xchg_wait(S, (const unsigned long long*)(X.mine + xchg_vflag(set, threadIdx.x)), seq)) S->err = 1;
  }
  __syncthreads();
  for (int i = gtid; i < n; i += gstride) {
    double a = 0.0;
    for (int r = 0; r < world; r++) {
      const double v = (r == me) ? (double)buf[i] : __ldcg(X.mine + xchg_vec(set, r) + i);
      a = IS_MAX ? fmax(a, v) : a + v;
    }
    buf[i] = (T)a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&S->ticket2, 1u) == gridDim.x - 1) {   // every CTA has read vseq and folded its part
      S->ticket2 = 0;
      S->vseq = seq;
    }
  }
}

// peer path for a vector head; false when unavailable
template <bool IS_MAX>
bool p2p_vector(T* d_buf, int n);
}  // namespace

namespace b200 {
XchgView g_xchg_view;
bool dist_p2p_ready() { return p2p.ready && dist_active() && g.world <= kXchgMaxWorld; }
static bool p2p_outside_loop() {
  static const bool off = getenv("B200_DIST_NO_P2P") != nullptr || getenv("B200_DIST_P2P_LOOP_ONLY") != nullptr;
  return dist_p2p_ready() && !off;
}
bool dist_p2p_small(double* d_vals, int count, unsigned max_mask, unsigned active_mask, double* d_mail,
                    unsigned long long seq) {
  if (!p2p_outside_loop() || count <= 0 || count > kXchgScSlot) return false;
  g_xchg_small<<<1, 64, 0, ctx().stream>>>(g_xchg_view, d_vals, count, max_mask, active_mask, d_mail, seq);
  count_launch("xchg(small)");
  return true;
}
const XchgView& dist_xchg_view() { return g_xchg_view; }
}  // namespace b200

namespace {
template <bool IS_MAX>
bool p2p_vector(T* d_buf, int n) {
  if (!p2p_outside_loop() || n + 8 > kXchgCap) return false;
  int grid = (n + kBlock * 2 - 1) / (kBlock * 2);
  if (grid > ctx().sm_count) grid = ctx().sm_count;
  if (grid < 1) grid = 1;
  g_xchg_plain<IS_MAX><<<grid, kBlock, 0, ctx().stream>>>(g_xchg_view, d_buf, n);
  count_launch("xchg(plain vector)");
  return true;
}
}  // namespace

static void p2p_release() {
  if (!p2p.mine) return;
  cudaDeviceSynchronize();
  for (int r = 0; r < kXchgMaxWorld; r++)
    if (p2p.peer[r] && p2p.peer[r] != p2p.mine) cudaIpcCloseMemHandle(p2p.peer[r]);
  cudaFree(p2p.mine);
  p2p = P2P();
}
static void p2p_release_hook() { p2p_release(); }

extern "C" {

int b200_dist_unique_id(unsigned char* id128) {
  if (!load_nccl()) return 1;
  nccl_uid u;
  if (!nccl_ok(g.get_uid(&u), "ncclGetUniqueId")) return 1;
  memcpy(id128, u.internal, 128);
  return 0;
}

int b200_dist_init(int rank, int world, const unsigned char* id128) {
  if (world <= 1) { g.rank = 0; g.world = 1; return 0; }
  if (!load_nccl()) return 1;
  if (g.comm) return 0;
  nccl_uid u;
  memcpy(u.internal, id128, 128);
  if (!nccl_ok(g.init_rank(&g.comm, world, u, rank), "ncclCommInitRank")) return 1;
  g.rank = rank;
  g.world = world;
  return 0;
}

void b200_dist_finalize(void) {
  p2p_release_hook();
  if (g.comm) {
    cudaStreamSynchronize(ctx().stream);
    g.destroy(g.comm);
    g.comm = nullptr;
  }
  g.world = 1;
  g.rank = 0;
}

int b200_dist_world(void) { return g.suspended ? 1 : g.world; }
int b200_dist_rank(void) { return g.suspended ? 0 : g.rank; }
void b200_dist_suspend(int on) { g.suspended = on ? 1 : 0; }
void b200_dist_scope(int sharded) { g.scope = sharded; }
void b200_dist_set_split(int n_shared) { g.n_shared = n_shared; }
int  b200_dist_n_shared(void) { return g.n_shared; }

void b200_dist_allreduce_sum(T* d_buf, int n) {
  if (!dist_active() || n <= 0) return;
  ctx().epoch++;
  if (p2p_vector<false>(d_buf, n)) return;
  nccl_ok(g.allreduce(d_buf, d_buf, (size_t)n, sizeof(T) == 8 ? NCCL_DOUBLE : NCCL_FLOAT, NCCL_SUM, g.comm,
                      ctx().stream), "ncclAllReduce(sum)");
  if (ctx().trace_on) trace_point("allreduce(vector sum)");
  g.n_allreduce++;
  g.bytes_allreduce += (unsigned long long)n * sizeof(T);
}

void b200_dist_allreduce_max(T* d_buf, int n) {
  if (!dist_active() || n <= 0) return;
  ctx().epoch++;
  if (p2p_vector<true>(d_buf, n)) return;
  nccl_ok(g.allreduce(d_buf, d_buf, (size_t)n, sizeof(T) == 8 ? NCCL_DOUBLE : NCCL_FLOAT, NCCL_MAX, g.comm,
                      ctx().stream), "ncclAllReduce(max)");
  if (ctx().trace_on) trace_point("allreduce(vector max)");
  g.n_allreduce++;
  g.bytes_allreduce += (unsigned long long)n * sizeof(T);
}

int b200_dist_p2p_export(unsigned char* handle64) {
  if (!p2p.mine) {
    const size_t bytes = sizeof(double) * kXchgTotalDoubles + sizeof(XchgState);
    if (!B200_CHECK(cudaMalloc((void**)&p2p.mine, bytes))) return 1;
    if (!B200_CHECK(cudaMemset(p2p.mine, 0, bytes))) return 1;
    if (!B200_CHECK(cudaDeviceSynchronize())) return 1;
  }
  cudaIpcMemHandle_t h;
  if (!B200_CHECK(cudaIpcGetMemHandle(&h, p2p.mine))) return 1;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  return 0;
}

int b200_dist_p2p_import(const unsigned char* handles, int world) {
  if (world != g.world || world > kXchgMaxWorld || !p2p.mine) return 1;
  for (int r = 0; r < world; r++) {
    if (r == g.rank) { p2p.peer[r] = p2p.mine; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    if (!B200_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess))) return 1;
    p2p.peer[r] = (double*)ptr;
  }
  XchgView v;
  memset(&v, 0, sizeof(v));
  v.mine = p2p.mine;
  for (int r = 0; r < world; r++) v.peer[r] = p2p.peer[r];
  v.state = (XchgState*)(p2p.mine + kXchgTotalDoubles);
  v.world = world;
  v.rank = g.rank;
  g_xchg_view = v;
  p2p.ready = 1;
  return 0;
}

int b200_dist_p2p_enabled(void) { return dist_p2p_ready() ? 1 : 0; }

// 1 if a wait of the peer-memory exchange ever ran into its time limit (a peer never arrived): the results
// since then are meaningless.  Synchronises the library stream.
int b200_dist_p2p_error(void) {
  if (!p2p.mine) return 0;
  XchgState h;
  memset(&h, 0, sizeof(h));
  cudaStreamSynchronize(ctx().stream);
  if (cudaMemcpy(&h, p2p.mine + kXchgTotalDoubles, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return 2;
  if (h.err && !ctx().last_error) ctx().last_error = 20000;
  return h.err;
}


void b200_dist_stats(unsigned long long* n_calls, unsigned long long* bytes) {
  if (n_calls) *n_calls = g.n_allreduce;
  if (bytes) *bytes = g.bytes_allreduce;
}

}  // extern "C"
