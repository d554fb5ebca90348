// context.cu -- library lifecycle and memory entry points of the C-ABI.
// Roles taken over from the reference: osqp_algebra_init_libs/free_libs
// (algebra/cuda/algebra_libs.cu:31-75, src/cuda_handler.cu:22-44) and the raw memory helpers
// (algebra/cuda/src/cuda_memory.cu:22-112).
#include "common.cuh"

#include <cstring>
#include <cstdlib>
#include <ctime>
#include <vector>
#include <atomic>
#include <thread>
#include <nvtx3/nvToolsExt.h>   // header-only: resolves the tools injection library at first use, a no-op otherwise

namespace b200 {
Context& ctx() {
  static thread_local Context c;
  return c;
}

namespace {
static_assert(kMaxContexts <= 64, "slot mask is one 64-bit word");
std::atomic<unsigned long long> g_slot_mask{0};
int slot_acquire() {
  for (;;) {
    unsigned long long cur = g_slot_mask.load();
    int free_bit = -1;
    for (int i = 0; i < kMaxContexts; i++)
      if (!(cur & (1ull << i))) { free_bit = i; break; }
    if (free_bit < 0) return -1;
    if (g_slot_mask.compare_exchange_weak(cur, cur | (1ull << free_bit))) return free_bit;
  }
}
void slot_release(int i) { g_slot_mask.fetch_and(~(1ull << i)); }
}  // namespace

// ---- launch tracer: after every launch an event is recorded on the library stream together
// with the host time of the call.  The dump lists, per launch, the device time since the previous
// launch finished (= idle gap + kernel duration) and the host time since the previous call, which
// is what separates "the GPU was busy" from "the GPU waited for the host".
namespace {
struct TracePoint { cudaEvent_t ev; const char* tag; double host_us; };
thread_local std::vector<TracePoint> g_trace;
double host_now_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
}  // namespace
namespace {
constexpr size_t kStageChunk   = 8u << 20;    // bytes per pinned staging buffer
constexpr size_t kStageMinSize = 16u << 20;   // smaller uploads: plain (driver-staged) copy
}  // namespace

bool upload(void* d_dst, const void* h_src, size_t bytes) {
  Context& c = ctx();
  if (bytes == 0) return true;
  static const bool off = getenv("B200_NO_STAGED_UPLOAD") != nullptr;
  if (bytes < kStageMinSize || off)
    return check(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c.stream), "upload");
  for (int b = 0; b < Context::kStageBufs; b++) {
    if (!c.h_stage[b]) {
      if (cudaHostAlloc(&c.h_stage[b], kStageChunk, cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&c.stage_ev[b], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        c.h_stage[b] = nullptr;
        return check(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c.stream), "upload");
      }
    }
  }
  // worker b owns buffer b and the chunks b, b + kStageBufs, ...: wait until the previous copy out
  // of its buffer is done, refill it, enqueue the next copy (the stream serialises the copies; their
  // destinations are disjoint, so their order does not matter)
  const size_t nchunks = (bytes + kStageChunk - 1) / kStageChunk;
  std::atomic<int> failed{0};
  const int device = c.device;
  cudaStream_t st = c.stream;
  void* const* stage = c.h_stage;
  cudaEvent_t const* ev = c.stage_ev;
  auto work = [&](int b) {
    if (cudaSetDevice(device) != cudaSuccess) { failed = 1; return; }
    bool used = false;
    for (size_t ch = (size_t)b; ch < nchunks; ch += Context::kStageBufs) {
      const size_t o = ch * kStageChunk, len = (bytes - o < kStageChunk) ? bytes - o : kStageChunk;
      if (used && cudaEventSynchronize(ev[b]) != cudaSuccess) { failed = 1; return; }
      memcpy(stage[b], (const char*)h_src + o, len);
      if (cudaMemcpyAsync((char*)d_dst + o, stage[b], len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaEventRecord(ev[b], st) != cudaSuccess) { failed = 1; return; }
      used = true;
    }
  };
  std::thread th[Context::kStageBufs];
  for (int b = 1; b < Context::kStageBufs; b++) th[b] = std::thread(work, b);
  work(0);
  for (int b = 1; b < Context::kStageBufs; b++) th[b].join();
  // the buffers are reused by the next upload: make sure the last copies have left them
  for (int b = 0; b < Context::kStageBufs; b++)
    if ((size_t)b < nchunks && cudaEventSynchronize(ev[b]) != cudaSuccess) failed = 1;
  if (failed) return check(cudaErrorUnknown, "staged upload");
  return true;
}

bool mail_wait(unsigned long long seq) {
  Context& c = ctx();
  volatile unsigned long long* word = reinterpret_cast<volatile unsigned long long*>(c.h_mail + kMailSlots);
  for (unsigned spins = 0;; spins++) {
    if (*word == seq) return true;
    if ((spins & 0xfff) == 0xfff) {
      // the stream draining without the word arriving means the kernel failed
      cudaError_t q = cudaStreamQuery(c.stream);
      if (q != cudaErrorNotReady) {
        if (*word == seq) return true;
        check(q == cudaSuccess ? cudaErrorUnknown : q, "reduction mailbox");
        cudaStreamSynchronize(c.stream);
        return false;
      }
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

void trace_point(const char* tag) {
  if (g_trace.size() >= (1u << 17)) return;
  TracePoint t;
  if (cudaEventCreate(&t.ev) != cudaSuccess) return;
  cudaEventRecord(t.ev, ctx().stream);
  t.tag = tag;
  t.host_us = host_now_us();
  g_trace.push_back(t);
}
static void trace_dump() {
  const char* base = getenv("B200_TRACE_FILE");
  if (!base || g_trace.empty()) return;
  char path[1024];
  if (ctx().slot > 0) snprintf(path, sizeof(path), "%s.ctx%d", base, ctx().slot);
  else snprintf(path, sizeof(path), "%s", base);
  cudaStreamSynchronize(ctx().stream);
  FILE* f = fopen(path, "w");
  if (f) {
    fprintf(f, "idx,tag,dev_us_since_prev,host_us_since_prev\n");
    for (size_t i = 1; i < g_trace.size(); i++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_trace[i - 1].ev, g_trace[i].ev);
      fprintf(f, "%zu,%s,%.2f,%.2f\n", i, g_trace[i].tag, ms * 1e3, g_trace[i].host_us - g_trace[i - 1].host_us);
    }
    fclose(f);
  }
  for (auto& t : g_trace) cudaEventDestroy(t.ev);
  g_trace.clear();
}
}  // namespace b200

using namespace b200;

void b200_csr_configure_kernels();
void b200_pcg_configure_kernels();

extern "C" {

int b200_init(int device) {
  Context& c = ctx();
  if (c.refcount > 0) {
    c.refcount++;
    return 0;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    fprintf(stderr, "[osqp_b200] no CUDA device visible: the B200 backend has no CPU fallback\n");
    return 1;
  }
  if (device < 0 || device >= count) device = 0;
  if (!B200_CHECK(cudaSetDevice(device))) return 1;
  c.slot = slot_acquire();
  if (c.slot < 0) {
    fprintf(stderr, "[osqp_b200] more than %d library contexts (host threads with live solvers)\n", kMaxContexts);
    return 1;
  }
  // every failure from here on gives the slot (and whatever was created) back
  auto fail = [&]() {
    if (c.stream) cudaStreamSynchronize(c.stream);
    dev_free(c.d_partials); dev_free(c.d_ticket); dev_free(c.d_scalar);
    if (c.h_scalar) cudaFreeHost(c.h_scalar);
    if (c.h_mail) cudaFreeHost(c.h_mail);
    if (c.stream) cudaStreamDestroy(c.stream);
    c.d_partials = nullptr; c.d_ticket = nullptr; c.d_scalar = nullptr; c.h_scalar = nullptr;
    c.h_mail = c.d_mail = nullptr; c.stream = nullptr;
    slot_release(c.slot);
    return 1;
  };
  cudaDeviceProp prop;
  if (!B200_CHECK(cudaGetDeviceProperties(&prop, device))) return fail();
  c.device   = device;
  c.sm_count = prop.multiProcessorCount;
  snprintf(c.name, sizeof(c.name), "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor,
           prop.multiProcessorCount);
  if (!B200_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking))) { c.stream = nullptr; return fail(); }
  {
    cudaMemPool_t pool;
    if (B200_CHECK(cudaDeviceGetDefaultMemPool(&pool, device))) {
      unsigned long long keep = ~0ull;   // never hand memory back to the driver between solvers
      B200_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
  }
  bool ok = true;
  ok &= B200_CHECK(dev_malloc(&c.d_partials, sizeof(double) * kMaxRedBlocks * 32));
  ok &= B200_CHECK(dev_malloc(&c.d_ticket, sizeof(unsigned) * 4));
  ok &= B200_CHECK(dev_malloc(&c.d_scalar, sizeof(double) * kScalarSlots));
  ok &= B200_CHECK(cudaMallocHost(&c.h_scalar, sizeof(double) * kScalarSlots));
  ok &= B200_CHECK(cudaHostAlloc(&c.h_mail, sizeof(double) * (kMailSlots + 1), cudaHostAllocMapped));
  if (!ok) return fail();
  memset(c.h_mail, 0, sizeof(double) * (kMailSlots + 1));
  ok &= B200_CHECK(cudaHostGetDevicePointer(&c.d_mail, c.h_mail, 0));
  c.mail_seq = 0;
  if (!ok) return fail();
  B200_CHECK(cudaMemsetAsync(c.d_ticket, 0, sizeof(unsigned) * 4, c.stream));
  B200_CHECK(cudaMemsetAsync(c.d_scalar, 0, sizeof(double) * kScalarSlots, c.stream));
  b200_csr_configure_kernels();
  b200_pcg_configure_kernels();
  c.last_error = 0;
  c.launches   = 0;
  c.refcount   = 1;
  c.trace_on   = getenv("B200_TRACE_FILE") ? 1 : 0;
  c.args_valid = 0;
  return 0;
}

void b200_shutdown(void) {
  Context& c = ctx();
  if (c.refcount <= 0) return;
  if (--c.refcount > 0) return;
  cudaStreamSynchronize(c.stream);
  if (c.trace_on) trace_dump();
  dev_free(c.d_partials);
  dev_free(c.d_ticket);
  dev_free(c.d_scalar);
  cudaFreeHost(c.h_scalar);
  cudaFreeHost(c.h_mail);
  c.h_mail = c.d_mail = nullptr;
  for (int b = 0; b < Context::kStageBufs; b++) {
    if (c.h_stage[b]) cudaFreeHost(c.h_stage[b]);
    if (c.stage_ev[b]) cudaEventDestroy(c.stage_ev[b]);
    c.h_stage[b] = nullptr;
    c.stage_ev[b] = nullptr;
  }
  cudaStreamDestroy(c.stream);
  slot_release(c.slot);
  c.d_partials = nullptr;
  c.d_ticket   = nullptr;
  c.d_scalar   = nullptr;
  c.h_scalar   = nullptr;
  c.stream     = nullptr;
}

int b200_device_name(char* name, int len) {
  if (!name || len <= 0) return 0;
  return snprintf(name, (size_t)len, "%s", ctx().refcount > 0 ? ctx().name : "");
}

int b200_sm_count(void) { return ctx().sm_count; }

void b200_sync(void) {
  if (ctx().stream) B200_CHECK(cudaStreamSynchronize(ctx().stream));
}

void* b200_stream_handle(void) { return (void*)ctx().stream; }

int b200_last_error(void) { return ctx().last_error; }

unsigned long long b200_launch_count(void) { return ctx().launches; }

unsigned long long b200_epoch(void) { return ctx().epoch; }

unsigned long long b200_graph_launch_count(void) { return ctx().graph_launches; }

void b200_trace_mark(const char* tag) {
  if (ctx().trace_on) trace_point(strdup(tag ? tag : "mark"));   // tags live until the dump
}

// NVTX ranges (Nsight Systems / Compute timelines): the sections the reference annotates through
// osqp_profiler_sec_push / pop (include/private/profilers.h:16-31; OSQP_PROFILER_SEC_LINSYS_SOLVE around
// cuda_pcg.cu:159-161).  Active with B200_NVTX=1 so that the unprofiled path pays one branch.
static bool nvtx_on() {
  static const bool on = getenv("B200_NVTX") != nullptr;
  return on;
}
void b200_range_push(const char* name) {
  if (nvtx_on()) nvtxRangePushA(name ? name : "b200");
}
void b200_range_pop(void) {
  if (nvtx_on()) nvtxRangePop();
}

void* b200_event_create(void) {
  cudaEvent_t e = nullptr;
  if (!B200_CHECK(cudaEventCreate(&e))) return nullptr;
  return (void*)e;
}
void b200_event_destroy(void* ev) {
  if (ev) cudaEventDestroy((cudaEvent_t)ev);
}
void b200_event_record(void* ev) { B200_CHECK(cudaEventRecord((cudaEvent_t)ev, ctx().stream)); }
float b200_event_elapsed_ms(void* a, void* b) {
  float ms = 0.f;
  B200_CHECK(cudaEventSynchronize((cudaEvent_t)b));
  B200_CHECK(cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return ms;
}

void* b200_malloc(size_t bytes) {
  void* p = nullptr;
  ctx().epoch++;   // addresses may be reused: cached scalars keyed by pointer go stale
  if (bytes == 0) bytes = 8;   // keep zero-length vectors addressable
  if (!B200_CHECK(dev_malloc(&p, bytes))) return nullptr;
  return p;
}

void* b200_calloc(size_t bytes) {
  void* p = b200_malloc(bytes);
  if (p) B200_CHECK(cudaMemsetAsync(p, 0, bytes ? bytes : 8, ctx().stream));
  return p;
}

void b200_free(void* d_ptr) {
  if (!d_ptr) return;
  ctx().epoch++;
  // stream-ordered: kernels already enqueued on the library stream finish before reuse
  dev_free(d_ptr);
}

int b200_ptr_is_device(const void* ptr) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int b200_copy_in(void* d_dst, const void* src, size_t bytes) {
  if (bytes == 0) return 0;
  Context& c = ctx();
  c.epoch++;
  if (b200_ptr_is_device(src)) {
    return B200_CHECK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyDeviceToDevice, c.stream)) ? 0 : 1;
  }
  // pageable host memory: cudaMemcpyAsync stages it before returning, so the caller may
  // reuse `src` immediately; ordering with earlier kernels is kept by the stream
  return upload(d_dst, src, bytes) ? 0 : 1;
}

int b200_copy_out(void* dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return 0;
  Context& c = ctx();
  if (b200_ptr_is_device(dst)) {
    return B200_CHECK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToDevice, c.stream)) ? 0 : 1;
  }
  bool ok = B200_CHECK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, c.stream));
  ok &= B200_CHECK(cudaStreamSynchronize(c.stream));
  return ok ? 0 : 1;
}

}  // extern "C"
