// b200_kernels.cu -- hand-written sm_100a kernels + the C-ABI of include/b200_sts.h.
//
// Everything on this path is FP64 and HBM-bound (~0.5 flop/byte): no tensor cores.
// What matters is (i) one pass over memory per STS stage, (ii) 16-byte coalesced
// loads with enough of them in flight, (iii) re-using each input row for its three
// vertical uses from registers, (iv) deterministic reductions, and (v) bit-faithful
// arithmetic: explicit __dmul_rn/__dadd_rn in the reference's association order
// so no FMA contraction can change a result.
//
// Reference behaviour restated (paths relative to /root/reference):
//   stencil            diffusion_2D/diffusion.cpp:34-55 (+ faces :68-205)
//   linear combination deps/sundials/src/sundials/sundials_nvector.c:557-565
//   vector ops         deps/sundials/src/nvector/parallel/nvector_parallel.c:424-730
//   halo pack          diffusion_2D/buffers.cpp:20-43
//   Jacobi setup       diffusion_2D/preconditioner_jacobi.cpp:9-46
//   adr callbacks      adr/advection_diffusion_reaction_2d.cpp:1406-1520

#ifdef B200_HOST_EMU
// tests/emu only: the same translation unit compiled with g++ against a host emulation of the CUDA runtime and
// execution model, so that the host stack above the C-ABI can be tested without a GPU.  Never part of the product.
#include "cuda_runtime_emu.h"
#else
#include <cuda_runtime.h>
#include <nccl.h> // types only: the library is bound with dlopen (see nccl_api below)
#endif
#include <dlfcn.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#ifdef B200_HOST_EMU
#pragma GCC visibility push(default) // the emulated build hides everything but the C-ABI
#endif
#include "b200_sts.h"
#ifdef B200_HOST_EMU
#pragma GCC visibility pop
#endif
#include "kernel_prims.cuh"

// every kernel launch of the library goes through here
// B200_TRACE_LAUNCHES=1: every launch is bracketed by stream synchronisations and its wall time (launch latency +
// execution) is accumulated per kernel, as is the host time that passes BETWEEN launches; b200_trace_report() prints
// the table (also called when a context is destroyed).  A diagnostic: it serialises host and device.
struct LaunchTrace
{
  struct Row { const char* name; const void* key; uint64_t n; double ms; unsigned gx, gy, block; size_t smem; };
  std::vector<Row> rows;
  double gap_ms = 0.0, last_end = 0.0;
  int on = -1;
};
static LaunchTrace g_trace;
static double trace_now_ms()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}
static bool trace_on()
{
  if (g_trace.on < 0) g_trace.on = getenv("B200_TRACE_LAUNCHES") ? 1 : 0;
  return g_trace.on == 1;
}
#if defined(__x86_64__) || defined(__i386__)
#define B200_CPU_RELAX() __builtin_ia32_pause()
#else
#define B200_CPU_RELAX() do { } while (0)
#endif
// B200_HOST_PROFILE=1: host time spent inside kernel launches and inside stream synchronisations (nothing is
// serialised, two clock reads per call); reported with the trace report.  Tells launch-bound small-grid runs apart
// from sync-bound ones.
struct HostProfile
{
  int on = -1;
  uint64_t launches = 0, syncs = 0;
  double launch_ms = 0.0, sync_ms = 0.0, launch_max_ms = 0.0; // (the slowest launch call is the module load: reported apart)
};
static HostProfile g_hprof;
static bool hprof_on()
{
  if (g_hprof.on < 0) g_hprof.on = getenv("B200_HOST_PROFILE") ? 1 : 0;
  return g_hprof.on == 1;
}
static inline cudaError_t stream_sync_profiled(cudaStream_t st)
{
  if (!hprof_on()) return cudaStreamSynchronize(st);
  const double t0 = trace_now_ms();
  cudaError_t e   = cudaStreamSynchronize(st);
  g_hprof.sync_ms += trace_now_ms() - t0;
  g_hprof.syncs++;
  return e;
}
extern "C" int b200_host_profile_on(void) { return hprof_on() ? 1 : 0; }
extern "C" void b200_host_profile_wait(double ms) // a wait for a result that happened outside this library (the vector's slot poll)
{
  g_hprof.sync_ms += ms;
  g_hprof.syncs++;
}
extern "C" void b200_trace_report(void)
{
  if (hprof_on() && g_hprof.launches)
  {
    const double rest = g_hprof.launch_ms - g_hprof.launch_max_ms;
    fprintf(stderr, "[b200 host profile] %llu launches: %.3f ms inside the launch calls without the slowest one (%.2f us each; slowest %.3f ms); "
                    "%llu waits for a result (stream sync or value poll): %.3f ms (%.2f us each)\n",
            (unsigned long long)g_hprof.launches, rest, g_hprof.launches > 1 ? 1e3 * rest / (double)(g_hprof.launches - 1) : 0.0,
            g_hprof.launch_max_ms, (unsigned long long)g_hprof.syncs, g_hprof.sync_ms,
            g_hprof.syncs ? 1e3 * g_hprof.sync_ms / (double)g_hprof.syncs : 0.0);
    g_hprof.launches = g_hprof.syncs = 0;
    g_hprof.launch_ms = g_hprof.sync_ms = g_hprof.launch_max_ms = 0.0;
  }
  if (!trace_on() || g_trace.rows.empty()) return;
  double total = 0.0;
  uint64_t n   = 0;
  for (auto& r : g_trace.rows) { total += r.ms; n += r.n; }
  fprintf(stderr, "[b200 trace] %llu launches, %.3f ms in kernels (launch + execution), %.3f ms of host time between launches\n",
          (unsigned long long)n, total, g_trace.gap_ms);
  for (auto& r : g_trace.rows)
    fprintf(stderr, "[b200 trace]   %-46s grid %5u x %5u x %3u smem %6zu  n=%6llu  total %10.3f ms  avg %9.1f us\n", r.name, r.gx, r.gy,
            r.block, r.smem, (unsigned long long)r.n, r.ms, 1e3 * r.ms / (double)r.n);
  g_trace.rows.clear();
  g_trace.gap_ms = 0.0;
  g_trace.last_end = 0.0;
}

template <class... KArgs, class... Args>
static inline void klaunch_named(const char* name, void (*kern)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args)
{
#ifdef B200_HOST_EMU
  (void)st;
  if (!trace_on()) { emu::launch_body(grid, block, smem, [&]() { kern(args...); }); return; }
  const double t0 = trace_now_ms();
  emu::launch_body(grid, block, smem, [&]() { kern(args...); });
  const double t1 = trace_now_ms();
#else
  if (!trace_on())
  {
    if (!hprof_on()) { kern<<<grid, block, smem, st>>>(args...); return; }
    const double h0 = trace_now_ms();
    kern<<<grid, block, smem, st>>>(args...);
    const double hd = trace_now_ms() - h0;
    g_hprof.launch_ms += hd;
    if (hd > g_hprof.launch_max_ms) g_hprof.launch_max_ms = hd;
    g_hprof.launches++;
    return;
  }
  cudaStreamSynchronize(st);
  const double t0 = trace_now_ms();
  if (g_trace.last_end > 0.0) g_trace.gap_ms += t0 - g_trace.last_end;
  kern<<<grid, block, smem, st>>>(args...);
  cudaStreamSynchronize(st);
  const double t1 = trace_now_ms();
  g_trace.last_end = t1;
#endif
  const void* key = reinterpret_cast<const void*>(kern); // one row per instantiation (the label is the call site's spelling)
  for (auto& r : g_trace.rows)
    if (r.key == key) { r.n++; r.ms += t1 - t0; return; }
  g_trace.rows.push_back({name, key, 1, t1 - t0, grid.x, grid.y, block, smem});
}
#define klaunch(kern, ...) klaunch_named(#kern, kern, __VA_ARGS__)

// --------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_alg_bytes{0}; // modelled bytes (full-vector reads + writes) of all launches
#define ALG_BYTES(nvec_touches, ndoubles) g_alg_bytes.fetch_add((uint64_t)(nvec_touches) * 8ull * (uint64_t)(ndoubles), std::memory_order_relaxed)

#define CU_TRY(expr)                                                          \
  do {                                                                        \
    cudaError_t e_ = (expr);                                                  \
    if (e_ != cudaSuccess)                                                    \
    {                                                                         \
      snprintf(g_err, sizeof(g_err), "%s:%d: %s -> %s", __FILE__, __LINE__,   \
               #expr, cudaGetErrorString(e_));                                \
      return (int)e_ ? (int)e_ : -1;                                          \
    }                                                                         \
  }                                                                           \
  while (0)

#define NCCL_TRY(expr)                                                        \
  do {                                                                        \
    ncclResult_t r_ = (expr);                                                 \
    if (r_ != ncclSuccess)                                                    \
    {                                                                         \
      snprintf(g_err, sizeof(g_err), "%s:%d: %s -> %s", __FILE__, __LINE__,   \
               #expr, ncclGetErrorString(r_));                                \
      return 10000 + (int)r_;                                                 \
    }                                                                         \
  }                                                                           \
  while (0)

// B200_FAIL_AFTER_LAUNCHES=n (fault injection for the tests of the failure path): launch number n reports an error
static long long g_fail_after = -2;
static bool injected_failure(uint64_t launch_no)
{
  if (g_fail_after == -2)
  {
    const char* e = getenv("B200_FAIL_AFTER_LAUNCHES");
    g_fail_after  = e ? atoll(e) : -1;
  }
  return g_fail_after >= 0 && (long long)launch_no == g_fail_after;
}
#define LAUNCH_CHECK()                                                        \
  do {                                                                        \
    const uint64_t n_ = g_launches.fetch_add(1, std::memory_order_relaxed);   \
    if (injected_failure(n_ + 1)) return fail("injected failure (B200_FAIL_AFTER_LAUNCHES)"); \
    CU_TRY(cudaGetLastError());                                               \
  }                                                                           \
  while (0)

static int fail(const char* msg)
{
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}

extern "C" const char* b200_last_error(void) { return g_err; }
extern "C" uint64_t b200_launch_count(void) { return g_launches.load(); }
extern "C" uint64_t b200_algorithmic_bytes(void) { return g_alg_bytes.load(); }

// -------------------------------------------------------------------- context
static const int kMaxPartials = 1 << 18;
static const int kSmallMax    = 64; // doubles read back through mapped host memory instead of a DMA copy
static ncclResult_t (*g_nccl_destroy)(ncclComm_t) = nullptr; // set once NCCL is bound

struct b200_ctx
{
  int device          = 0;
  int sm_count        = 148;
  cudaStream_t stream = nullptr;
  bool own_stream     = false;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_compute   = nullptr;
  cudaEvent_t ev_comm      = nullptr;
  bool comm_pending        = false;
  double* partials    = nullptr; // [kMaxPartials] block partials
  unsigned* ticket    = nullptr; // last-block-done counter
  double* dev_result  = nullptr; // [8] device scalars
  double* host_result = nullptr; // [kSmallMax] pinned, MAPPED: kernels write results straight into it
  double* host_result_dev = nullptr; // device alias of host_result
  ncclComm_t comm     = nullptr;
  int rank = 0, nranks = 1;
  double* strips      = nullptr; // W/E send staging of b200_deep_halo_exchange
  size_t strip_cap    = 0;
};

extern "C" int b200_ctx_create(int device, void* stream, b200_ctx** out)
{
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) return fail("b200_ctx_create: no CUDA device (there is no CPU fallback)");
  CU_TRY(cudaSetDevice(device));
  b200_ctx* c = new b200_ctx();
  c->device   = device;
  CU_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  if (stream) { c->stream = (cudaStream_t)stream; }
  else
  {
    CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  CU_TRY(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_compute, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
  CU_TRY(cudaMalloc(&c->partials, sizeof(double) * kMaxPartials));
  CU_TRY(cudaMalloc(&c->ticket, sizeof(unsigned) * 4));
  CU_TRY(cudaMemset(c->ticket, 0, sizeof(unsigned) * 4));
  CU_TRY(cudaMalloc(&c->dev_result, sizeof(double) * 8));
  CU_TRY(cudaHostAlloc(&c->host_result, sizeof(double) * kSmallMax, cudaHostAllocMapped));
  CU_TRY(cudaHostGetDevicePointer(&c->host_result_dev, c->host_result, 0));
  *out = c;
  return 0;
}

extern "C" int b200_ctx_destroy(b200_ctx* c)
{
  if (!c) return 0;
  b200_trace_report();
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->comm_stream);
  if (c->comm && g_nccl_destroy) g_nccl_destroy(c->comm);
  if (c->strips) cudaFree(c->strips);
  cudaFree(c->partials);
  cudaFree(c->ticket);
  cudaFree(c->dev_result);
  cudaFreeHost(c->host_result);
  cudaEventDestroy(c->ev_compute);
  cudaEventDestroy(c->ev_comm);
  cudaStreamDestroy(c->comm_stream);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" void* b200_ctx_stream(b200_ctx* c) { return (void*)c->stream; }

extern "C" int b200_ctx_sync(b200_ctx* c)
{
  CU_TRY(stream_sync_profiled(c->stream));
  return 0;
}

extern "C" int b200_malloc(b200_ctx* c, int64_t n, double** dptr)
{
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMalloc((void**)dptr, sizeof(double) * (size_t)(n > 0 ? n : 1)));
  return 0;
}

extern "C" int b200_free(b200_ctx* c, double* dptr)
{
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaFree(dptr));
  return 0;
}

extern "C" int b200_host_alloc(int64_t n, double** hptr)
{
  CU_TRY(cudaMallocHost((void**)hptr, sizeof(double) * (size_t)(n > 0 ? n : 1)));
  return 0;
}

// pinned host memory that kernels can write directly (mapped): results a kernel's last block stores there are
// visible to the host after b200_ctx_sync, without a copy or a publishing launch
extern "C" int b200_mapped_alloc(b200_ctx* c, int64_t n, double** hptr, double** dptr)
{
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaHostAlloc((void**)hptr, sizeof(double) * (size_t)(n > 0 ? n : 1), cudaHostAllocMapped));
  CU_TRY(cudaHostGetDevicePointer((void**)dptr, *hptr, 0));
  return 0;
}

extern "C" int b200_host_free(double* hptr)
{
  CU_TRY(cudaFreeHost(hptr));
  return 0;
}

extern "C" int b200_h2d(b200_ctx* c, double* dst, const double* src, int64_t n)
{
  CU_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

// Scalars and other tiny read-backs (reduction results, fused WRMS slots) do not go through the copy
// engine: a 1-block kernel stores them into mapped pinned host memory and the host waits for the
// stream.  A DMA copy would queue behind whatever bulk transfer is in flight on the same engine
// (b200_pipe_*: a 2 GiB download delays every reduction of the next batch by ~40 ms).
__global__ void k_publish_small(double* __restrict__ host_mapped, const double* __restrict__ src, int n)
{
  for (int i = threadIdx.x; i < n; i += blockDim.x) host_mapped[i] = src[i];
}
static int read_small(b200_ctx* c, double* dst, const double* src, int n)
{
  klaunch(k_publish_small, 1, 32, 0, c->stream, c->host_result_dev, src, n);
  CU_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CU_TRY(cudaStreamSynchronize(c->stream));
  memcpy(dst, c->host_result, sizeof(double) * (size_t)n);
  return 0;
}

extern "C" int b200_d2h(b200_ctx* c, double* dst, const double* src, int64_t n)
{
  if (n <= kSmallMax) return read_small(c, dst, src, (int)n);
  CU_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------- pipelined host <-> device staging
// A stream of independent states (ensemble members, parameter sweeps, bench.py's end-to-end leg):
// the upload of state i+1 and the download of result i-1 run on their own copy streams while
// state i is integrated on the compute stream.  Two input and two output staging buffers in HBM;
// the hand-over to / from the integrator's vectors is a device-to-device copy on the compute
// stream, so pooled vector buffers are never touched from a copy stream.
struct b200_pipe
{
  b200_ctx* ctx;
  int64_t n;
  double* in[2];
  double* out[2];
  cudaStream_t h2d, d2h;
  cudaEvent_t in_ready[2], in_free[2], out_ready[2], out_free[2];
  // B200_PIPE_TRACE=1: timing events around every transfer, printed by b200_pipe_drain
  bool trace;
  struct Span { const char* what; int64_t seq; cudaEvent_t a, b; };
  std::vector<Span>* spans;
};

static void pipe_mark(b200_pipe* p, const char* what, int64_t seq, cudaStream_t st, bool begin)
{
  if (!p->trace) return;
  if (begin)
  {
    b200_pipe::Span sp{what, seq, nullptr, nullptr};
    cudaEventCreate(&sp.a);
    cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, st);
    p->spans->push_back(sp);
  }
  else cudaEventRecord(p->spans->back().b, st);
}

extern "C" int b200_pipe_destroy(b200_pipe* p)
{
  if (!p) return 0;
  cudaSetDevice(p->ctx->device);
  if (p->h2d) cudaStreamSynchronize(p->h2d);
  if (p->d2h) cudaStreamSynchronize(p->d2h);
  cudaStreamSynchronize(p->ctx->stream);
  for (int s = 0; s < 2; s++)
  {
    if (p->in[s]) cudaFree(p->in[s]);
    if (p->out[s]) cudaFree(p->out[s]);
    if (p->in_ready[s]) cudaEventDestroy(p->in_ready[s]);
    if (p->in_free[s]) cudaEventDestroy(p->in_free[s]);
    if (p->out_ready[s]) cudaEventDestroy(p->out_ready[s]);
    if (p->out_free[s]) cudaEventDestroy(p->out_free[s]);
  }
  if (p->h2d) cudaStreamDestroy(p->h2d);
  if (p->d2h) cudaStreamDestroy(p->d2h);
  if (p->spans)
  {
    for (auto& sp : *p->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    delete p->spans;
  }
  delete p;
  return 0;
}

static int pipe_init(b200_pipe* p)
{
  CU_TRY(cudaSetDevice(p->ctx->device));
  CU_TRY(cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking));
  for (int s = 0; s < 2; s++)
  {
    CU_TRY(cudaMalloc(&p->in[s], sizeof(double) * (size_t)p->n));
    CU_TRY(cudaMalloc(&p->out[s], sizeof(double) * (size_t)p->n));
    CU_TRY(cudaEventCreateWithFlags(&p->in_ready[s], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&p->in_free[s], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&p->out_ready[s], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&p->out_free[s], cudaEventDisableTiming));
  }
  return 0;
}

extern "C" int b200_pipe_create(b200_ctx* c, int64_t n, b200_pipe** out)
{
  if (!c || n <= 0 || !out) return fail("b200_pipe_create: bad argument");
  b200_pipe* p = new b200_pipe();
  memset(p, 0, sizeof(*p));
  p->ctx = c;
  p->n   = n;
  p->trace = getenv("B200_PIPE_TRACE") != nullptr;
  p->spans = new std::vector<b200_pipe::Span>();
  int rc = pipe_init(p);
  if (rc)
  {
    b200_pipe_destroy(p);
    return rc;
  }
  *out = p;
  return 0;
}

// A wait on an event that was never recorded is a no-op, which is what the first use of a slot needs.
extern "C" int b200_pipe_upload(b200_pipe* p, int64_t seq, const double* host_src)
{
  const int s = (int)(seq & 1);
  CU_TRY(cudaStreamWaitEvent(p->h2d, p->in_free[s], 0));
  pipe_mark(p, "h2d", seq, p->h2d, true);
  CU_TRY(cudaMemcpyAsync(p->in[s], host_src, sizeof(double) * (size_t)p->n, cudaMemcpyHostToDevice, p->h2d));
  pipe_mark(p, "h2d", seq, p->h2d, false);
  CU_TRY(cudaEventRecord(p->in_ready[s], p->h2d));
  return 0;
}

extern "C" int b200_pipe_take(b200_pipe* p, int64_t seq, double* dst_dev)
{
  const int s = (int)(seq & 1);
  CU_TRY(cudaStreamWaitEvent(p->ctx->stream, p->in_ready[s], 0));
  pipe_mark(p, "take", seq, p->ctx->stream, true);
  CU_TRY(cudaMemcpyAsync(dst_dev, p->in[s], sizeof(double) * (size_t)p->n, cudaMemcpyDeviceToDevice, p->ctx->stream));
  pipe_mark(p, "take", seq, p->ctx->stream, false);
  CU_TRY(cudaEventRecord(p->in_free[s], p->ctx->stream));
  return 0;
}

extern "C" int b200_pipe_put(b200_pipe* p, int64_t seq, const double* src_dev, double* host_dst)
{
  const int s = (int)(seq & 1);
  CU_TRY(cudaStreamWaitEvent(p->ctx->stream, p->out_free[s], 0));
  pipe_mark(p, "put", seq, p->ctx->stream, true);
  CU_TRY(cudaMemcpyAsync(p->out[s], src_dev, sizeof(double) * (size_t)p->n, cudaMemcpyDeviceToDevice, p->ctx->stream));
  pipe_mark(p, "put", seq, p->ctx->stream, false);
  CU_TRY(cudaEventRecord(p->out_ready[s], p->ctx->stream));
  CU_TRY(cudaStreamWaitEvent(p->d2h, p->out_ready[s], 0));
  pipe_mark(p, "d2h", seq, p->d2h, true);
  CU_TRY(cudaMemcpyAsync(host_dst, p->out[s], sizeof(double) * (size_t)p->n, cudaMemcpyDeviceToHost, p->d2h));
  pipe_mark(p, "d2h", seq, p->d2h, false);
  CU_TRY(cudaEventRecord(p->out_free[s], p->d2h));
  return 0;
}

extern "C" int b200_pipe_drain(b200_pipe* p)
{
  CU_TRY(cudaStreamSynchronize(p->h2d));
  CU_TRY(cudaStreamSynchronize(p->ctx->stream));
  CU_TRY(cudaStreamSynchronize(p->d2h));
  if (p->trace && !p->spans->empty())
  {
    cudaEvent_t base = p->spans->front().a;
    for (auto& sp : *p->spans)
    {
      float t0 = 0.f, t1 = 0.f;
      cudaEventElapsedTime(&t0, base, sp.a);
      cudaEventElapsedTime(&t1, base, sp.b);
      fprintf(stderr, "[b200_pipe] %-4s seq %2lld  %9.3f -> %9.3f ms  (%.3f ms)\n", sp.what, (long long)sp.seq, t0, t1, t1 - t0);
    }
    for (auto& sp : *p->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    p->spans->clear();
  }
  return 0;
}

// ------------------------------------------------------------ device helpers
#include "kernel_prims.cuh"

#include "reduce_prims.cuh"

// ------------------------------------------------------ elementwise kernels
#include "vector_kernels.cuh"

template <int OP>
static int launch_ew(b200_ctx* c, const EwArgs& a)
{
  if (a.n <= 0) return 0;
  int64_t n2     = (a.n + 1) >> 1;
  int64_t blocks = (n2 + 2 * kThreads - 1) / (2 * kThreads);
  int64_t cap    = (int64_t)c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  klaunch((k_elementwise<OP>), (unsigned)blocks, kThreads, 0, c->stream, a);
  LAUNCH_CHECK();
  const int reads = (OP == EW_LINCOMB) ? a.t.n
                    : (OP == EW_CONST) ? 0
                    : (OP == EW_SCALESUM || OP == EW_SCALEDIFF || OP == EW_PROD || OP == EW_DIV) ? 2 : 1;
  ALG_BYTES(reads + 1, a.n);
  return 0;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

extern "C" int b200_lincomb(b200_ctx* c, int nterms, const double* cf,
                            const double* const* v, double* z, int64_t n)
{
  if (nterms < 1 || nterms > B200_MAX_TERMS) return fail("b200_lincomb: nterms out of range");
  EwArgs a;
  memset(&a, 0, sizeof(a));
  a.t.n = nterms;
  for (int k = 0; k < nterms; k++)
  {
    a.t.c[k] = cf[k];
    a.t.v[k] = v[k];
    if (!aligned16(v[k])) return fail("b200_lincomb: operand not 16-byte aligned");
  }
  if (!aligned16(z)) return fail("b200_lincomb: output not 16-byte aligned");
  a.z = z;
  a.n = n;
  return launch_ew<EW_LINCOMB>(c, a);
}

#define EW_ENTRY_CHECK(ptr) \
  if (!aligned16(ptr)) return fail("b200 elementwise: pointer not 16-byte aligned")

extern "C" int b200_scale_sumdiff(b200_ctx* c, double s, const double* x, const double* y,
                                  int sign, double* z, int64_t n)
{
  EW_ENTRY_CHECK(x); EW_ENTRY_CHECK(y); EW_ENTRY_CHECK(z);
  EwArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.y = y; a.a = s; a.z = z; a.n = n;
  return sign >= 0 ? launch_ew<EW_SCALESUM>(c, a) : launch_ew<EW_SCALEDIFF>(c, a);
}
extern "C" int b200_const(b200_ctx* c, double v, double* z, int64_t n)
{
  EW_ENTRY_CHECK(z);
  EwArgs a;
  memset(&a, 0, sizeof(a));
  a.a = v; a.z = z; a.n = n;
  return launch_ew<EW_CONST>(c, a);
}
#define EW_BINARY(NAME, OP)                                                          \
  extern "C" int NAME(b200_ctx* c, const double* x, const double* y, double* z, int64_t n) \
  {                                                                                  \
    EW_ENTRY_CHECK(x); EW_ENTRY_CHECK(y); EW_ENTRY_CHECK(z);                         \
    EwArgs a;                                                                        \
    memset(&a, 0, sizeof(a));                                                        \
    a.x = x; a.y = y; a.z = z; a.n = n;                                              \
    return launch_ew<OP>(c, a);                                                      \
  }
EW_BINARY(b200_prod, EW_PROD)
EW_BINARY(b200_div, EW_DIV)
#define EW_UNARY(NAME, OP)                                                \
  extern "C" int NAME(b200_ctx* c, const double* x, double* z, int64_t n) \
  {                                                                       \
    EW_ENTRY_CHECK(x); EW_ENTRY_CHECK(z);                                 \
    EwArgs a;                                                             \
    memset(&a, 0, sizeof(a));                                             \
    a.x = x; a.z = z; a.n = n;                                            \
    return launch_ew<OP>(c, a);                                           \
  }
EW_UNARY(b200_abs, EW_ABS)
EW_UNARY(b200_inv, EW_INV)
extern "C" int b200_addconst(b200_ctx* c, const double* x, double b, double* z, int64_t n)
{
  EW_ENTRY_CHECK(x); EW_ENTRY_CHECK(z);
  EwArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.b = b; a.z = z; a.n = n;
  return launch_ew<EW_ADDCONST>(c, a);
}
extern "C" int b200_ewt_ss(b200_ctx* c, const double* y, double rtol, double atol,
                           double* ewt, int64_t n)
{
  EW_ENTRY_CHECK(y); EW_ENTRY_CHECK(ewt);
  EwArgs a;
  memset(&a, 0, sizeof(a));
  a.x = y; a.a = rtol; a.b = atol; a.z = ewt; a.n = n;
  return launch_ew<EW_EWT>(c, a);
}

// ---------------------------------------------------------------- reductions
static int nccl_allreduce_inplace(b200_ctx* c, double* buf, int n, int op);

// Where a reduction kernel's last block stores the result: on one rank straight into mapped pinned host memory
// (no copy engine, no extra launch -- the host only waits for the stream); with a communicator into device memory,
// all-reduced over the ranks and then published.
// One rank: the host does not wait for the stream (cudaStreamSynchronize costs several microseconds after the kernel
// has ended) but for the value itself: the slot is armed with a NaN bit pattern no reduction produces, the last block's
// 8-byte store replaces it, the host spins on the (cache-coherent, pinned) word.  Falls back to the stream
// synchronisation after ~2 ms of spinning (long queues of large kernels) -- or always, with B200_NO_POLL.
static const unsigned long long kArmed = 0x7ff8dead0000beefULL;
static int g_poll = -1;
static bool poll_on()
{
  if (g_poll < 0) g_poll = getenv("B200_NO_POLL") ? 0 : 1;
  return g_poll == 1;
}
static double* reduce_target(b200_ctx* c)
{
  if (c->comm && c->nranks > 1) return c->dev_result;
  *reinterpret_cast<volatile unsigned long long*>(c->host_result) = kArmed;
  return c->host_result_dev;
}
static int reduce_fetch(b200_ctx* c, int rop, double* out)
{
  if (c->comm && c->nranks > 1)
  {
    int rc = nccl_allreduce_inplace(c, c->dev_result, 1, rop);
    if (rc) return rc;
    return read_small(c, out, c->dev_result, 1);
  }
  volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(c->host_result);
  if (poll_on())
  {
    const double t0 = hprof_on() ? trace_now_ms() : 0.0;
    for (int spin = 0; spin < 40000; spin++)
    {
      const unsigned long long b = *w;
      if (b != kArmed)
      {
        memcpy(out, &b, sizeof(double));
        if (hprof_on()) { g_hprof.sync_ms += trace_now_ms() - t0; g_hprof.syncs++; }
        return 0;
      }
      B200_CPU_RELAX();
    }
  }
  CU_TRY(stream_sync_profiled(c->stream));
  *out = c->host_result[0];
  return 0;
}
static unsigned reduce_blocks(b200_ctx* c, int64_t n, int per_thread)
{
  int64_t n2     = (n + 1) >> 1;
  int64_t blocks = (n2 + per_thread * kThreads - 1) / (per_thread * kThreads);
  int64_t cap    = (int64_t)c->sm_count * 4;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <int KIND, int ROP>
static int run_reduce(b200_ctx* c, const double* x, const double* y, int64_t n, double* result, double ys = 0.0)
{
  if (!aligned16(x) || (y && !aligned16(y))) return fail("b200 reduce: pointer not 16-byte aligned");
  klaunch((k_reduce<KIND, ROP>), reduce_blocks(c, n, 4), kThreads, 0, c->stream, x, y, ys, n, c->partials, c->ticket, reduce_target(c));
  LAUNCH_CHECK();
  ALG_BYTES((KIND == RD_DOT || KIND == RD_WSQR) ? 2 : 1, n);
  return reduce_fetch(c, ROP, result);
}

extern "C" int b200_dot(b200_ctx* c, const double* x, const double* y, int64_t n, double* r)
{
  return run_reduce<RD_DOT, RED_SUM>(c, x, y, n, r);
}
extern "C" int b200_wsqrsum(b200_ctx* c, const double* x, const double* w, int64_t n, double* r)
{
  return run_reduce<RD_WSQR, RED_SUM>(c, x, w, n, r);
}
extern "C" int b200_wsqrsum_scalar(b200_ctx* c, const double* x, double w, int64_t n, double* r)
{
  return run_reduce<RD_WSQRC, RED_SUM>(c, x, nullptr, n, r, w);
}
extern "C" int b200_maxnorm(b200_ctx* c, const double* x, int64_t n, double* r)
{
  return run_reduce<RD_MAXNORM, RED_MAX>(c, x, nullptr, n, r);
}
extern "C" int b200_min(b200_ctx* c, const double* x, int64_t n, double* r)
{
  return run_reduce<RD_MIN, RED_MIN>(c, x, nullptr, n, r);
}
extern "C" int b200_l1norm(b200_ctx* c, const double* x, int64_t n, double* r)
{
  return run_reduce<RD_L1, RED_SUM>(c, x, nullptr, n, r);
}

// ------------------------------------------------------- fused stage kernels
#include "stage_kernels.cuh"

template <int NT, uint32_t PAT>
static void launch_march(const StageArgs& a, dim3 grid, cudaStream_t st)
{
  if (a.region == 2) klaunch((k_stage_march<NT, PAT, 2, 0>), grid, kThreads, 0, st, a);
  else if (a.rw && a.ewt_out) klaunch((k_stage_march<NT, PAT, 0, 2>), grid, kThreads, 0, st, a);
  else if (a.rw) klaunch((k_stage_march<NT, PAT, 0, 1>), grid, kThreads, 0, st, a);
  else klaunch((k_stage_march<NT, PAT, 0, 0>), grid, kThreads, 0, st, a);
}

// pick the compiled pattern for the term sequence, else the general kernel
static void dispatch_march(const StageArgs& a, dim3 grid, cudaStream_t st)
{
  uint32_t pat = 0;
  for (int k = 0; k < a.t.n; k++) pat |= (uint32_t)a.t.src[k] << (2 * k);
  const int V = B200_SRC_VECTOR, C = B200_SRC_CENTRE, S = B200_SRC_STENCIL;
  switch (a.t.n)
  {
  case 1:
    if (pat == PAT1(S)) return launch_march<1, PAT1(S)>(a, grid, st);                       // f = L(x)
    break;
  case 2:
    if (pat == PAT2(C, S)) return launch_march<2, PAT2(C, S)>(a, grid, st);                 // y + c L(y): SSP stage, STS stage 1
    if (pat == PAT2(V, S)) return launch_march<2, PAT2(V, S)>(a, grid, st);
    if (pat == PAT2(S, V)) return launch_march<2, PAT2(S, V)>(a, grid, st);                 // DQ: sig*v + y style sums
    break;
  case 3:
    if (pat == PAT3(C, V, S)) return launch_march<3, PAT3(C, V, S)>(a, grid, st);           // SSP closing LC3
    break;
  case 4:
    if (pat == PAT4(V, C, V, S)) return launch_march<4, PAT4(V, C, V, S)>(a, grid, st);     // STS embedding LC4
    break;
  case 5:
    if (pat == PAT5(S, V, V, C, V)) return launch_march<5, PAT5(S, V, V, C, V)>(a, grid, st); // RKC/RKL stage
    break;
  }
  return launch_march<B200_MAX_TERMS, PAT_RUNTIME>(a, grid, st);
}

static int g_rows_per_block = 8; // measured best on B200 at 4096^2 and 16384^2 (profiles/)
static bool g_rows_per_block_set = false;

extern "C" int b200_set_rows_per_block(int r)
{
  if (r < 1) return -1;
  g_rows_per_block     = r;
  g_rows_per_block_set = true;
  return 0;
}
// Rows a block of the one-stage kernels marches over.  Small grids are latency-bound: every block is resident at
// once and the launch lasts as long as one block's dependent row steps, so rows are halved until the grid has at
// least two blocks per SM (128^2: 128 blocks of 1 row instead of 16 blocks of 8).
static int stage_rows_auto(const b200_ctx* c, int64_t gx, int64_t ny)
{
  int rows = g_rows_per_block;
  if (g_rows_per_block_set) return rows;
  const int64_t want = 2 * (int64_t)(c->sm_count > 0 ? c->sm_count : 148);
  while (rows > 1 && gx * ((ny + rows - 1) / rows) < want) rows >>= 1;
  return rows;
}

extern "C" int b200_stencil_lincomb(b200_ctx* c, const b200_stencil_geom* g, const double* x,
                                    int nterms, const double* cf, const int* src,
                                    const double* const* v, double* z,
                                    const b200_stage_extras* ex, int region)
{
  if (nterms < 1 || nterms > B200_MAX_TERMS) return fail("b200_stencil_lincomb: nterms out of range");
  if (z == x) return fail("b200_stencil_lincomb: z must not alias the stencil input");
  if (g->nx < 2 || g->ny < 2) return fail("b200_stencil_lincomb: sub-domain must be at least 2x2");
  StageArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = g->nx; a.ny = g->ny;
  a.cxw = g->cxw; a.cxe = g->cxe; a.cys = g->cys; a.cyn = g->cyn;
  a.hw = g->halo_w; a.he = g->halo_e; a.hs = g->halo_s; a.hn = g->halo_n;
  a.x = x; a.z = z;
  a.t.n = nterms;
  int nst = 0;
  for (int k = 0; k < nterms; k++)
  {
    a.t.c[k]   = cf[k];
    a.t.src[k] = src[k];
    a.t.v[k]   = (src[k] == B200_SRC_VECTOR) ? v[k] : nullptr;
    if (src[k] == B200_SRC_STENCIL) nst++;
    if (src[k] == B200_SRC_VECTOR && !v[k]) return fail("b200_stencil_lincomb: NULL operand");
  }
  if (nst > 1) return fail("b200_stencil_lincomb: more than one stencil term");
  if (ex)
  {
    a.f_out = ex->f_out;
    a.send_w = ex->send_w; a.send_e = ex->send_e; a.send_s = ex->send_s; a.send_n = ex->send_n;
    a.rw = ex->wrms_w;
    a.result = ex->wrms_result;
    if (a.f_out == x) return fail("b200_stencil_lincomb: f_out must not alias the stencil input");
    if (ex->ewt_out)
    {
      if (!a.rw || !ex->ewt_result) return fail("b200_stencil_lincomb: ewt_out needs the fused WRMS norm and ewt_result");
      if (ex->ewt_out == x || ex->ewt_out == z || !aligned16(ex->ewt_out)) return fail("b200_stencil_lincomb: bad ewt_out");
      a.ewt_out = ex->ewt_out; a.ewt_rtol = ex->ewt_rtol; a.ewt_atol = ex->ewt_atol; a.result2 = ex->ewt_result;
    }
  }
  if (a.rw && region != 0) return fail("b200_stencil_lincomb: fused WRMS needs region 0");
  if (a.rw && !a.result) return fail("b200_stencil_lincomb: wrms_result missing");
  a.partials = c->partials;
  a.ticket   = c->ticket;
  a.region   = region;

  if (region == 1)
  {
    int64_t cells  = 2 * a.nx + 2 * (a.ny - 2);
    unsigned blocks = (unsigned)((cells + kThreads - 1) / kThreads);
    klaunch(k_stage_ring, blocks, kThreads, 0, c->stream, a);
    LAUNCH_CHECK();
    return 0;
  }

  bool fast = (a.nx % 2 == 0) && aligned16(x) && aligned16(z) && (!a.f_out || aligned16(a.f_out)) &&
              aligned16(a.cxw) && aligned16(a.cxe) && (!a.rw || aligned16(a.rw)) &&
              (!a.hs || aligned16(a.hs)) && (!a.hn || aligned16(a.hn)) &&
              (!a.send_s || aligned16(a.send_s)) && (!a.send_n || aligned16(a.send_n));
  for (int k = 0; k < nterms; k++)
    if (a.t.v[k] && !aligned16(a.t.v[k])) fast = false;

  if (fast)
  {
    int64_t gx    = (a.nx / 2 + kThreads - 1) / kThreads;
    a.rows        = stage_rows_auto(c, gx, a.ny);
    int64_t gy    = (a.ny + a.rows - 1) / a.rows;
    if (a.ny >= (int64_t)1 << 31) return fail("b200_stencil_lincomb: ny must be below 2^31");
    if (gy > 65535)
    {
      a.rows = (int)((a.ny + 65534) / 65535);
      gy     = (a.ny + a.rows - 1) / a.rows;
    }
    if (a.rw && gx * gy * (a.ewt_out ? 2 : 1) > kMaxPartials) return fail("b200_stencil_lincomb: too many blocks for fused WRMS");
    dim3 grid((unsigned)gx, (unsigned)gy);
    dispatch_march(a, grid, c->stream);
    LAUNCH_CHECK();
  }
  else
  {
    int64_t gx = (a.nx + kThreads - 1) / kThreads;
    if (a.ny > 65535) return fail("b200_stencil_lincomb: generic path supports ny <= 65535");
    if (a.rw && gx * a.ny * (a.ewt_out ? 2 : 1) > kMaxPartials) return fail("b200_stencil_lincomb: too many blocks for fused WRMS");
    dim3 grid((unsigned)gx, (unsigned)a.ny);
    klaunch(k_stage_generic, grid, kThreads, 0, c->stream, a);
    LAUNCH_CHECK();
  }
  {
    int touches = 2 + (a.f_out ? 1 : 0) + (a.rw ? 1 : 0) + (a.ewt_out ? 1 : 0); // x, z
    for (int k = 0; k < nterms; k++) touches += (src[k] == B200_SRC_VECTOR);
    ALG_BYTES(touches, a.nx * a.ny);
  }
  return 0;
}

// ------------------------------------------------ implicit path: fused vector work
#include "dq_kernels.cuh"

extern "C" int b200_lin2_wsqrsum(b200_ctx* c, double ca, const double* a, double cb, const double* b, const double* w,
                                 double wscalar, double* z, int64_t n, double* result)
{
  if (!aligned16(a) || !aligned16(b) || !aligned16(z) || (w && !aligned16(w))) return fail("b200_lin2_wsqrsum: pointer not 16-byte aligned");
  Lin2RedArgs r;
  memset(&r, 0, sizeof(r));
  r.a = a; r.b = b; r.w = w; r.ca = ca; r.cb = cb; r.ws = wscalar; r.z = z; r.n = n;
  r.partials = c->partials; r.ticket = c->ticket; r.result = reduce_target(c);
  if (w) klaunch((k_lin2_wsqr<true>), reduce_blocks(c, n, 2), kThreads, 0, c->stream, r);
  else klaunch((k_lin2_wsqr<false>), reduce_blocks(c, n, 2), kThreads, 0, c->stream, r);
  LAUNCH_CHECK();
  ALG_BYTES(w ? 4 : 3, n);
  return reduce_fetch(c, RED_SUM, result);
}

extern "C" int b200_ewt_ss_wsqrsum(b200_ctx* c, const double* y, double rtol, double atol, double* ewt, int64_t n, double* result)
{
  if (!aligned16(y) || !aligned16(ewt)) return fail("b200_ewt_ss_wsqrsum: pointer not 16-byte aligned");
  EwtArgs r;
  memset(&r, 0, sizeof(r));
  r.y = y; r.rtol = rtol; r.atol = atol; r.ewt = ewt; r.n = n;
  r.partials = c->partials; r.ticket = c->ticket; r.result = reduce_target(c);
  klaunch(k_ewt_wsqr, reduce_blocks(c, n, 2), kThreads, 0, c->stream, r);
  LAUNCH_CHECK();
  ALG_BYTES(2, n);
  return reduce_fetch(c, RED_SUM, result);
}

extern "C" int b200_prod_dot(b200_ctx* c, const double* a, const double* b, const double* cc, double* z, int64_t n, double* result)
{
  if (!aligned16(a) || !aligned16(b) || !aligned16(cc) || !aligned16(z)) return fail("b200_prod_dot: pointer not 16-byte aligned");
  ProdDotArgs r;
  memset(&r, 0, sizeof(r));
  r.a = a; r.b = b; r.c = cc; r.z = z; r.n = n;
  r.partials = c->partials; r.ticket = c->ticket; r.result = reduce_target(c);
  klaunch(k_prod_dot, reduce_blocks(c, n, 2), kThreads, 0, c->stream, r);
  LAUNCH_CHECK();
  ALG_BYTES((cc == a || cc == b) ? 3 : 4, n);
  return reduce_fetch(c, RED_SUM, result);
}

extern "C" int b200_stencil_dq(b200_ctx* c, const b200_stencil_geom* g, const double* v, const double* y, const double* fy,
                               double sigma, double siginv, int outer, double ca, double cb, double* z, double* dot_result)
{
  if (g->halo_w || g->halo_e || g->halo_s || g->halo_n) return fail("b200_stencil_dq: one periodic rank only");
  if ((g->nx & 1) || g->nx < 2 || g->ny < 2) return fail("b200_stencil_dq: needs even nx >= 2 and ny >= 2");
  if (g->ny >= (int64_t)1 << 31) return fail("b200_stencil_dq: ny too large");
  if (!aligned16(v) || !aligned16(y) || !aligned16(fy) || !aligned16(z) || !aligned16(g->cxw) || !aligned16(g->cxe))
    return fail("b200_stencil_dq: pointer not 16-byte aligned");
  if (z == v || z == y || z == fy) return fail("b200_stencil_dq: z aliases an input");
  DqArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = g->nx; a.ny = g->ny; a.cxw = g->cxw; a.cxe = g->cxe; a.cys = g->cys; a.cyn = g->cyn;
  a.v = v; a.y = y; a.fy = fy; a.sigma = sigma; a.siginv = siginv; a.ca = ca; a.cb = cb; a.outer = outer;
  a.want_dot = dot_result ? 1 : 0;
  a.z = z;
  int64_t gx    = (a.nx / 2 + kThreads - 1) / kThreads;
  a.rows        = stage_rows_auto(c, gx, a.ny);
  int64_t gy    = (a.ny + a.rows - 1) / a.rows;
  if (gy > 65535) { a.rows = (int)((a.ny + 65534) / 65535); gy = (a.ny + a.rows - 1) / a.rows; }
  if (dot_result && gx * gy > kMaxPartials) return fail("b200_stencil_dq: too many blocks for the fused dot product");
  a.partials = c->partials; a.ticket = c->ticket; a.result = reduce_target(c);
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (dot_result) klaunch((k_dq_march<true>), grid, kThreads, 0, c->stream, a);
  else klaunch((k_dq_march<false>), grid, kThreads, 0, c->stream, a);
  LAUNCH_CHECK();
  ALG_BYTES(4, a.nx * a.ny);
  if (dot_result) return reduce_fetch(c, RED_SUM, dot_result);
  return 0;
}

// ------------------------------------------- temporally blocked STS stages
#include "chain_march_inst.cuh"
#ifndef B200_HOST_EMU // (the emulated build is one translation unit: it instantiates the kernels implicitly)
B200_CHAIN_K2(B200_CHAIN_DECLARE)
B200_CHAIN_K3(B200_CHAIN_DECLARE)
B200_CHAIN_K4(B200_CHAIN_DECLARE)
B200_CHAIN_K4S(B200_CHAIN_DECLARE)
B200_CHAIN_K4B(B200_CHAIN_DECLARE_BULK)
B200_CHAIN_K4P(B200_CHAIN_DECLARE)
B200_CHAIN_K5(B200_CHAIN_DECLARE)
B200_CHAIN_K6(B200_CHAIN_DECLARE)
#endif
#include "chain_quad.cuh"

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: remember what was configured per
// (instantiation, device); each instantiation has its own table (a static of the template function)
static const int kMaxDevices = 64;
static int current_device()
{
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}

static bool g_chain_preload_only = false;
// SPLIT flavour of k_chain_march (chain_march.cuh): the upper half of the levels one row late, computed first -- two
// independent instruction streams per row step.  Instantiated for the depth(s) listed in chain_split_available (each one costs 8 more instantiations to compile).
static int g_chain_split = -1; // -1: not decided yet (B200_CHAIN_SPLIT, else the default below)
static const int kChainSplitDefault = 0;
extern "C" int b200_set_chain_split(int on)
{
  g_chain_split = on ? 1 : 0;
  return 0;
}
extern "C" int b200_get_chain_split(void)
{
  if (g_chain_split < 0)
  {
    const char* e = getenv("B200_CHAIN_SPLIT");
    g_chain_split = e ? (atoi(e) != 0) : kChainSplitDefault;
  }
  return g_chain_split;
}
template <int K, int PF, bool FMA> constexpr bool chain_split_available() { return !FMA && K == 4 && PF == 3; }

// BULK flavour of k_chain_march (chain_march.cuh): the operand ring is filled by cp.async.bulk (one 512-byte copy per
// warp, operand and row, issued by one lane, completed on an mbarrier) instead of one 16-byte cp.async per thread.
static int g_chain_bulk = -1; // -1: not decided yet (B200_CHAIN_BULK, else the default below)
static const int kChainBulkDefault = 2; // measured: profiles/r02_bench_n1_default_ab_call_aa_*, DESIGN 4.1b
extern "C" int b200_set_chain_bulk(int on)
{ // < 0: back to the initial value (B200_CHAIN_BULK, else the default)
  g_chain_bulk = on < 0 ? -1 : (on > 2 ? 2 : on); // 1: prefetch depth of the plain flavour, 2: one row deeper
  return 0;
}
extern "C" int b200_get_chain_bulk(void)
{
  if (g_chain_bulk < 0)
  {
    const char* e = getenv("B200_CHAIN_BULK");
    g_chain_bulk  = e ? atoi(e) : kChainBulkDefault;
    if (g_chain_bulk < 0 || g_chain_bulk > 2) g_chain_bulk = kChainBulkDefault;
  }
  return g_chain_bulk;
}
template <int K, int PF, bool FMA> constexpr bool chain_bulk_available() { return !FMA && K == 4 && PF == 3; }

template <int K, int PF, bool HALO, bool FMA, bool UNI, bool HEAD, bool SPLIT, bool BULK = false>
static int launch_chain_s(const ChainArgs& a, dim3 grid, cudaStream_t st)
{
  const size_t smem = chain_march_smem(K, PF, a.rows, SPLIT, BULK);
  static size_t configured_on[kMaxDevices] = {};
  size_t& configured = configured_on[current_device()];
  if (smem > configured)
  {
    CU_TRY(cudaFuncSetAttribute(k_chain_march<K, PF, HALO, FMA, UNI, HEAD, SPLIT, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  if (g_chain_preload_only)
  { // b200_stencil_chain_preload: make the driver load this instantiation now (lazy module loading would do it at
    // the first launch, milliseconds into somebody's time step), launch nothing
    cudaFuncAttributes fa;
    CU_TRY(cudaFuncGetAttributes(&fa, k_chain_march<K, PF, HALO, FMA, UNI, HEAD, SPLIT, BULK>));
    return 0;
  }
  klaunch((k_chain_march<K, PF, HALO, FMA, UNI, HEAD, SPLIT, BULK>), grid, kChainThreads, smem, st, a);
  return 0;
}
template <int K, int PF, bool HALO, bool FMA, bool UNI, bool HEAD>
static int launch_chain_k(const ChainArgs& a, dim3 grid, cudaStream_t st)
{
  if constexpr (chain_split_available<K, PF, FMA>()) // (opt-in: an explicit request wins over the default ring)
    if (b200_get_chain_split()) return launch_chain_s<K, PF, HALO, FMA, UNI, HEAD, true>(a, grid, st);
  if constexpr (chain_bulk_available<K, PF, FMA>())
  {
    if (b200_get_chain_bulk() == 2) return launch_chain_s<K, PF + 1, HALO, FMA, UNI, HEAD, false, true>(a, grid, st);
    if (b200_get_chain_bulk()) return launch_chain_s<K, PF, HALO, FMA, UNI, HEAD, false, true>(a, grid, st);
  }
  return launch_chain_s<K, PF, HALO, FMA, UNI, HEAD, false>(a, grid, st);
}

// 0: two roundings per multiply-add, bit-identical to the reference's baseline x86-64 build (default);
// 1: the chain kernel's multiply-adds are contracted to FMA (one rounding)
static int g_contract = 0;

extern "C" int b200_set_contract(int on)
{
  g_contract = on ? 1 : 0;
  return 0;
}
extern "C" int b200_get_contract(void) { return g_contract; }

template <int K, int PF, bool HALO, bool FMA, bool UNI>
static int launch_chain_h(const ChainArgs& a, dim3 grid, cudaStream_t st)
{ // HEAD: the chain begins with stage 1 of the step
  return a.head ? launch_chain_k<K, PF, HALO, FMA, UNI, true>(a, grid, st) : launch_chain_k<K, PF, HALO, FMA, UNI, false>(a, grid, st);
}
template <int K, int PF, bool HALO, bool FMA>
static int launch_chain_u(const ChainArgs& a, dim3 grid, cudaStream_t st, bool uni)
{
  return uni ? launch_chain_h<K, PF, HALO, FMA, true>(a, grid, st) : launch_chain_h<K, PF, HALO, FMA, false>(a, grid, st);
}
// Depth 4, exact arithmetic, plain flavour with one more row in flight (PF = 4: the yn / fn ring is then 8 rows deep and
// its slot arithmetic a mask).  B200_CHAIN_PF=4 / b200_set_chain_pf(4); instantiated in csrc/chain_inst_k4p.cu.
static int g_chain_pf = -1;
static const int kChainPfDefault = 3;
extern "C" int b200_set_chain_pf(int pf)
{
  g_chain_pf = (pf == 3 || pf == 4) ? pf : -1;
  return 0;
}
extern "C" int b200_get_chain_pf(void)
{
  if (g_chain_pf < 0)
  {
    const char* e = getenv("B200_CHAIN_PF");
    g_chain_pf    = (e && atoi(e) == 4) ? 4 : ((e && atoi(e) == 3) ? 3 : kChainPfDefault);
  }
  return g_chain_pf;
}
template <int K, int PF>
static int launch_chain(const ChainArgs& a, dim3 grid, cudaStream_t st, bool uni)
{
  if constexpr (K == 4 && PF == 3)
    if (!g_contract && !b200_get_chain_bulk() && !b200_get_chain_split() && b200_get_chain_pf() == 4)
      return a.hx ? launch_chain_u<4, 4, true, false>(a, grid, st, uni) : launch_chain_u<4, 4, false, false>(a, grid, st, uni);
  if (g_contract)
    return a.hx ? launch_chain_u<K, PF, true, true>(a, grid, st, uni) : launch_chain_u<K, PF, false, true>(a, grid, st, uni);
  return a.hx ? launch_chain_u<K, PF, true, false>(a, grid, st, uni) : launch_chain_u<K, PF, false, false>(a, grid, st, uni);
}

template <int K, int PF, bool HALO, bool FMA, int MINB>
static int launch_quad_k(const ChainArgs& a, dim3 grid, cudaStream_t st)
{
  const size_t smem = chain_quad_smem(K, PF, a.rows);
  static size_t configured_on[kMaxDevices] = {};
  size_t& configured = configured_on[current_device()];
  if (smem > configured)
  {
    CU_TRY(cudaFuncSetAttribute(k_chain_quad<K, PF, HALO, FMA, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  klaunch((k_chain_quad<K, PF, HALO, FMA, MINB>), grid, kQuadThreads, smem, st, a);
  return 0;
}
template <int K, int PF, int MINB>
static int launch_quad(const ChainArgs& a, dim3 grid, cudaStream_t st)
{
  if (g_contract)
    return a.hx ? launch_quad_k<K, PF, true, true, MINB>(a, grid, st) : launch_quad_k<K, PF, false, true, MINB>(a, grid, st);
  return a.hx ? launch_quad_k<K, PF, true, false, MINB>(a, grid, st) : launch_quad_k<K, PF, false, false, MINB>(a, grid, st);
}

// A grid of slightly more than a whole number of waves (all blocks of a wave start together and, on a small grid, end
// together) leaves a few straggler blocks running on an otherwise empty machine for one more block duration: 2048^2,
// depth 4, 32 rows = 320 blocks for 296 slots took 0.056 ms per launch, 37 rows = 280 blocks 0.045 ms
// (profiles/r02_kbench_chain_rows_wave_fit.log).  Where the block count exceeds w = 1..3 waves by at most a quarter
// wave, the rows are raised until the grid fits w waves.  B200_NO_WAVE_FIT=1 switches it off (A/B).
static int rows_fit_waves(int64_t ny, int64_t gx, int rows, int64_t resident)
{
  static const bool off = getenv("B200_NO_WAVE_FIT") != nullptr;
  if (off || gx <= 0 || resident <= 0 || rows <= 0) return rows;
  const int64_t blocks = gx * ((ny + rows - 1) / rows);
  const int64_t w      = blocks / resident;
  if (w < 1 || w > 3 || blocks == w * resident || 4 * (blocks - w * resident) > resident) return rows;
  const int64_t gy = (w * resident) / gx; // block rows that fit w waves
  if (gy < 1) return rows;
  const int r2 = (int)((ny + gy - 1) / gy);
  if (r2 <= rows || r2 > 2 * rows || gx * ((ny + r2 - 1) / r2) > w * resident) return rows;
  return r2;
}

static int g_chain_rows = 0; // 0 = automatic (chain_rows_auto), else what b200_set_chain_rows asked for
// Rows each block marches over: more rows = less redundant work (a block starts K-1 rows early and the
// K-1 rows around it are recomputed: (rows + 2(K-1)) / rows) but fewer blocks.  256 rows where that still
// leaves >= 4 waves of blocks (2 resident blocks on each SM), else 128, else 64, else 32.  Measured at 16384^2,
// K = 4: 62.4 / 57.5 / 54.4 / 53.6 ms per step for 32 / 64 / 128 / 256 rows (profiles/r01_bench_chain_rows.log).
static int chain_rows_auto_sm(int sm_count, int64_t nx, int64_t ny, int nstages, bool quad)
{
  const int64_t sms   = sm_count > 0 ? sm_count : 148;
  const int use       = quad ? 128 - 4 * ((nstages + 1) / 2) : 64 - 4 * ((nstages + 1) / 2);
  const int wpb       = quad ? kQuadThreads / 32 : kChainThreads / 32;
  const int64_t gx    = ((nx + use - 1) / use + wpb - 1) / wpb;
  const int64_t waves = 4 * 2 * sms;
  const int cand[4]   = {256, 128, 64, 32};
  const int64_t slots = 2 * sms;
  for (int r : cand)
    if (gx * ((ny + r - 1) / r) >= waves) return r;
  if (gx * ((ny + 31) / 32) > slots) return rows_fit_waves(ny, gx, 32, slots);
  // Small grids: every block is resident at once and the launch takes as long as ONE block needs for its
  // rows + 2(K-1) row steps, so the fewest rows that still fit the machine in one wave win (128^2, K = 6: 16 blocks
  // of 8 rows = 18 row steps instead of 4 blocks of 42).
  const int64_t resident = 2 * sms;
  const int small[2]     = {8, 16};
  for (int r : small)
    if (gx * ((ny + r - 1) / r) <= resident) return r;
  return 32;
}
static int chain_rows_auto(const b200_ctx* c, int64_t nx, int64_t ny, int nstages, bool quad)
{
  return chain_rows_auto_sm(c->sm_count, nx, ny, nstages, quad);
}
// the automatic choice for an nx x ny block and a chain of nstages stages (k_chain_march) on a device with sm_count
// multiprocessors (0: the context's own device) -- for tests and tuning scripts; launches nothing
extern "C" int b200_chain_rows_query(b200_ctx* c, int64_t nx, int64_t ny, int nstages, int sm_count)
{
  if (nstages < 2 || nstages > B200_MAX_CHAIN || nx < 2 || ny < 1) return -1;
  return chain_rows_auto_sm(sm_count > 0 ? sm_count : (c ? c->sm_count : 148), nx, ny, nstages, false);
}
static int g_chain_uniform = 1; // honour b200_stencil_geom.uniform (0: always load the tables; for A/B tests)
extern "C" int b200_set_chain_uniform(int on)
{
  g_chain_uniform = on ? 1 : 0;
  return 0;
}
static const char* g_last_chain_kernel = "";
extern "C" const char* b200_last_chain_kernel(void) { return g_last_chain_kernel; }
// 0 (default): k_chain_march, two cells per thread; 1: k_chain_quad, four cells per thread.
// B200_CHAIN_VARIANT overrides the initial value.
static int g_chain_variant = -1;

extern "C" int b200_set_chain_variant(int v)
{
  if (v != 0 && v != 1) return -1;
  g_chain_variant = v;
  return 0;
}
extern "C" int b200_get_chain_variant(void)
{
  if (g_chain_variant < 0)
  {
    const char* e   = getenv("B200_CHAIN_VARIANT");
    g_chain_variant = (e && e[0] == '1') ? 1 : 0;
  }
  return g_chain_variant;
}

extern "C" int b200_set_chain_rows(int r)
{
  if (r < 0) return -1; // 0 = back to automatic
  g_chain_rows = r;
  return 0;
}

static int stencil_chain_common(b200_ctx* c, const b200_stencil_geom* g, int nstages, const double* x,
                                const double* prev2, const double* yn, const double* fn,
                                const double* coeffs, double* const* z_out, const double* const* halos,
                                int hg, int hg2, double* f_out = nullptr, bool head = false)
{
  if (nstages < 2 || nstages > B200_MAX_CHAIN) return fail("b200_stencil_chain: nstages must be 2..B200_MAX_CHAIN");
  if (f_out && (f_out == x || !aligned16(f_out))) return fail("b200_stencil_chain_head: bad f_out");
  if (g->halo_w || g->halo_e || g->halo_s || g->halo_n)
    return fail("b200_stencil_chain: the one-deep halo buffers of b200_stencil_geom are not used here");
  if ((g->nx & 1) || g->nx < 128 || g->ny < 16) return fail("b200_stencil_chain: needs even nx >= 128 and ny >= 16");
  if (g->ny >= (int64_t)1 << 30) return fail("b200_stencil_chain: ny too large");
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = g->nx; a.ny = g->ny;
  a.cxw = g->cxw; a.cxe = g->cxe; a.cys = g->cys; a.cyn = g->cyn;
  a.x = x; a.prev2 = prev2; a.yn = yn; a.fn = fn;
  // (the uniform flavour shares the x-direction products between neighbours: it needs Dx_w == Dx_e to the bit --
  // always the case for this problem, kx / dx^2 on both faces; anything else runs on the tables, which are still valid)
  const bool uni = g->uniform != 0 && g_chain_uniform && memcmp(&g->u_cxw, &g->u_cxe, sizeof(double)) == 0;
  if (uni)
  { // centre coefficient in the reference's association, diffusion.cpp:48 (host IEEE adds = device DADD)
    a.u_cxw = g->u_cxw; a.u_cxe = g->u_cxe; a.u_cys = g->u_cys; a.u_cyn = g->u_cyn;
    volatile double sx = g->u_cxw + g->u_cxe, sy = g->u_cys + g->u_cyn;
    volatile double sc = sx + sy;
    a.u_ndc = -sc;
  }
  if (!aligned16(x) || !aligned16(prev2) || !aligned16(yn) || !aligned16(fn) || !aligned16(a.cxw) || !aligned16(a.cxe))
    return fail("b200_stencil_chain: operand not 16-byte aligned");
  if (halos)
  {
    if (hg < nstages || (hg2 & 1) || hg2 < 2 * ((nstages + 1) / 2)) return fail("b200_stencil_chain_halo: halo too shallow");
    a.hx = halos[0]; a.hp = halos[1]; a.hy = halos[2]; a.hf = halos[3];
    a.g = hg; a.g2 = hg2;
    if (!a.hx || !a.hp || !a.hy || !a.hf) return fail("b200_stencil_chain_halo: NULL halo buffer");
    if (!aligned16(a.hx) || !aligned16(a.hp) || !aligned16(a.hy) || !aligned16(a.hf))
      return fail("b200_stencil_chain_halo: halo buffer not 16-byte aligned");
  }
  bool any = false;
  for (int l = 0; l < nstages; l++)
  {
    for (int k = 0; k < 5; k++) a.c[l][k] = coeffs[5 * l + k];
    a.out[l] = z_out[l];
    if (a.out[l])
    {
      any = true;
      if (!aligned16(a.out[l])) return fail("b200_stencil_chain: output not 16-byte aligned");
      if (a.out[l] == x || a.out[l] == prev2 || a.out[l] == yn || a.out[l] == fn || a.out[l] == f_out)
        return fail("b200_stencil_chain: an output aliases an input");
    }
  }
  if (!any || !a.out[nstages - 1]) return fail("b200_stencil_chain: the last stage must be stored");
  a.f_out = f_out;
  a.head  = head ? 1 : 0;
  // (k_chain_quad has no stage-1 flavour: a chain that begins the step runs on k_chain_march)
  const bool use_quad = !head && b200_get_chain_variant() == 1 && chain_quad_supported(a.nx, a.ny, nstages, halos ? hg2 : -1);
  a.rows              = g_chain_rows > 0 ? g_chain_rows : chain_rows_auto(c, a.nx, a.ny, nstages, use_quad);
  int rc = 0;
  if (use_quad)
  {
    dim3 grid = chain_quad_grid(a.nx, a.ny, nstages, &a.rows);
    g_last_chain_kernel = "k_chain_quad";
    switch (nstages)
    {
    case 2: rc = launch_quad<2, kQuadPF, 2>(a, grid, c->stream); break;
    case 3: rc = launch_quad<3, kQuadPF, 2>(a, grid, c->stream); break;
    case 4: rc = launch_quad<4, kQuadPF, 2>(a, grid, c->stream); break;
    case 5: rc = launch_quad<5, kQuadPF, 2>(a, grid, c->stream); break;
    default: rc = launch_quad<6, kQuadPF, 2>(a, grid, c->stream); break;
    }
  }
  else
  {
    dim3 grid = chain_march_grid(a.nx, a.ny, nstages, &a.rows);
    g_last_chain_kernel = "k_chain_march";
    switch (nstages)
    {
    case 2: rc = launch_chain<2, 4>(a, grid, c->stream, uni); break;
    case 3: rc = launch_chain<3, 4>(a, grid, c->stream, uni); break;
    case 4: rc = launch_chain<4, 3>(a, grid, c->stream, uni); break;
    case 5: rc = launch_chain<5, 3>(a, grid, c->stream, uni); break;
    default: rc = launch_chain<6, 3>(a, grid, c->stream, uni); break;
    }
  }
  if (rc) return rc;
  LAUNCH_CHECK();
  {
    int touches = head ? (f_out ? 2 : 1) : 4; // x, prev2, yn, fn -- or x (and the stored f_n) when the chain begins the step
    for (int l = 0; l < nstages; l++) touches += (a.out[l] != nullptr);
    ALG_BYTES(touches, a.nx * a.ny);
  }
  return 0;
}

// Load every k_chain_march instantiation a session with these properties can reach -- depths 2..B200_MAX_CHAIN, with
// and without the stage-1 head -- and opt it into its shared memory.  With CUDA's lazy module loading a kernel is
// loaded at its first launch (several milliseconds); an adaptive run changes its stage count, hence the depth of the
// last chain of a step, at any time, and would pay that inside some time step.
extern "C" int b200_stencil_chain_preload(b200_ctx* c, int halo, int uniform)
{
  CU_TRY(cudaSetDevice(c->device));
  static const double dummy = 0.0;
  int rc = 0;
  g_chain_preload_only = true;
  for (int head = 0; head < 2 && !rc; head++)
    for (int k = 2; k <= B200_MAX_CHAIN && !rc; k++)
    {
      ChainArgs a;
      memset(&a, 0, sizeof(a));
      a.rows = 256;
      a.head = head;
      a.hx   = halo ? &dummy : nullptr;
      const bool uni = uniform && g_chain_uniform;
      const dim3 grid(1, 1);
      switch (k)
      {
      case 2: rc = launch_chain<2, 4>(a, grid, c->stream, uni); break;
      case 3: rc = launch_chain<3, 4>(a, grid, c->stream, uni); break;
      case 4: rc = launch_chain<4, 3>(a, grid, c->stream, uni); break;
      case 5: rc = launch_chain<5, 3>(a, grid, c->stream, uni); break;
      default: rc = launch_chain<6, 3>(a, grid, c->stream, uni); break;
      }
    }
  g_chain_preload_only = false;
  return rc;
}

// The chain that BEGINS a step: stage 1 is z_1 = x + c1 L(x) (x = y_n; coeffs[0] = c1, coeffs[1..4] unused) and
// f_n = L(x) is written to f_out and used by the later stages straight from the kernel's ring -- y_n is read once,
// no z_{-1}, no f_n stream.
extern "C" int b200_stencil_chain_head(b200_ctx* c, const b200_stencil_geom* g, int nstages, const double* x,
                                       const double* coeffs, double* const* z_out, double* f_out,
                                       const double* halo_x, int halo_rows, int halo_cols)
{
  if (halo_x)
  {
    const double* halos[4] = {halo_x, halo_x, halo_x, halo_x};
    return stencil_chain_common(c, g, nstages, x, x, x, x, coeffs, z_out, halos, halo_rows, halo_cols, f_out, true);
  }
  return stencil_chain_common(c, g, nstages, x, x, x, x, coeffs, z_out, nullptr, 0, 0, f_out, true);
}

extern "C" int b200_stencil_chain(b200_ctx* c, const b200_stencil_geom* g, int nstages,
                                  const double* x, const double* prev2, const double* yn,
                                  const double* fn, const double* coeffs, double* const* z_out)
{
  return stencil_chain_common(c, g, nstages, x, prev2, yn, fn, coeffs, z_out, nullptr, 0, 0);
}

extern "C" int b200_stencil_chain_halo(b200_ctx* c, const b200_stencil_geom* g, int nstages,
                                       const double* x, const double* prev2, const double* yn,
                                       const double* fn, const double* coeffs, double* const* z_out,
                                       const double* const* halos, int halo_rows, int halo_cols)
{
  if (!halos) return fail("b200_stencil_chain_halo: halos missing");
  return stencil_chain_common(c, g, nstages, x, prev2, yn, fn, coeffs, z_out, halos, halo_rows, halo_cols);
}

// ------------------------------------------------------------ deep halo exchange
#include "halo_kernels.cuh"

extern "C" int64_t b200_deep_halo_doubles(int64_t nx, int64_t ny, int g, int g2)
{
  return 2 * (int64_t)g * nx + 2 * (ny + 2 * (int64_t)g) * g2;
}

static int nccl_load();
static int deep_halo_nccl_phase(b200_ctx* c, int nfields, const int peers[2], const double* const* send_lo,
                                const double* const* send_hi, double* const* recv_lo, double* const* recv_hi,
                                size_t count);

extern "C" int b200_deep_halo_exchange(b200_ctx* c, const int peers[4], int x_split, int y_split,
                                       int64_t nx, int64_t ny, int g, int g2, int nfields,
                                       const double* const* fields, double* const* halos)
{
  if (nfields < 1 || nfields > 4) return fail("b200_deep_halo_exchange: 1..4 fields");
  if (g < 1 || g > ny || g2 < 2 || (g2 & 1) || g2 > nx) return fail("b200_deep_halo_exchange: bad halo depth");
  if ((x_split || y_split) && !c->comm) return fail("b200_deep_halo_exchange: communicator not initialised");
  const int64_t srow  = (int64_t)g * nx;             // doubles in one S / N block
  const int64_t strip = (ny + 2 * (int64_t)g) * g2;  // doubles in one W / E strip
  // ---- phase 1: S / N blocks (contiguous rows, no packing)
  if (y_split)
  {
    const double* lo[4]; const double* hi[4]; double* rlo[4]; double* rhi[4];
    for (int f = 0; f < nfields; f++)
    {
      lo[f]  = fields[f];                   // my rows 0..g-1      -> S neighbour's N halo
      hi[f]  = fields[f] + (ny - g) * nx;   // my rows ny-g..ny-1  -> N neighbour's S halo
      rlo[f] = halos[f];                    // S halo  <- S neighbour's top rows
      rhi[f] = halos[f] + srow;             // N halo  <- N neighbour's bottom rows
    }
    const int py[2] = {peers[2], peers[3]};
    int rc = deep_halo_nccl_phase(c, nfields, py, lo, hi, rlo, rhi, (size_t)srow);
    if (rc) return rc;
  }
  else
  {
    for (int f = 0; f < nfields; f++)
    { // one rank in y: the periodic neighbour is this rank itself
      CU_TRY(cudaMemcpyAsync(halos[f], fields[f] + (ny - g) * nx, sizeof(double) * srow, cudaMemcpyDeviceToDevice, c->stream));
      CU_TRY(cudaMemcpyAsync(halos[f] + srow, fields[f], sizeof(double) * srow, cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  // ---- phase 2: W / E strips including the corner rows received in phase 1
  const unsigned blocks = (unsigned)((strip + kThreads - 1) / kThreads);
  if (x_split)
  {
    const size_t need = (size_t)nfields * 2 * strip;
    if (c->strip_cap < need)
    {
      if (c->strips) CU_TRY(cudaFree(c->strips));
      CU_TRY(cudaMalloc(&c->strips, sizeof(double) * need));
      c->strip_cap = need;
    }
    const double* lo[4]; const double* hi[4]; double* rlo[4]; double* rhi[4];
    for (int f = 0; f < nfields; f++)
    {
      double* ws = c->strips + (size_t)(2 * f) * strip;
      double* es = ws + strip;
      klaunch(k_pack_strips, blocks, kThreads, 0, c->stream, fields[f], halos[f], nx, ny, g, g2, ws, es);
      LAUNCH_CHECK();
      lo[f]  = ws;                              // my west columns -> W neighbour's E halo
      hi[f]  = es;                              // my east columns -> E neighbour's W halo
      rlo[f] = halos[f] + 2 * srow;             // W halo <- W neighbour's east columns
      rhi[f] = halos[f] + 2 * srow + strip;     // E halo <- E neighbour's west columns
    }
    const int px[2] = {peers[0], peers[1]};
    int rc = deep_halo_nccl_phase(c, nfields, px, lo, hi, rlo, rhi, (size_t)strip);
    if (rc) return rc;
  }
  else
  {
    for (int f = 0; f < nfields; f++)
    { // one rank in x: my own east columns are my W halo and vice versa
      klaunch(k_pack_strips, blocks, kThreads, 0, c->stream, fields[f], halos[f], nx, ny, g, g2,
                                                        halos[f] + 2 * srow + strip, halos[f] + 2 * srow);
      LAUNCH_CHECK();
    }
  }
  return 0;
}


// ------------------------------------------------------------------ halo pack
extern "C" int b200_pack_halo(b200_ctx* c, const double* u, int64_t nx, int64_t ny,
                              double* sw, double* se, double* ss, double* sn)
{
  int64_t m = nx > ny ? nx : ny;
  klaunch(k_pack, (unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, c->stream, u, nx, ny, sw, se, ss, sn);
  LAUNCH_CHECK();
  return 0;
}

// --------------------------------------------------------------- Jacobi setup
extern "C" int b200_jacobi_setup(b200_ctx* c, int64_t nx, int64_t ny, const double* pxw,
                                 const double* pxe, const double* pys, const double* pyn,
                                 double gamma, double* diag)
{
  if (ny > 65535) return fail("b200_jacobi_setup: ny <= 65535");
  dim3 grid((unsigned)((nx + kThreads - 1) / kThreads), (unsigned)ny);
  klaunch(k_jacobi, grid, kThreads, 0, c->stream, nx, ny, pxw, pxe, pys, pyn, gamma, diag);
  LAUNCH_CHECK();
  ALG_BYTES(1, nx * ny);
  return 0;
}

// -------------------------------------------------------------- adr kernels
#include "adr_kernels.cuh"

static int launch_adr(b200_ctx* c, AdrArgs& a, int mode)
{
  const int64_t gx = (a.nx + kThreads - 1) / kThreads;
  // enough blocks for >= 2 waves of 148 SMs x 3 resident blocks (76-80 registers) before the strips get longer
  a.rows = (gx * ((a.ny + 15) / 16) >= 2 * 148 * 3) ? 16 : 8;
  int64_t gy = (a.ny + a.rows - 1) / a.rows;
  if (gy > 65535)
  {
    a.rows = (int)((a.ny + 65534) / 65535);
    gy     = (a.ny + a.rows - 1) / a.rows;
  }
  dim3 grid((unsigned)gx, (unsigned)gy);
  switch (mode)
  {
  case 1: klaunch((k_adr_march<1>), grid, kThreads, 0, c->stream, a); break;
  case 2: klaunch((k_adr_march<2>), grid, kThreads, 0, c->stream, a); break;
  case 3: klaunch((k_adr_march<3>), grid, kThreads, 0, c->stream, a); break;
  case 4: klaunch((k_adr_march<4>), grid, kThreads, 0, c->stream, a); break;
  case 5: klaunch((k_adr_march<5>), grid, kThreads, 0, c->stream, a); break;
  case 6: klaunch((k_adr_march<6>), grid, kThreads, 0, c->stream, a); break;
  default: klaunch((k_adr_march<7>), grid, kThreads, 0, c->stream, a); break;
  }
  LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_adr_rhs(b200_ctx* c, const b200_adr_params* p, int mode,
                            const double* y, double* f)
{
  if (mode < 1 || mode > 7) return fail("b200_adr_rhs: mode must be in 1..7");
  if (y == f) return fail("b200_adr_rhs: f must not alias y");
  if (p->ny >= (int64_t)1 << 30) return fail("b200_adr_rhs: ny too large");
  if (!aligned16(y) || !aligned16(f)) return fail("b200_adr_rhs: operand not 16-byte aligned");
  AdrArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = p->nx; a.ny = p->ny; a.k = adr_consts(*p); a.y = y; a.f = f;
  if (launch_adr(c, a, mode)) return -1;
  ALG_BYTES(2, 2 * p->nx * p->ny);
  return 0;
}

extern "C" int b200_adr_lincomb(b200_ctx* c, const b200_adr_params* p, int mode, const double* y,
                                int nterms, const double* cf, const int* src,
                                const double* const* v, double* z, double* f_out)
{
  if (mode < 1 || mode > 7) return fail("b200_adr_lincomb: mode must be in 1..7");
  if (nterms < 1 || nterms > B200_MAX_TERMS) return fail("b200_adr_lincomb: nterms out of range");
  if (z == y || f_out == y) return fail("b200_adr_lincomb: output aliases the stencil input");
  if (p->ny >= (int64_t)1 << 30) return fail("b200_adr_lincomb: ny too large");
  if (!aligned16(y) || !aligned16(z) || !aligned16(f_out)) return fail("b200_adr_lincomb: operand not 16-byte aligned");
  AdrArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = p->nx; a.ny = p->ny; a.k = adr_consts(*p); a.y = y; a.f = f_out; a.z = z;
  a.t.n = nterms;
  for (int k = 0; k < nterms; k++)
  {
    a.t.c[k] = cf[k]; a.t.src[k] = src[k];
    a.t.v[k] = (src[k] == B200_SRC_VECTOR) ? v[k] : nullptr;
    if (src[k] == B200_SRC_VECTOR && !v[k]) return fail("b200_adr_lincomb: NULL operand");
    if (src[k] == B200_SRC_VECTOR && !aligned16(v[k])) return fail("b200_adr_lincomb: operand not 16-byte aligned");
  }
  if (launch_adr(c, a, mode)) return -1;
  {
    int touches = 2 + (f_out ? 1 : 0);
    for (int k = 0; k < nterms; k++) touches += (src[k] == B200_SRC_VECTOR);
    ALG_BYTES(touches, 2 * p->nx * p->ny);
  }
  return 0;
}

extern "C" int b200_adr_diffusion_lincomb(b200_ctx* c, const b200_adr_params* p, const double* y,
                                          int nterms, const double* cf, const int* src,
                                          const double* const* v, double* z, double* f_out)
{
  return b200_adr_lincomb(c, p, 2, y, nterms, cf, src, v, z, f_out);
}

// ----------------------------------------- adr: temporally blocked diffusion stages
#include "adr_chain.cuh"

template <int K>
static int launch_adr_chain_k(const AdrChainArgs& a, dim3 grid, cudaStream_t st)
{
  const size_t smem = adr_chain_smem(K, kAdrChainPF);
  static bool configured_on[kMaxDevices] = {};
  bool& configured = configured_on[current_device()];
  if (!configured)
  {
    CU_TRY(cudaFuncSetAttribute(k_adr_chain<K, kAdrChainPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  klaunch((k_adr_chain<K, kAdrChainPF>), grid, kAdrChainThreads, smem, st, a);
  return 0;
}

extern "C" int b200_adr_chain(b200_ctx* c, const b200_adr_params* p, int nstages, const double* x,
                              const double* prev2, const double* yn, const double* fn, const double* coeffs,
                              double* const* z_out)
{
  if (!adr_chain_supported(p->nx, p->ny, nstages)) return fail("b200_adr_chain: needs 2 <= nstages <= B200_MAX_CHAIN, nx >= 64, ny >= 16");
  if (p->ny >= (int64_t)1 << 30) return fail("b200_adr_chain: ny too large");
  if (!aligned16(x) || !aligned16(prev2) || !aligned16(yn) || !aligned16(fn)) return fail("b200_adr_chain: operand not 16-byte aligned");
  AdrChainArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = p->nx; a.ny = p->ny; a.k = adr_consts(*p);
  a.x = x; a.prev2 = prev2; a.yn = yn; a.fn = fn;
  int stored = 0;
  for (int l = 0; l < nstages; l++)
  {
    for (int q = 0; q < 5; q++) a.c[l][q] = coeffs[5 * l + q];
    a.out[l] = z_out[l];
    if (a.out[l])
    {
      stored++;
      if (!aligned16(a.out[l])) return fail("b200_adr_chain: output not 16-byte aligned");
      if (a.out[l] == x || a.out[l] == prev2 || a.out[l] == yn || a.out[l] == fn) return fail("b200_adr_chain: an output aliases an input");
    }
  }
  if (!a.out[nstages - 1]) return fail("b200_adr_chain: the last stage must be stored");
  // rows per block: the largest of 128 / 64 / 32 / 16 that leaves >= 2 waves of blocks (2 per SM)
  const int64_t gx    = adr_chain_grid(a.nx, a.ny, nstages, 16).x;
  const int64_t waves = 2 * 2 * (int64_t)(c->sm_count > 0 ? c->sm_count : 148);
  a.rows              = 16;
  for (int r : {128, 64, 32})
    if (gx * ((a.ny + r - 1) / r) >= waves) { a.rows = r; break; }
  a.rows = rows_fit_waves(a.ny, gx, a.rows, waves / 2);
  if ((a.ny + a.rows - 1) / a.rows > 65535) a.rows = (int)((a.ny + 65534) / 65535);
  dim3 grid = adr_chain_grid(a.nx, a.ny, nstages, a.rows);
  int rc    = 0;
  switch (nstages)
  {
  case 2: rc = launch_adr_chain_k<2>(a, grid, c->stream); break;
  case 3: rc = launch_adr_chain_k<3>(a, grid, c->stream); break;
  case 4: rc = launch_adr_chain_k<4>(a, grid, c->stream); break;
  case 5: rc = launch_adr_chain_k<5>(a, grid, c->stream); break;
  default: rc = launch_adr_chain_k<6>(a, grid, c->stream); break;
  }
  if (rc) return rc;
  LAUNCH_CHECK();
  ALG_BYTES(4 + stored, 2 * p->nx * p->ny);
  return 0;
}

// ------------------------------------- adr: implicit reaction (block-diagonal Newton systems)
#include "react_kernels.cuh"

static unsigned blocks_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

extern "C" int b200_adr_jac_reaction(b200_ctx* c, const b200_adr_params* p, const double* y, double* J)
{
  if (!aligned16(y) || !aligned16(J)) return fail("b200_adr_jac_reaction: pointer not 16-byte aligned");
  const int64_t npts = p->nx * p->ny;
  if (npts < 1) return fail("b200_adr_jac_reaction: empty grid");
  const AdrConsts k = adr_consts(*p);
  klaunch(k_adr_jac_reaction, blocks_for(npts), kThreads, 0, c->stream, npts, k.B, k.Bp1, y, J);
  LAUNCH_CHECK();
  ALG_BYTES(3, 2 * npts);
  return 0;
}
extern "C" int b200_blk2_scale_add_i(b200_ctx* c, double cc, double* A, int64_t npts)
{
  if (!aligned16(A) || npts < 1) return fail("b200_blk2_scale_add_i: bad argument");
  klaunch(k_blk2_scale_add_i, blocks_for(npts), kThreads, 0, c->stream, npts, cc, A);
  LAUNCH_CHECK();
  ALG_BYTES(4, 2 * npts);
  return 0;
}
// *info = 0, or the 1-based column of the first zero pivot (what SUNDlsMat_bandGBTRF returns); synchronises
extern "C" int b200_blk2_factor(b200_ctx* c, double* A, double* piv, int64_t npts, long long* info)
{
  if (!aligned16(A) || !piv || npts < 1 || !info) return fail("b200_blk2_factor: bad argument");
  unsigned long long* flag_h = reinterpret_cast<unsigned long long*>(c->host_result + 1);
  unsigned long long* flag_d = reinterpret_cast<unsigned long long*>(c->host_result_dev + 1);
  *reinterpret_cast<volatile unsigned long long*>(flag_h) = ~0ull;
  klaunch(k_blk2_factor, blocks_for(npts), kThreads, 0, c->stream, npts, A, piv, flag_d);
  LAUNCH_CHECK();
  ALG_BYTES(5, 2 * npts);
  CU_TRY(stream_sync_profiled(c->stream));
  const unsigned long long f = *reinterpret_cast<volatile unsigned long long*>(flag_h);
  *info = (f == ~0ull) ? 0 : (long long)f;
  return 0;
}
extern "C" int b200_blk2_solve(b200_ctx* c, const double* A, const double* piv, const double* b, double* x, int64_t npts)
{
  if (!aligned16(A) || !aligned16(b) || !aligned16(x) || !piv || npts < 1) return fail("b200_blk2_solve: bad argument");
  klaunch(k_blk2_solve, blocks_for(npts), kThreads, 0, c->stream, npts, A, piv, b, x);
  LAUNCH_CHECK();
  ALG_BYTES(5, 2 * npts);
  return 0;
}

// --------------------------------------------------------------------- NCCL
// NCCL is resolved at first use with dlopen("libnccl.so.2") instead of being a link-time
// dependency: inside a Python process torch has usually loaded its own bundled NCCL
// already (same SONAME -> the same handle is returned), in the standalone C++ driver
// the system library is used.  Either way one NCCL per process.
struct NcclApi
{
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  bool ok = false;
};
static NcclApi g_nccl;

static int nccl_load()
{
  if (g_nccl.ok) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail("cannot dlopen libnccl.so.2 (needed for multi-GPU runs)");
#define NCCL_SYM(field, name)                                  \
  *(void**)(&g_nccl.field) = dlsym(h, name);                   \
  if (!g_nccl.field) return fail("libnccl: missing symbol " name);
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  NCCL_SYM(CommInitRank, "ncclCommInitRank")
  NCCL_SYM(CommDestroy, "ncclCommDestroy")
  NCCL_SYM(GetErrorString, "ncclGetErrorString")
  NCCL_SYM(AllReduce, "ncclAllReduce")
  NCCL_SYM(Send, "ncclSend")
  NCCL_SYM(Recv, "ncclRecv")
  NCCL_SYM(AllGather, "ncclAllGather")
  NCCL_SYM(GroupStart, "ncclGroupStart")
  NCCL_SYM(GroupEnd, "ncclGroupEnd")
#undef NCCL_SYM
  g_nccl_destroy = g_nccl.CommDestroy;
  g_nccl.ok = true;
  return 0;
}
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclGetErrorString g_nccl.GetErrorString
#define ncclAllReduce g_nccl.AllReduce
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclAllGather g_nccl.AllGather
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
extern "C" int b200_comm_unique_id(unsigned char id[128])
{
  if (nccl_load()) return -1;
  ncclUniqueId u;
  NCCL_TRY(ncclGetUniqueId(&u));
  static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id, &u, 128);
  return 0;
}

extern "C" int b200_comm_init(b200_ctx* c, int rank, int nranks, const unsigned char id[128])
{
  if (nccl_load()) return -1;
  CU_TRY(cudaSetDevice(c->device));
  ncclUniqueId u;
  memcpy(&u, id, 128);
  NCCL_TRY(ncclCommInitRank(&c->comm, nranks, u, rank));
  c->rank   = rank;
  c->nranks = nranks;
  return 0;
}

extern "C" int b200_comm_rank(b200_ctx* c, int* rank, int* nranks)
{
  *rank   = c->rank;
  *nranks = c->nranks;
  return 0;
}

static int nccl_allreduce_inplace(b200_ctx* c, double* buf, int n, int op)
{
  ncclRedOp_t o = (op == RED_SUM) ? ncclSum : (op == RED_MAX) ? ncclMax : ncclMin;
  NCCL_TRY(ncclAllReduce(buf, buf, (size_t)n, ncclDouble, o, c->comm, c->stream));
  return 0;
}

extern "C" int b200_allreduce(b200_ctx* c, double* dev_buf, int n, int op)
{
  if (!c->comm || c->nranks == 1) return 0;
  return nccl_allreduce_inplace(c, dev_buf, n, op);
}

// one direction pair of the deep halo exchange: "lo" data goes to peers[0] (W or S) and arrives
// as that rank's "hi" halo; "hi" data goes to peers[1]; per-peer posting order is the same on
// every rank (send lo, recv hi, send hi, recv lo), which is what NCCL's in-group matching needs.
static int deep_halo_nccl_phase(b200_ctx* c, int nfields, const int peers[2], const double* const* send_lo,
                                const double* const* send_hi, double* const* recv_lo, double* const* recv_hi,
                                size_t count)
{
  if (nccl_load()) return -1;
  NCCL_TRY(ncclGroupStart());
  for (int f = 0; f < nfields; f++)
  {
    NCCL_TRY(ncclSend(send_lo[f], count, ncclDouble, peers[0], c->comm, c->stream));
    NCCL_TRY(ncclRecv(recv_hi[f], count, ncclDouble, peers[1], c->comm, c->stream));
    NCCL_TRY(ncclSend(send_hi[f], count, ncclDouble, peers[1], c->comm, c->stream));
    NCCL_TRY(ncclRecv(recv_lo[f], count, ncclDouble, peers[0], c->comm, c->stream));
  }
  NCCL_TRY(ncclGroupEnd());
  return 0;
}

extern "C" int b200_halo_exchange(b200_ctx* c, const int peers[4], const double* sw,
                                  const double* se, const double* ss, const double* sn,
                                  double* rw, double* re, double* rs, double* rn,
                                  int64_t nx, int64_t ny)
{
  if (!c->comm) return fail("b200_halo_exchange: communicator not initialised");
  // comm stream picks up after everything enqueued so far on the compute stream
  CU_TRY(cudaEventRecord(c->ev_compute, c->stream));
  CU_TRY(cudaStreamWaitEvent(c->comm_stream, c->ev_compute, 0));
  // The reference pairs messages by tag (diffusion_2D.cpp:421-503): what I send west
  // is my west neighbour's "from east" halo, and so on.  Within one NCCL group the
  // per-peer send/recv order must match on both sides: every rank posts
  // W-send, E-recv, E-send, W-recv, S-send, N-recv, N-send, S-recv.
  NCCL_TRY(ncclGroupStart());
  if (sw) NCCL_TRY(ncclSend(sw, (size_t)ny, ncclDouble, peers[0], c->comm, c->comm_stream));
  if (re) NCCL_TRY(ncclRecv(re, (size_t)ny, ncclDouble, peers[1], c->comm, c->comm_stream));
  if (se) NCCL_TRY(ncclSend(se, (size_t)ny, ncclDouble, peers[1], c->comm, c->comm_stream));
  if (rw) NCCL_TRY(ncclRecv(rw, (size_t)ny, ncclDouble, peers[0], c->comm, c->comm_stream));
  if (ss) NCCL_TRY(ncclSend(ss, (size_t)nx, ncclDouble, peers[2], c->comm, c->comm_stream));
  if (rn) NCCL_TRY(ncclRecv(rn, (size_t)nx, ncclDouble, peers[3], c->comm, c->comm_stream));
  if (sn) NCCL_TRY(ncclSend(sn, (size_t)nx, ncclDouble, peers[3], c->comm, c->comm_stream));
  if (rs) NCCL_TRY(ncclRecv(rs, (size_t)nx, ncclDouble, peers[2], c->comm, c->comm_stream));
  NCCL_TRY(ncclGroupEnd());
  CU_TRY(cudaEventRecord(c->ev_comm, c->comm_stream));
  c->comm_pending = true;
  return 0;
}

extern "C" int b200_halo_wait(b200_ctx* c)
{
  if (c->comm_pending)
  {
    CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
    c->comm_pending = false;
  }
  return 0;
}

// ------------------------------------------------------- peer-mapped deep-halo exchange
// The deep halos of temporally blocked launches without NCCL: a ring of halo slots per rank, mapped into the
// neighbours' address spaces (CUDA IPC, one process per GPU), filled by the NEIGHBOURS' k_peer_exchange kernels with
// plain stores over NVLink.  Slot numbers are chosen by a deterministic allocator that every rank runs in lockstep
// (all ranks issue the same sequence of vector operations), so "slot s" names the same logical halo everywhere.
struct b200_peer_halo
{
  b200_ctx* ctx = nullptr;
  int64_t nx = 0, ny = 0, ny_s = 0, ny_n = 0; // my block, and the heights of the blocks south / north of me
  int g = 0, g2 = 0, nslots = 0;
  int64_t slot_doubles = 0;      // doubles per slot, sized for the tallest block of the decomposition
  int64_t ny_max = 0;
  double* base = nullptr;        // [nslots * slot_doubles doubles | 8 arrival counters | ticket]
  size_t bytes = 0;
  double* nbr_base[8] = {};      // the same region of each neighbour, in my address space
  void* opened[8] = {};          // IPC mappings to close (distinct, non-self neighbours)
  int nbr_rank[8] = {};
  unsigned long long epoch = 0;  // exchanges enqueued so far
  std::vector<unsigned long long> avail; // slot s may be handed out for an exchange with epoch >= avail[s]; ~0 = in use
  int* host_err = nullptr;       // mapped
  int* host_err_dev = nullptr;
  uint64_t exchanges = 0, doubles_pushed = 0;
};

static unsigned long long* ph_flags(const b200_peer_halo* ph, double* base)
{
  return reinterpret_cast<unsigned long long*>(base + (size_t)ph->nslots * ph->slot_doubles);
}

extern "C" int b200_peer_halo_destroy(b200_peer_halo* ph)
{
  if (!ph) return 0;
  cudaSetDevice(ph->ctx->device);
  cudaStreamSynchronize(ph->ctx->stream);
#ifndef B200_HOST_EMU
  for (int d = 0; d < 8; d++)
    if (ph->opened[d]) cudaIpcCloseMemHandle(ph->opened[d]);
#endif
  if (ph->base) cudaFree(ph->base);
  if (ph->host_err) cudaFreeHost(ph->host_err);
  delete ph;
  return 0;
}

static int peer_halo_init(b200_peer_halo* ph, const int nbr[8])
{
  b200_ctx* c = ph->ctx;
  CU_TRY(cudaSetDevice(c->device));
  ph->slot_doubles = (b200_deep_halo_doubles(ph->nx, ph->ny_max, ph->g, ph->g2) + 31) & ~(int64_t)31;
  ph->bytes        = sizeof(double) * (size_t)ph->nslots * ph->slot_doubles + 256;
  CU_TRY(cudaMalloc(&ph->base, ph->bytes));
  CU_TRY(cudaMemset(ph->base, 0, ph->bytes));
  CU_TRY(cudaHostAlloc(&ph->host_err, sizeof(int) * 4, cudaHostAllocMapped));
  ph->host_err[0] = 0;
  CU_TRY(cudaHostGetDevicePointer(&ph->host_err_dev, ph->host_err, 0));
  CU_TRY(cudaStreamSynchronize(c->stream));
  ph->avail.assign((size_t)ph->nslots, 0ull);
  bool remote = false;
  for (int d = 0; d < 8; d++)
  {
    ph->nbr_rank[d] = nbr[d];
    if (nbr[d] == c->rank || c->nranks <= 1) ph->nbr_base[d] = ph->base;
    else remote = true;
  }
  if (!remote) return 0;
#ifdef B200_HOST_EMU
  return fail("b200_peer_halo_create: the emulated build is single-rank");
#else
  if (!c->comm) return fail("b200_peer_halo_create: communicator not initialised");
  // every rank publishes the IPC handle of its region; all-gather over the communicator
  cudaIpcMemHandle_t mine;
  CU_TRY(cudaIpcGetMemHandle(&mine, ph->base));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  char *dsend = nullptr, *drecv = nullptr;
  CU_TRY(cudaMalloc(&dsend, 64));
  CU_TRY(cudaMalloc(&drecv, 64 * (size_t)c->nranks));
  CU_TRY(cudaMemcpyAsync(dsend, &mine, 64, cudaMemcpyHostToDevice, c->stream));
  NCCL_TRY(ncclAllGather(dsend, drecv, 64, ncclChar, c->comm, c->stream));
  std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
  CU_TRY(cudaMemcpyAsync(all.data(), drecv, 64 * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  CU_TRY(cudaFree(dsend));
  CU_TRY(cudaFree(drecv));
  for (int d = 0; d < 8; d++)
  {
    if (ph->nbr_base[d]) continue;
    for (int e = 0; e < d; e++) // the same rank in two directions (2 x 1, 2 x 2 layouts): map it once
      if (nbr[e] == nbr[d] && ph->nbr_base[e]) { ph->nbr_base[d] = ph->nbr_base[e]; break; }
    if (ph->nbr_base[d]) continue;
    void* p = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&p, all[(size_t)nbr[d]], cudaIpcMemLazyEnablePeerAccess));
    ph->opened[d]   = p;
    ph->nbr_base[d] = static_cast<double*>(p);
  }
  // nobody may push before everybody has mapped (and zeroed) everything: one barrier
  NCCL_TRY(ncclAllReduce(c->dev_result, c->dev_result, 1, ncclDouble, ncclSum, c->comm, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return 0;
#endif
}

extern "C" int b200_peer_halo_create(b200_ctx* c, const int nbr[8], int64_t nx, int64_t ny, int64_t ny_south,
                                     int64_t ny_north, int64_t ny_max, int g, int g2, int nslots, b200_peer_halo** out)
{
  if (!c || !out || nslots < 8 || g < 1 || g2 < 2 || (g2 & 1) || g > ny || g2 > nx) return fail("b200_peer_halo_create: bad argument");
  b200_peer_halo* ph = new b200_peer_halo();
  ph->ctx = c; ph->nx = nx; ph->ny = ny; ph->ny_s = ny_south; ph->ny_n = ny_north; ph->ny_max = ny_max;
  ph->g = g; ph->g2 = g2; ph->nslots = nslots;
  int rc = peer_halo_init(ph, nbr);
  if (rc) { b200_peer_halo_destroy(ph); return rc; }
  *out = ph;
  return 0;
}

// Slot allocator.  A slot released while `epoch` exchanges have been enqueued may still be read by this rank's
// launch that follows exchange `epoch`; a neighbour can write into it from exchange epoch + 2 on at the earliest
// (it starts that exchange only after it has seen this rank's flag of exchange epoch + 1, raised after that launch).
extern "C" double* b200_peer_halo_slot_alloc(b200_peer_halo* ph)
{
  const unsigned long long next = ph->epoch + 1;
  for (int s = 0; s < ph->nslots; s++)
    if (ph->avail[(size_t)s] <= next)
    {
      ph->avail[(size_t)s] = ~0ull;
      return ph->base + (size_t)s * ph->slot_doubles;
    }
  fail("b200_peer_halo_slot_alloc: all halo slots are in use");
  return nullptr;
}

extern "C" int b200_peer_halo_slot_free(b200_peer_halo* ph, double* slot)
{
  const int64_t s = (slot - ph->base) / ph->slot_doubles;
  if (s < 0 || s >= ph->nslots || ph->avail[(size_t)s] != ~0ull) return fail("b200_peer_halo_slot_free: not an allocated slot");
  ph->avail[(size_t)s] = ph->epoch + 2;
  return 0;
}

extern "C" int b200_peer_halo_stats(const b200_peer_halo* ph, uint64_t* exchanges, uint64_t* doubles_pushed)
{
  *exchanges      = ph->exchanges;
  *doubles_pushed = ph->doubles_pushed;
  return 0;
}

extern "C" int b200_peer_halo_exchange(b200_peer_halo* ph, int nfields, const double* const* fields, double* const* slots)
{
  if (nfields < 1 || nfields > 4) return fail("b200_peer_halo_exchange: 1..4 fields");
  b200_ctx* c = ph->ctx;
  if (ph->host_err[0]) return fail("b200_peer_halo_exchange: a neighbour did not arrive in an earlier exchange (time limit)");
  PeerXArgs a;
  memset(&a, 0, sizeof(a));
  a.nf = nfields; a.nx = ph->nx; a.ny = ph->ny; a.g = ph->g; a.g2 = ph->g2;
  a.epoch = ++ph->epoch;
  for (int f = 0; f < nfields; f++)
  {
    const int64_t off = slots[f] - ph->base; // the same offset in every rank's region (lockstep allocator)
    if (off < 0 || off % ph->slot_doubles || off / ph->slot_doubles >= ph->nslots) return fail("b200_peer_halo_exchange: not a slot");
    if (!aligned16(fields[f])) return fail("b200_peer_halo_exchange: field not 16-byte aligned");
    a.field[f] = fields[f];
    double* nbr_slot[8];
    for (int d = 0; d < 8; d++) nbr_slot[d] = ph->nbr_base[d] + off;
    peer_dst_pointers(nbr_slot, ph->nx, ph->ny, ph->ny_s, ph->ny_n, ph->g, ph->g2, a.dst[f]);
  }
  for (int d = 0; d < 8; d++) a.peer_flag[d] = ph_flags(ph, ph->nbr_base[d]) + kPeerOpposite[d];
  a.my_flag    = ph_flags(ph, ph->base);
  a.ticket     = reinterpret_cast<unsigned*>(ph_flags(ph, ph->base) + 8);
  a.err        = ph->host_err_dev;
  a.timeout_ns = 20ull * 1000000000ull;
  const int64_t total = 2 * (int64_t)ph->g * ph->nx + 2 * ph->ny * ph->g2 + 4 * (int64_t)ph->g * ph->g2;
  int64_t blocks      = (total + kThreads - 1) / kThreads;
  if (blocks > (int64_t)c->sm_count / 2) blocks = (int64_t)c->sm_count / 2; // a few MB: latency-, not bandwidth-bound
  if (blocks < 1) blocks = 1;
  klaunch(k_peer_exchange, dim3((unsigned)blocks, (unsigned)nfields), kThreads, 0, c->stream, a);
  LAUNCH_CHECK();
  ph->exchanges++;
  ph->doubles_pushed += (uint64_t)total * (uint64_t)nfields;
  return 0;
}
