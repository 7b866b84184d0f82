// cuda_runtime_emu.h -- TEST INFRASTRUCTURE ONLY: the slice of the CUDA runtime (and the NCCL types) that
// csrc/b200_kernels.cu uses, restated for the host so that the WHOLE library -- contexts, launchers, C-ABI --
// compiles with g++ on top of the execution-model emulation in cuda_emu.h.  "Device" memory is host memory,
// a stream is synchronous, an event is a time stamp.  With it the unchanged host stack (ARKODE, N_Vector_B200's
// lazy fusion, the problem layers) runs end to end on a CPU in the `-m "not gpu"` suite and is compared with the
// reference's fixtures bit for bit; see tests/emu/Makefile (libb200_fullstack_emu.so) and
// tests/test_fullstack_emulation.py.
//
// Nothing in the product includes this file: b200_kernels.cu pulls it in only under B200_HOST_EMU, which only
// tests/emu/Makefile defines, and the emulated library has its own name and is never on the product's load path.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "cuda_emu.h"

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct emu_stream_s* cudaStream_t;
struct emu_event_s { double t; };
typedef emu_event_s* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
// a small "GPU": grid-size heuristics that look at the SM count stay cheap to emulate
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(new char); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete reinterpret_cast<char*>(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emu_event_s{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
  e->t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return cudaSuccess;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t emu_alloc(void** p, size_t bytes)
{ // cudaMalloc alignment; NaN-poisoned so that a read of never-written "device" memory shows up
  const size_t n = (bytes + 255) & ~(size_t)255;
  *p             = aligned_alloc(256, n ? n : 256);
  if (!*p) return cudaErrorMemoryAllocation;
  memset(*p, 0xff, n ? n : 256);
  return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return emu_alloc((void**)p, bytes); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return emu_alloc((void**)p, bytes); }
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned) { return emu_alloc((void**)p, bytes); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostGetDevicePointer(T** d, void* h, unsigned) { *d = (T*)h; return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
struct cudaFuncAttributes { int numRegs; };
template <class F> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { a->numRegs = 0; return cudaSuccess; }

// ---- NCCL: types only.  The library binds NCCL with dlopen at first use (nccl_load), so a multi-rank call in the
// emulated build fails there with a clear message; the single-rank paths never touch it.
typedef struct emu_nccl_comm* ncclComm_t;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef enum { ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
typedef struct { char internal[128]; } ncclUniqueId;
