// kernel_prims.cuh -- the handful of device primitives the stage kernels are written in:
// exactly-rounded FP64 multiply / add (no contraction), the optional FMA form, 16-byte loads,
// cp.async staging and dynamic shared memory.  With B200_HOST_EMU defined (tests/emu only) the same
// names are provided by tests/emu/cuda_emu.h so that the kernel sources run lane by lane on CPU
// threads; the product never defines B200_HOST_EMU.
#pragma once

#ifdef B200_HOST_EMU
#include "cuda_emu.h"
#else

#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DFMA(a, b, c) __fma_rn((a), (b), (c))
#define B200_DYN_SMEM(type, name) extern __shared__ type name[]

// streaming (read-once) loads: keep them out of L1, evict-first in L2
__device__ __forceinline__ double2 ld_stream2(const double* p)
{
  double2 r;
  asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_keep2(const double* p)
{
  return *reinterpret_cast<const double2*>(p);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}


// system-scope flag traffic of the peer-mapped halo exchange (flags live in another GPU's memory, reached over NVLink)
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_sys() { __threadfence_system(); }
__device__ __forceinline__ unsigned long long global_timer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void nap_ns(unsigned ns) { __nanosleep(ns); }

#endif // B200_HOST_EMU

// c + a*b: two roundings (the reference's baseline x86-64 build) or one (FMA-contracted arithmetic,
// what gcc's default -ffp-contract=fast makes of the same source on an FMA-baseline ISA)
template <bool FMA>
__device__ __forceinline__ double mad(double a, double b, double c)
{
  return FMA ? DFMA(a, b, c) : DADD(c, DMUL(a, b));
}
