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

// Bulk asynchronous copies (the TMA unit's 1-D path, SASS UBLKCP) completed on a shared-memory mbarrier: ONE thread
// moves a whole contiguous run of a row (16-byte aligned, a multiple of 16 bytes) global -> shared, the bytes are
// counted down on the barrier's transaction count, and every consumer waits for the barrier's phase by parity.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned arrivals)
{
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ba), "r"(arrivals) : "memory");
}
// makes freshly initialised barriers visible to the async proxy (call once after the mbar_init's, before the first copy)
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one arrival that also announces `bytes` of bulk-copy traffic for the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa),
               "l"(gmem), "r"(bytes), "r"(ba)
               : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one()
{
  unsigned pred;
  asm volatile("{\n"
               ".reg .pred p;\n"
               "elect.sync _|p, 0xffffffff;\n"
               "selp.u32 %0, 1, 0, p;\n"
               "}"
               : "=r"(pred));
  return pred != 0;
}
// a value that is the same in every lane, in a form the compiler can keep in a uniform register
__device__ __forceinline__ int warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }
// blocks until the phase of parity `parity` has completed (try_wait suspends the warp for a while instead of spinning)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n"
               ".reg .pred p;\n"
               "MBAR_WAIT:\n" // (labels are local to the { } block)
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra MBAR_DONE;\n"
               "bra MBAR_WAIT;\n"
               "MBAR_DONE:\n"
               "}" ::"r"(ba),
               "r"(parity)
               : "memory");
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
