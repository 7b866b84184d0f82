// reduce_prims.cuh -- term lists, block size and the deterministic warp / block / grid reduction used by
// the reduction kernels and by the fused WRMS norm of the stage kernel.  Included by b200_kernels.cu
// (nvcc) and, under B200_HOST_EMU, by the host emulation harness tests/emu.
#pragma once
#include "kernel_prims.cuh"

struct LinTerms
{
  int n;
  int src[B200_MAX_TERMS];
  double c[B200_MAX_TERMS];
  const double* v[B200_MAX_TERMS];
};

static const int kThreads = 256;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = DADD(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

enum RedOp { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };

template <int ROP>
__device__ __forceinline__ double red_combine(double a, double b)
{
  if (ROP == RED_SUM) return DADD(a, b);
  if (ROP == RED_MAX) return fmax(a, b);
  return fmin(a, b);
}
template <int ROP>
__device__ __forceinline__ double red_identity()
{
  if (ROP == RED_SUM) return 0.0;
  if (ROP == RED_MAX) return 0.0; // max-norm of |x| >= 0
  return __longlong_as_double(0x7ff0000000000000LL);
}

// Block-level reduce (fixed shuffle tree -> deterministic), result valid in thread 0.
template <int ROP>
__device__ __forceinline__ double block_reduce(double v, double* smem /* >= 32 */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x * blockDim.y + 31) >> 5;
  if (ROP == RED_SUM) v = warp_sum(v);
  else if (ROP == RED_MAX) v = warp_max(v);
  else v = warp_min(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0)
  {
    v = (lane < nwarp) ? smem[lane] : red_identity<ROP>();
    if (ROP == RED_SUM) v = warp_sum(v);
    else if (ROP == RED_MAX) v = warp_max(v);
    else v = warp_min(v);
  }
  return v;
}

// Grid-level finish: every block stores its partial; the block that takes the
// last ticket re-reduces all partials in index order (so the result does not
// depend on which block happens to be last) and resets the ticket.
template <int ROP>
__device__ __forceinline__ void grid_finish(double block_val, unsigned nblocks,
                                            unsigned bid, double* partials,
                                            unsigned* ticket, double* result,
                                            double* smem)
{
  __shared__ bool is_last;
  if (threadIdx.x == 0)
  {
    partials[bid] = block_val;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last    = (t == nblocks - 1);
  }
  __syncthreads();
  if (is_last)
  {
    __threadfence();
    double acc = red_identity<ROP>();
    for (unsigned k = threadIdx.x; k < nblocks; k += blockDim.x)
      acc = red_combine<ROP>(acc, ((volatile double*)partials)[k]);
    acc = block_reduce<ROP>(acc, smem);
    if (threadIdx.x == 0)
    {
      *result = acc;
      *ticket = 0;
    }
  }
}


// Two sums finished together (one ticket): partials[bid] / partials[nblocks + bid] -> result[0] / result2[0].
__device__ __forceinline__ void grid_finish2(double v0, double v1, unsigned nblocks, unsigned bid, double* partials,
                                             unsigned* ticket, double* result, double* result2, double* smem)
{
  __shared__ bool is_last2;
  if (threadIdx.x == 0)
  {
    partials[bid]           = v0;
    partials[nblocks + bid] = v1;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last2   = (t == nblocks - 1);
  }
  __syncthreads();
  if (is_last2)
  {
    __threadfence();
    double a0 = 0.0, a1 = 0.0;
    for (unsigned k = threadIdx.x; k < nblocks; k += blockDim.x)
    {
      a0 = DADD(a0, ((volatile double*)partials)[k]);
      a1 = DADD(a1, ((volatile double*)partials)[nblocks + k]);
    }
    a0 = block_reduce<RED_SUM>(a0, smem);
    a1 = block_reduce<RED_SUM>(a1, smem);
    if (threadIdx.x == 0)
    {
      *result  = a0;
      *result2 = a1;
      *ticket  = 0;
    }
  }
}
