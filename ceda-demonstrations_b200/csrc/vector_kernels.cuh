// vector_kernels.cuh -- the elementwise N_Vector kernels (k_elementwise<OP>) and the deterministic
// reductions (k_reduce<KIND, ROP>).  Included by b200_kernels.cu (nvcc) and, under B200_HOST_EMU, by the
// host emulation harness tests/emu.
#pragma once
#include "reduce_prims.cuh"

enum EwOp
{
  EW_LINCOMB = 0,
  EW_SCALESUM,
  EW_SCALEDIFF,
  EW_CONST,
  EW_PROD,
  EW_DIV,
  EW_ABS,
  EW_INV,
  EW_ADDCONST,
  EW_EWT
};

struct EwArgs
{
  LinTerms t;   // LINCOMB
  const double* x;
  const double* y;
  double a, b;
  double* z;
  int64_t n;
};

template <int OP>
__device__ __forceinline__ double ew_apply(const EwArgs& a, int64_t i)
{
  if (OP == EW_LINCOMB)
  {
    double acc = DMUL(a.t.c[0], a.t.v[0][i]);
#pragma unroll
    for (int k = 1; k < B200_MAX_TERMS; k++)
      if (k < a.t.n) acc = DADD(acc, DMUL(a.t.c[k], a.t.v[k][i]));
    return acc;
  }
  if (OP == EW_SCALESUM) return DMUL(a.a, DADD(a.x[i], a.y[i]));
  if (OP == EW_SCALEDIFF) return DMUL(a.a, DSUB(a.x[i], a.y[i]));
  if (OP == EW_CONST) return a.a;
  if (OP == EW_PROD) return DMUL(a.x[i], a.y[i]);
  if (OP == EW_DIV) return __ddiv_rn(a.x[i], a.y[i]);
  if (OP == EW_ABS) return fabs(a.x[i]);
  if (OP == EW_INV) return __ddiv_rn(1.0, a.x[i]);
  if (OP == EW_ADDCONST) return DADD(a.x[i], a.b);
  // EW_EWT: N_VAbs, N_VScale(rtol), N_VAddConst(atol), N_VInv
  return __ddiv_rn(1.0, DADD(DMUL(a.a, fabs(a.x[i])), a.b));
}

template <int OP>
__device__ __forceinline__ double2 ew_apply2(const EwArgs& a, int64_t i)
{
  double2 r;
  if (OP == EW_LINCOMB)
  {
    double2 v = ld_keep2(a.t.v[0] + i);
    r.x = DMUL(a.t.c[0], v.x);
    r.y = DMUL(a.t.c[0], v.y);
#pragma unroll
    for (int k = 1; k < B200_MAX_TERMS; k++)
      if (k < a.t.n)
      {
        v   = ld_keep2(a.t.v[k] + i);
        r.x = DADD(r.x, DMUL(a.t.c[k], v.x));
        r.y = DADD(r.y, DMUL(a.t.c[k], v.y));
      }
    return r;
  }
  if (OP == EW_CONST) { r.x = a.a; r.y = a.a; return r; }
  double2 x = ld_keep2(a.x + i);
  if (OP == EW_SCALESUM || OP == EW_SCALEDIFF || OP == EW_PROD || OP == EW_DIV)
  {
    double2 y = ld_keep2(a.y + i);
    if (OP == EW_SCALESUM) { r.x = DMUL(a.a, DADD(x.x, y.x)); r.y = DMUL(a.a, DADD(x.y, y.y)); }
    if (OP == EW_SCALEDIFF) { r.x = DMUL(a.a, DSUB(x.x, y.x)); r.y = DMUL(a.a, DSUB(x.y, y.y)); }
    if (OP == EW_PROD) { r.x = DMUL(x.x, y.x); r.y = DMUL(x.y, y.y); }
    if (OP == EW_DIV) { r.x = __ddiv_rn(x.x, y.x); r.y = __ddiv_rn(x.y, y.y); }
    return r;
  }
  if (OP == EW_ABS) { r.x = fabs(x.x); r.y = fabs(x.y); }
  if (OP == EW_INV) { r.x = __ddiv_rn(1.0, x.x); r.y = __ddiv_rn(1.0, x.y); }
  if (OP == EW_ADDCONST) { r.x = DADD(x.x, a.b); r.y = DADD(x.y, a.b); }
  if (OP == EW_EWT)
  {
    r.x = __ddiv_rn(1.0, DADD(DMUL(a.a, fabs(x.x)), a.b));
    r.y = __ddiv_rn(1.0, DADD(DMUL(a.a, fabs(x.y)), a.b));
  }
  return r;
}

// grid-stride, 2 x double2 per thread per trip (4 independent 16-byte loads / vector)
template <int OP>
__global__ void __launch_bounds__(kThreads) k_elementwise(const EwArgs a)
{
  const int64_t n2     = a.n >> 1; // number of double2
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t p            = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; p + stride < n2; p += 2 * stride)
  {
    double2 r0 = ew_apply2<OP>(a, 2 * p);
    double2 r1 = ew_apply2<OP>(a, 2 * (p + stride));
    *reinterpret_cast<double2*>(a.z + 2 * p)            = r0;
    *reinterpret_cast<double2*>(a.z + 2 * (p + stride)) = r1;
  }
  if (p < n2) { *reinterpret_cast<double2*>(a.z + 2 * p) = ew_apply2<OP>(a, 2 * p); }
  if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) a.z[a.n - 1] = ew_apply<OP>(a, a.n - 1);
}


// RD_WSQRC: sum (x_i * w)^2 with ONE weight for every entry (a constant-valued weight vector is never stored;
// fixed-step LSRKStep sets ewt = N_VConst(SUN_SMALL_REAL) every step, arkode.c:2985-2990)
enum RdKind { RD_DOT = 0, RD_WSQR, RD_MAXNORM, RD_MIN, RD_L1, RD_WSQRC };

template <int KIND>
__device__ __forceinline__ double rd_term(double x, double y)
{
  if (KIND == RD_DOT) return DMUL(x, y);
  if (KIND == RD_WSQR || KIND == RD_WSQRC) { double p = DMUL(x, y); return DMUL(p, p); }
  if (KIND == RD_MAXNORM) return fabs(x);
  if (KIND == RD_MIN) return x;
  return fabs(x);
}

template <int KIND, int ROP>
__global__ void __launch_bounds__(kThreads)
  k_reduce(const double* __restrict__ x, const double* __restrict__ y, double ys, int64_t n,
           double* partials, unsigned* ticket, double* result)
{
  __shared__ double smem[32];
  const bool two       = (KIND == RD_DOT || KIND == RD_WSQR);
  const int64_t n2     = n >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc0 = red_identity<ROP>(), acc1 = red_identity<ROP>();
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n2; p += stride)
  {
    double2 a = ld_keep2(x + 2 * p);
    double2 b = two ? ld_keep2(y + 2 * p) : make_double2(ys, ys);
    acc0      = red_combine<ROP>(acc0, rd_term<KIND>(a.x, b.x));
    acc1      = red_combine<ROP>(acc1, rd_term<KIND>(a.y, b.y));
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
    acc0 = red_combine<ROP>(acc0, rd_term<KIND>(x[n - 1], two ? y[n - 1] : ys));
  double v = block_reduce<ROP>(red_combine<ROP>(acc0, acc1), smem);
  grid_finish<ROP>(v, gridDim.x, blockIdx.x, partials, ticket, result, smem);
}

