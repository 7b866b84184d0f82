// adr_kernels.cuh -- the adr 2-D Brusselator kernels (k_adr_march<MODE>) and their constants.
// Included by b200_kernels.cu (nvcc) and, under B200_HOST_EMU, by the host emulation harness tests/emu.
#pragma once
#include "reduce_prims.cuh"

// Scalar factors of the three operators, computed ONCE on the host with the reference's own
// expressions (IEEE division, so the bits are those of the reference's per-call scalars):
//   advection  ...2d.cpp:1417-1420: c = ONE*cu/(TWO*dx)
//   diffusion  ...2d.cpp:1461-1462: d*dxinv2 with dxinv2 = ONE/(dx*dx)
//   reaction   ...2d.cpp:1515:      (B + 1)
struct AdrConsts
{
  double cux, cuy, cvx, cvy;
  double kx, ky;
  double A, B, Bp1;
};

static AdrConsts adr_consts(const b200_adr_params& p)
{
  AdrConsts k;
  k.cux = (1.0 * p.cux) / (2.0 * p.dx);
  k.cuy = (1.0 * p.cuy) / (2.0 * p.dy);
  k.cvx = (1.0 * p.cvx) / (2.0 * p.dx);
  k.cvy = (1.0 * p.cvy) / (2.0 * p.dy);
  k.kx  = p.d * (1.0 / (p.dx * p.dx));
  k.ky  = p.d * (1.0 / (p.dy * p.dy));
  k.A   = p.A;
  k.B   = p.B;
  k.Bp1 = p.B + 1.0;
  return k;
}

struct AdrArgs
{
  int64_t nx, ny;
  AdrConsts k;
  const double* y;
  double* f;  // plain RHS output (b200_adr_rhs) or f_out
  LinTerms t; // fused combination (b200_adr_lincomb); t.n == 0: plain RHS
  double* z;
  int rows;   // rows marched per block
};

// One grid point, both species: c = centre, l/r = west/east, b/t = south/north.  Composite callbacks add in
// the order advection, diffusion, reaction (f_adv_react ...2d.cpp:1602-1619, f_adv_diff_react :1622-1646,
// f_diff_react, f_adv_diff; the N_VLinearSum(1,f,1,temp,f) there is Vaxpy: f += temp).
template <int MODE>
__device__ __forceinline__ double2 adr_point(const AdrConsts& k, double2 c, double2 l, double2 r, double2 b, double2 t)
{
  double2 res = make_double2(0, 0);
  if (MODE & 1)
  { // ...2d.cpp:1440-1441: f = cx*(r-l) + cy*(t-b)
    res.x = DADD(DMUL(k.cux, DSUB(r.x, l.x)), DMUL(k.cuy, DSUB(t.x, b.x)));
    res.y = DADD(DMUL(k.cvx, DSUB(r.y, l.y)), DMUL(k.cvy, DSUB(t.y, b.y)));
  }
  if (MODE & 2)
  { // ...2d.cpp:1483-1486: d*dxinv2*(l + r - 2c) + d*dyinv2*(b + t - 2c)
    const double c2x = DMUL(2.0, c.x), c2y = DMUL(2.0, c.y);
    double2 fd;
    fd.x = DADD(DMUL(k.kx, DSUB(DADD(l.x, r.x), c2x)), DMUL(k.ky, DSUB(DADD(b.x, t.x), c2x)));
    fd.y = DADD(DMUL(k.kx, DSUB(DADD(l.y, r.y), c2y)), DMUL(k.ky, DSUB(DADD(b.y, t.y), c2y)));
    res  = (MODE & 1) ? make_double2(DADD(res.x, fd.x), DADD(res.y, fd.y)) : fd;
  }
  if (MODE & 4)
  { // ...2d.cpp:1515-1516: A + u*u*v - (B+1)*u ; B*u - u*u*v
    const double uuv = DMUL(DMUL(c.x, c.x), c.y);
    double2 fr;
    fr.x = DSUB(DADD(k.A, uuv), DMUL(k.Bp1, c.x));
    fr.y = DSUB(DMUL(k.B, c.x), uuv);
    res  = (MODE & 3) ? make_double2(DADD(res.x, fr.x), DADD(res.y, fr.y)) : fr;
  }
  return res;
}

// Marching kernel: a thread owns one grid point (both species = one 16-byte access) of a 256-point strip and
// marches down `rows` rows with the three live rows of y in registers, so a row of y is fetched once per
// block; west/east neighbours come from warp shuffles (lanes 0 / 31 and the strip ends issue one extra
// load; the domain is periodic, the reference wraps indices the same way, ...2d.cpp:1425-1433).  With
// t.n > 0 the operator value is consumed in registers by z = sum_k c[k]*T_k, T_k in {vector, y, F(y)}
// (left to right like SUNDIALS' N_VLinearCombination fallback) and optionally stored as well.
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_adr_march(const AdrArgs a)
{
  constexpr bool NB = (MODE & 3) != 0; // the reaction alone is pointwise
  const int64_t nx  = a.nx;
  const int ny      = (int)a.ny;
  const int lane    = threadIdx.x & 31;
  const int64_t i0  = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const bool active = i0 < nx;
  const int64_t i   = active ? i0 : 0;
  const int j0      = (int)blockIdx.y * a.rows;
  int j1            = j0 + a.rows;
  if (j1 > ny) j1 = ny;
  const bool wload = NB && active && (lane == 0 || i == 0);
  const bool eload = NB && active && (lane == 31 || i == nx - 1);
  const int64_t iw = (i > 0) ? i - 1 : nx - 1, ie = (i < nx - 1) ? i + 1 : 0;
  const double* yb = a.y;
  int64_t off      = 2 * ((int64_t)j0 * nx + i);
  double2 ym = make_double2(0, 0), yc = make_double2(0, 0);
  if (active && j0 < j1)
  {
    yc = ld_keep2(yb + off);
    if (NB) ym = ld_keep2(yb + 2 * ((int64_t)(j0 > 0 ? j0 - 1 : ny - 1) * nx + i));
  }
  const int nt = a.t.n;
#pragma unroll 1
  for (int j = j0; j < j1; j++)
  {
    double2 yp = make_double2(0, 0), wv = make_double2(0, 0), ev = make_double2(0, 0);
    double2 tv[B200_MAX_TERMS];
    if (active)
    {
      if (NB) yp = ld_keep2(yb + 2 * ((int64_t)(j < ny - 1 ? j + 1 : 0) * nx + i));
      else if (j + 1 < j1) yp = ld_keep2(yb + off + 2 * nx);
      if (wload) wv = ld_keep2(yb + 2 * ((int64_t)j * nx + iw));
      if (eload) ev = ld_keep2(yb + 2 * ((int64_t)j * nx + ie));
#pragma unroll
      for (int k = 0; k < B200_MAX_TERMS; k++)
        if (k < nt && a.t.src[k] == B200_SRC_VECTOR) tv[k] = ld_stream2(a.t.v[k] + off);
    }
    double2 l = make_double2(0, 0), r = make_double2(0, 0);
    if (NB)
    {
      l.x = __shfl_up_sync(0xffffffffu, yc.x, 1);
      l.y = __shfl_up_sync(0xffffffffu, yc.y, 1);
      r.x = __shfl_down_sync(0xffffffffu, yc.x, 1);
      r.y = __shfl_down_sync(0xffffffffu, yc.y, 1);
      if (wload) l = wv;
      if (eload) r = ev;
    }
    if (active)
    {
      const double2 F = adr_point<MODE>(a.k, yc, l, r, ym, yp);
      if (nt > 0)
      {
        double2 acc = make_double2(0, 0);
#pragma unroll
        for (int k = 0; k < B200_MAX_TERMS; k++)
          if (k < nt)
          {
            double2 v;
            if (a.t.src[k] == B200_SRC_STENCIL) v = F;
            else if (a.t.src[k] == B200_SRC_CENTRE) v = yc;
            else v = tv[k];
            const double p0 = DMUL(a.t.c[k], v.x), p1 = DMUL(a.t.c[k], v.y);
            acc.x = (k == 0) ? p0 : DADD(acc.x, p0);
            acc.y = (k == 0) ? p1 : DADD(acc.y, p1);
          }
        *reinterpret_cast<double2*>(a.z + off) = acc;
      }
      if (a.f) *reinterpret_cast<double2*>(a.f + off) = F;
    }
    ym = yc;
    yc = yp;
    off += 2 * nx;
  }
}

