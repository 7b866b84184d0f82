// dq_kernels.cuh -- the implicit path's fused kernels (config 5: DIRK + PCG + Jacobi; also the difference-quotient
// Jacobian-vector products of the power iteration, --internaleig):
//   k_lin2_wsqr   z = ca*A + cb*B stored, sum (z*W)^2 reduced          (PCG: r -= alpha*Ap with ||r||_w; p = z + beta*p
//                                                                       with the WRMS norm arkLsDQJtimes takes next)
//   k_prod_dot    z = A.*B stored, sum z*C reduced                      (PCG: z = P^-1 r (Jacobi), <r, z>)
//   k_dq_march    z = ca*v + cb*( siginv*( L(sigma*v + y) - fy ) ) stored, optionally sum z*v reduced
//                 = arkLsATimes o arkLsDQJtimes (SUN/src/arkode/arkode_ls.c:2316-2372, :2839-2877) in ONE stencil pass:
//                 the perturbed state y + sigma*v, L of it, the difference quotient and v - gamma*Jv never reach memory.
// Every element sees the instruction sequence of the separate N_V* kernels (two roundings per multiply-add, left to
// right), so results are bit-identical to the unfused path; the reductions are the deterministic block-tree /
// last-ticket ones of reduce_prims.cuh.  Included by b200_kernels.cu (nvcc) and, under B200_HOST_EMU, by tests/emu.
#pragma once
#include "reduce_prims.cuh"

struct Lin2RedArgs
{
  const double *a, *b, *w; // w == nullptr: one weight ws for every entry
  double ca, cb, ws;
  double* z;
  int64_t n;
  double *partials, *result;
  unsigned* ticket;
};

template <bool WVEC>
__global__ void __launch_bounds__(kThreads) k_lin2_wsqr(const Lin2RedArgs a)
{
  __shared__ double smem[32];
  const int64_t n2     = a.n >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc0 = 0.0, acc1 = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n2; p += stride)
  {
    const double2 x = ld_keep2(a.a + 2 * p), y = ld_keep2(a.b + 2 * p);
    const double2 w = WVEC ? ld_keep2(a.w + 2 * p) : make_double2(a.ws, a.ws);
    double2 z;
    z.x = DADD(DMUL(a.ca, x.x), DMUL(a.cb, y.x)); // sundials_nvector.c:557-565 / nvector_parallel.c:424-517
    z.y = DADD(DMUL(a.ca, x.y), DMUL(a.cb, y.y));
    *reinterpret_cast<double2*>(a.z + 2 * p) = z;
    const double q0 = DMUL(z.x, w.x), q1 = DMUL(z.y, w.y);
    acc0 = DADD(acc0, DMUL(q0, q0));
    acc1 = DADD(acc1, DMUL(q1, q1));
  }
  if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
  {
    const int64_t i = a.n - 1;
    const double z  = DADD(DMUL(a.ca, a.a[i]), DMUL(a.cb, a.b[i]));
    a.z[i]          = z;
    const double q  = DMUL(z, WVEC ? a.w[i] : a.ws);
    acc0            = DADD(acc0, DMUL(q, q));
  }
  const double v = block_reduce<RED_SUM>(DADD(acc0, acc1), smem);
  grid_finish<RED_SUM>(v, gridDim.x, blockIdx.x, a.partials, a.ticket, a.result, smem);
}

// ewt = 1 / (rtol*|y| + atol) stored (arkEwtSetSS: N_VAbs, N_VScale, N_VAddConst, N_VInv, SUN/src/arkode/arkode.c:2932-2944,
// the rounding sequence of the four separate kernels) and, in the same pass, sum (y_i*ewt_i)^2 -- the "too much accuracy"
// norm ARKODE takes of the new y_n with the new weights at the top of the next step (arkode.c:835)
struct EwtArgs
{
  const double* y;
  double rtol, atol;
  double* ewt;
  int64_t n;
  double *partials, *result;
  unsigned* ticket;
};

__global__ void __launch_bounds__(kThreads) k_ewt_wsqr(const EwtArgs a)
{
  __shared__ double smem[32];
  const int64_t n2     = a.n >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc0 = 0.0, acc1 = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n2; p += stride)
  {
    const double2 y = ld_keep2(a.y + 2 * p);
    double2 w;
    w.x = __ddiv_rn(1.0, DADD(DMUL(a.rtol, fabs(y.x)), a.atol));
    w.y = __ddiv_rn(1.0, DADD(DMUL(a.rtol, fabs(y.y)), a.atol));
    *reinterpret_cast<double2*>(a.ewt + 2 * p) = w;
    const double q0 = DMUL(y.x, w.x), q1 = DMUL(y.y, w.y);
    acc0 = DADD(acc0, DMUL(q0, q0));
    acc1 = DADD(acc1, DMUL(q1, q1));
  }
  if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
  {
    const int64_t i = a.n - 1;
    const double w  = __ddiv_rn(1.0, DADD(DMUL(a.rtol, fabs(a.y[i])), a.atol));
    a.ewt[i]        = w;
    const double q  = DMUL(a.y[i], w);
    acc0            = DADD(acc0, DMUL(q, q));
  }
  const double v = block_reduce<RED_SUM>(DADD(acc0, acc1), smem);
  grid_finish<RED_SUM>(v, gridDim.x, blockIdx.x, a.partials, a.ticket, a.result, smem);
}

struct ProdDotArgs
{
  const double *a, *b, *c;
  double* z;
  int64_t n;
  double *partials, *result;
  unsigned* ticket;
};

__global__ void __launch_bounds__(kThreads) k_prod_dot(const ProdDotArgs a)
{
  __shared__ double smem[32];
  const int64_t n2     = a.n >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc0 = 0.0, acc1 = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n2; p += stride)
  {
    const double2 x = ld_keep2(a.a + 2 * p), y = ld_keep2(a.b + 2 * p), c = ld_keep2(a.c + 2 * p);
    double2 z;
    z.x = DMUL(x.x, y.x);
    z.y = DMUL(x.y, y.y);
    *reinterpret_cast<double2*>(a.z + 2 * p) = z;
    acc0 = DADD(acc0, DMUL(c.x, z.x));
    acc1 = DADD(acc1, DMUL(c.y, z.y));
  }
  if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
  {
    const int64_t i = a.n - 1;
    const double z  = DMUL(a.a[i], a.b[i]);
    a.z[i]          = z;
    acc0            = DADD(acc0, DMUL(a.c[i], z));
  }
  const double v = block_reduce<RED_SUM>(DADD(acc0, acc1), smem);
  grid_finish<RED_SUM>(v, gridDim.x, blockIdx.x, a.partials, a.ticket, a.result, smem);
}

// ------------------------------------------------------------------ difference-quotient matvec
struct DqArgs
{
  int64_t nx, ny;
  const double *cxw, *cxe, *cys, *cyn;
  const double *v, *y, *fy;
  double sigma, siginv, ca, cb;
  int outer;   // 1: z = ca*v + cb*Jv (arkLsATimes) ; 0: z = Jv (arkLsDQJtimes / lsrkStep_DQJtimes alone)
  int want_dot;
  double* z;
  int rows;
  double *partials, *result;
  unsigned* ticket;
};

// w = sigma*v + 1*y, the operand of the stencil (N_VLinearSum(sig, v, ONE, y, work), arkode_ls.c:2858)
__device__ __forceinline__ double dq_w(const DqArgs& a, double v, double y) { return DADD(DMUL(a.sigma, v), DMUL(1.0, y)); }

// One periodic rank (index wrap), even nx: the thread layout of k_stage_march (two adjacent cells per thread, a block
// marches down `rows` rows with the three live rows of w in registers, west / east neighbours by warp shuffle).
template <bool DOT>
__global__ void __launch_bounds__(kThreads, 4) k_dq_march(const DqArgs a)
{
  __shared__ double smem[32];
  const int64_t nx = a.nx, ny = a.ny;
  const int lane    = threadIdx.x & 31;
  const int64_t i0  = 2 * ((int64_t)blockIdx.x * kThreads + threadIdx.x);
  const bool active = (i0 < nx);
  const int64_t ic  = active ? i0 : 0;
  const int j0      = (int)blockIdx.y * a.rows;
  int j1            = j0 + a.rows;
  if (j1 > (int)ny) j1 = (int)ny;
  const bool wedge = (i0 == 0), eedge = (i0 + 2 >= nx);
  const bool wload = active && (lane == 0 || wedge);
  const bool eload = active && (lane == 31 || eedge);
  double cw0 = 0, cw1 = 0, ce0 = 0, ce1 = 0;
  if (active)
  {
    const double2 w = ld_keep2(a.cxw + ic), e = ld_keep2(a.cxe + ic);
    cw0 = w.x; cw1 = w.y; ce0 = e.x; ce1 = e.y;
  }
  const double sx0 = DADD(cw0, ce0), sx1 = DADD(cw1, ce1);
  const int64_t wcol = wedge ? nx - 1 : ic - 1; // column of the west neighbour of cell 0 / east neighbour of cell 1
  const int64_t ecol = eedge ? 0 : ic + 2;

  double2 wm = make_double2(0, 0), wc = make_double2(0, 0), vc = make_double2(0, 0);
  if (active && j0 < j1)
  {
    const int64_t rb = (int64_t)((j0 > 0) ? j0 - 1 : (int)ny - 1) * nx + ic;
    const double2 vb = ld_keep2(a.v + rb), yb = ld_keep2(a.y + rb);
    wm = make_double2(dq_w(a, vb.x, yb.x), dq_w(a, vb.y, yb.y));
    const int64_t r0 = (int64_t)j0 * nx + ic;
    vc               = ld_keep2(a.v + r0);
    const double2 yc = ld_keep2(a.y + r0);
    wc = make_double2(dq_w(a, vc.x, yc.x), dq_w(a, vc.y, yc.y));
  }
  double acc = 0.0;
#pragma unroll 1
  for (int j = j0; j < j1; j++)
  {
    double2 wp = make_double2(0, 0), vp = make_double2(0, 0), fy = make_double2(0, 0);
    double uw_edge = 0.0, ue_edge = 0.0;
    const int64_t row = (int64_t)j * nx;
    if (active)
    {
      const int64_t ra = (int64_t)((j < (int)ny - 1) ? j + 1 : 0) * nx + ic;
      vp               = ld_keep2(a.v + ra);
      const double2 yp = ld_keep2(a.y + ra);
      wp = make_double2(dq_w(a, vp.x, yp.x), dq_w(a, vp.y, yp.y));
      if (wload) uw_edge = dq_w(a, a.v[row + wcol], a.y[row + wcol]);
      if (eload) ue_edge = dq_w(a, a.v[row + ecol], a.y[row + ecol]);
      fy = ld_stream2(a.fy + row + ic);
    }
    const double dys = a.cys[j], dyn = a.cyn[j];
    const double sy  = DADD(dys, dyn);
    double uw0 = __shfl_up_sync(0xffffffffu, wc.y, 1);
    double ue1 = __shfl_down_sync(0xffffffffu, wc.x, 1);
    if (wload) uw0 = uw_edge;
    if (eload) ue1 = ue_edge;
    if (active)
    {
      // diffusion.cpp:48-53 in k_stage_march's association
      double L0 = DMUL(-DADD(sx0, sy), wc.x);
      double L1 = DMUL(-DADD(sx1, sy), wc.y);
      L0 = DADD(L0, DMUL(cw0, uw0));  L1 = DADD(L1, DMUL(cw1, wc.x));
      L0 = DADD(L0, DMUL(ce0, wc.y)); L1 = DADD(L1, DMUL(ce1, ue1));
      L0 = DADD(L0, DMUL(dys, wm.x)); L1 = DADD(L1, DMUL(dys, wm.y));
      L0 = DADD(L0, DMUL(dyn, wp.x)); L1 = DADD(L1, DMUL(dyn, wp.y));
      L0 = DADD(0.0, L0);             L1 = DADD(0.0, L1);
      // Jv = siginv*Jv - siginv*fy with equal magnitudes: a*(x - y)  (nvector_parallel.c:424-517, arkode_ls.c:2874)
      double2 z = make_double2(DMUL(a.siginv, DSUB(L0, fy.x)), DMUL(a.siginv, DSUB(L1, fy.y)));
      if (a.outer)
      { // z = v - gamma*Jv  (N_VLinearSum(ONE, v, -gamma, z, z), arkode_ls.c:2366)
        z.x = DADD(DMUL(a.ca, vc.x), DMUL(a.cb, z.x));
        z.y = DADD(DMUL(a.ca, vc.y), DMUL(a.cb, z.y));
      }
      *reinterpret_cast<double2*>(a.z + row + ic) = z;
      if (DOT) acc = DADD(acc, DADD(DMUL(z.x, vc.x), DMUL(z.y, vc.y)));
    }
    wm = wc; wc = wp; vc = vp;
  }
  if (DOT)
  {
    const double v = block_reduce<RED_SUM>(acc, smem);
    grid_finish<RED_SUM>(v, gridDim.x * gridDim.y, blockIdx.y * gridDim.x + blockIdx.x, a.partials, a.ticket, a.result, smem);
  }
}
