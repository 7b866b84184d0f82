// stage_kernels.cuh -- the fused one-stage kernels: k_stage_march (fast path), k_stage_generic (any
// width), k_stage_ring (boundary ring of a decomposed rank).  Included by b200_kernels.cu (nvcc) and,
// under B200_HOST_EMU, by the host emulation harness tests/emu.
#pragma once
#include "reduce_prims.cuh"

struct StageArgs
{
  int64_t nx, ny;
  const double *cxw, *cxe, *cys, *cyn;
  const double *hw, *he, *hs, *hn;
  const double* x;
  LinTerms t;
  double* z;
  double* f_out;
  double *send_w, *send_e, *send_s, *send_n;
  const double* rw;
  double* partials;
  unsigned* ticket;
  double* result;
  // RED == 2: also the error weights of the stencil input itself, e = 1/(rtol*|x| + atol) (arkEwtSetSS,
  // arkode.c:2932-2944: N_VAbs, N_VScale, N_VAddConst, N_VInv) stored to ewt_out, and sum (x*e)^2 into result2 --
  // what ARKODE asks for next if x becomes y_{n+1}
  double* ewt_out;
  double ewt_rtol, ewt_atol;
  double* result2;
  int rows;   // rows marched per block
  int region; // 0 all, 1 ring, 2 interior
};

// 5-point operator in the reference's association order
// (diffusion_2D/diffusion.cpp:48-53; f starts at 0 and is "+="-ed):
//   0 + (((( -((Dxw+Dxe)+(Dys+Dyn)) * uc + Dxw*uw ) + Dxe*ue ) + Dys*us ) + Dyn*un )
__device__ __forceinline__ double lap5(double dxw, double dxe, double dys, double dyn,
                                       double uc, double uw, double ue, double us, double un)
{
  const double dc = -DADD(DADD(dxw, dxe), DADD(dys, dyn));
  double r        = DMUL(dc, uc);
  r               = DADD(r, DMUL(dxw, uw));
  r               = DADD(r, DMUL(dxe, ue));
  r               = DADD(r, DMUL(dys, us));
  r               = DADD(r, DMUL(dyn, un));
  return DADD(0.0, r);
}

// one cell, generic neighbour access (ring kernel, odd-width fallback)
__device__ __forceinline__ void stage_cell(const StageArgs& a, int64_t i, int64_t j, double* wr_acc, double* we_acc = nullptr)
{
  const int64_t nx = a.nx, ny = a.ny;
  const int64_t id = j * nx + i;
  const double uc  = a.x[id];
  const double uw  = (i > 0) ? a.x[id - 1] : (a.hw ? a.hw[j] : a.x[id + nx - 1]);
  const double ue  = (i < nx - 1) ? a.x[id + 1] : (a.he ? a.he[j] : a.x[id - (nx - 1)]);
  const double us  = (j > 0) ? a.x[id - nx] : (a.hs ? a.hs[i] : a.x[(ny - 1) * nx + i]);
  const double un  = (j < ny - 1) ? a.x[id + nx] : (a.hn ? a.hn[i] : a.x[i]);
  const double L   = lap5(a.cxw[i], a.cxe[i], a.cys[j], a.cyn[j], uc, uw, ue, us, un);
  double acc       = 0.0;
#pragma unroll
  for (int k = 0; k < B200_MAX_TERMS; k++)
    if (k < a.t.n)
    {
      const double tv = (a.t.src[k] == B200_SRC_STENCIL) ? L
                        : (a.t.src[k] == B200_SRC_CENTRE) ? uc
                                                          : a.t.v[k][id];
      const double pr = DMUL(a.t.c[k], tv);
      acc             = (k == 0) ? pr : DADD(acc, pr);
    }
  a.z[id] = acc;
  if (a.f_out) a.f_out[id] = L;
  if (a.send_w && i == 0) a.send_w[j] = acc;
  if (a.send_e && i == nx - 1) a.send_e[j] = acc;
  if (a.send_s && j == 0) a.send_s[i] = acc;
  if (a.send_n && j == ny - 1) a.send_n[i] = acc;
  if (wr_acc)
  {
    const double p = DMUL(acc, a.rw[id]);
    *wr_acc        = DADD(*wr_acc, DMUL(p, p));
  }
  if (we_acc)
  {
    const double e = __ddiv_rn(1.0, DADD(DMUL(a.ewt_rtol, fabs(uc)), a.ewt_atol));
    a.ewt_out[id]  = e;
    const double q = DMUL(uc, e);
    *we_acc        = DADD(*we_acc, DMUL(q, q));
  }
}

// Generic kernel: one thread per cell, any nx/ny; region-aware.
__global__ void __launch_bounds__(kThreads) k_stage_generic(const StageArgs a)
{
  __shared__ double smem[32];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  double wr = 0.0, we = 0.0;
  if (i < a.nx)
  {
    const bool ring = (i == 0 || i == a.nx - 1 || j == 0 || j == a.ny - 1);
    if (a.region == 0 || (a.region == 1 && ring) || (a.region == 2 && !ring))
      stage_cell(a, i, j, a.rw ? &wr : nullptr, a.ewt_out ? &we : nullptr);
  }
  if (a.rw && a.ewt_out)
  {
    double v  = block_reduce<RED_SUM>(wr, smem);
    double v2 = block_reduce<RED_SUM>(we, smem);
    grid_finish2(v, v2, gridDim.x * gridDim.y, blockIdx.y * gridDim.x + blockIdx.x, a.partials, a.ticket, a.result,
                 a.result2, smem);
  }
  else if (a.rw)
  {
    double v = block_reduce<RED_SUM>(wr, smem);
    grid_finish<RED_SUM>(v, gridDim.x * gridDim.y, blockIdx.y * gridDim.x + blockIdx.x,
                         a.partials, a.ticket, a.result, smem);
  }
}

// Ring kernel: the 2*nx + 2*(ny-2) boundary cells only (they are the only ones
// that read halos and the only ones that are packed for the neighbours).
__global__ void __launch_bounds__(kThreads) k_stage_ring(const StageArgs a)
{
  const int64_t t  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nx = a.nx, ny = a.ny;
  int64_t i, j;
  if (t < nx) { i = t; j = 0; }
  else if (t < 2 * nx) { i = t - nx; j = ny - 1; }
  else if (t < 2 * nx + (ny - 2)) { i = 0; j = t - 2 * nx + 1; }
  else if (t < 2 * nx + 2 * (ny - 2)) { i = nx - 1; j = t - 2 * nx - (ny - 2) + 1; }
  else return;
  if (ny == 1 && t >= nx) return;
  stage_cell(a, i, j, nullptr);
}

// Fast path (nx even): each thread owns two adjacent cells (one double2) of a
// 512-cell-wide strip and marches down `rows` rows keeping the three live rows of
// x in registers, so every x row is loaded from L2/HBM once per block.  West/east
// neighbours come from warp shuffles; only lanes 0 / 31 (and the strip ends) issue
// an extra scalar load.  Per cell-update: 4 x 8 B streamed in + 8 B out.
//
// The term pattern (which of the NT terms is a vector / the stencil input / L(x))
// is a template parameter for the sequences LSRKStep actually issues, so the row
// loop carries no dispatch; PAT_RUNTIME keeps a fully general fallback.  All
// addresses are running pointers (one add per row).
#define PAT_RUNTIME 0xffffffffu
#define PAT1(a) (uint32_t)(a)
#define PAT2(a, b) (uint32_t)((a) | ((b) << 2))
#define PAT3(a, b, c) (uint32_t)((a) | ((b) << 2) | ((c) << 4))
#define PAT4(a, b, c, d) (uint32_t)((a) | ((b) << 2) | ((c) << 4) | ((d) << 6))
#define PAT5(a, b, c, d, e) (uint32_t)((a) | ((b) << 2) | ((c) << 4) | ((d) << 6) | ((e) << 8))

// HAS_RED: 0 = no reduction, 1 = sum (z*w)^2, 2 = that and the error weights of x with their own norm (StageArgs::ewt_out)
template <int NT, uint32_t PAT, int REGION, int HAS_RED>
__global__ void __launch_bounds__(kThreads, 4) k_stage_march(const StageArgs a)
{
  __shared__ double smem[32];
  const int64_t nx = a.nx, ny = a.ny;
  const int lane    = threadIdx.x & 31;
  const int64_t i0  = 2 * ((int64_t)blockIdx.x * kThreads + threadIdx.x);
  const bool active = (i0 < nx);
  const int64_t ic  = active ? i0 : 0; // clamp so address arithmetic stays in range
  int j0            = (int)blockIdx.y * a.rows; // ny < 2^31 (checked by the launcher)
  int j1            = j0 + a.rows;
  if (j1 > (int)ny) j1 = (int)ny;
  if (REGION == 2)
  {
    if (j0 < 1) j0 = 1;
    if (j1 > (int)ny - 1) j1 = (int)ny - 1;
  }
  const int jlast = (int)ny - 1;
  const int nt = (PAT == PAT_RUNTIME) ? a.t.n : NT;
#define SRC_OF(k) ((PAT == PAT_RUNTIME) ? a.t.src[k] : (int)((PAT >> (2 * (k))) & 3u))

  const bool wedge = (i0 == 0);      // west neighbour lies outside the field
  const bool eedge = (i0 + 2 >= nx); // east neighbour lies outside the field
  const bool wload = active && (lane == 0 || wedge);
  const bool eload = active && (lane == 31 || eedge);

  // x-direction face coefficients of my two cells, and their sums (diffusion.cpp:48)
  double cw0 = 0, cw1 = 0, ce0 = 0, ce1 = 0;
  if (active)
  {
    const double2 w = ld_keep2(a.cxw + ic), e = ld_keep2(a.cxe + ic);
    cw0 = w.x; cw1 = w.y; ce0 = e.x; ce1 = e.y;
  }
  const double sx0 = DADD(cw0, ce0), sx1 = DADD(cw1, ce1);

  // running pointers: current x row, west/east edge values, and the element offset
  int64_t off        = (int64_t)j0 * nx + ic;
  const double* xrow = a.x + off;
  const double* wptr;
  const double* eptr;
  int64_t wstep = nx, estep = nx;
  if (!wedge) wptr = xrow - 1;
  else if (a.hw) { wptr = a.hw + j0; wstep = 1; }
  else wptr = xrow + (nx - 1);
  if (!eedge) eptr = xrow + 2;
  else if (a.he) { eptr = a.he + j0; estep = 1; }
  else eptr = a.x + (int64_t)j0 * nx;
  const bool wvalid = wload && !(REGION == 2 && wedge);
  const bool evalid = eload && !(REGION == 2 && eedge);

  double2 xm = make_double2(0, 0), xc = make_double2(0, 0);
  if (active && j0 < j1)
  {
    const double* below = (j0 > 0) ? (xrow - nx) : (a.hs ? a.hs + ic : a.x + (ny - 1) * nx + ic);
    xm = ld_keep2(below);
    xc = ld_keep2(xrow);
  }
  double wr = 0.0, we = 0.0;
  // loop-invariant switches (all uniform or per-thread constants)
  const bool has_f  = (a.f_out != nullptr);
  const bool do_sw  = (REGION == 0) && a.send_w && wedge;
  const bool do_se  = (REGION == 0) && a.send_e && eedge;
  const bool do_ss  = (REGION == 0) && a.send_s && (j0 == 0);
  const bool do_sn  = (REGION == 0) && a.send_n && (j1 == (int)ny);
  const double* wrapn = a.hn ? a.hn + ic : a.x + ic; // row "ny": north halo or periodic wrap

#pragma unroll 1
  for (int j = j0; j < j1; j++)
  {
    double2 xp = make_double2(0, 0);
    double uw_edge = 0.0, ue_edge = 0.0;
    double2 tv[NT];
    if (active)
    {
      const double* above = (j < jlast) ? (xrow + nx) : wrapn;
      xp = ld_keep2(above);
      if (wvalid) uw_edge = *wptr;
      if (evalid) ue_edge = *eptr;
#pragma unroll
      for (int k = 0; k < NT; k++)
        if (k < nt && SRC_OF(k) == B200_SRC_VECTOR) tv[k] = ld_stream2(a.t.v[k] + off);
    }
    const double dys = a.cys[j], dyn = a.cyn[j];
    const double sy  = DADD(dys, dyn);
    // west of cell0 = previous lane's cell1 ; east of cell1 = next lane's cell0
    double uw0 = __shfl_up_sync(0xffffffffu, xc.y, 1);
    double ue1 = __shfl_down_sync(0xffffffffu, xc.x, 1);
    if (wload) uw0 = uw_edge;
    if (eload) ue1 = ue_edge;
    if (active)
    {
      // diffusion.cpp:48-53, same association: ((((dc*uc + Dxw*uw) + Dxe*ue) + Dys*us) + Dyn*un)
      double L0 = DMUL(-DADD(sx0, sy), xc.x);
      double L1 = DMUL(-DADD(sx1, sy), xc.y);
      L0 = DADD(L0, DMUL(cw0, uw0));  L1 = DADD(L1, DMUL(cw1, xc.x));
      L0 = DADD(L0, DMUL(ce0, xc.y)); L1 = DADD(L1, DMUL(ce1, ue1));
      L0 = DADD(L0, DMUL(dys, xm.x)); L1 = DADD(L1, DMUL(dys, xm.y));
      L0 = DADD(L0, DMUL(dyn, xp.x)); L1 = DADD(L1, DMUL(dyn, xp.y));
      L0 = DADD(0.0, L0);             L1 = DADD(0.0, L1);
      double2 acc = make_double2(0, 0);
#pragma unroll
      for (int k = 0; k < NT; k++)
        if (k < nt)
        {
          double2 v;
          const int sk = SRC_OF(k);
          if (sk == B200_SRC_STENCIL) v = make_double2(L0, L1);
          else if (sk == B200_SRC_CENTRE) v = xc;
          else v = tv[k];
          const double p0 = DMUL(a.t.c[k], v.x), p1 = DMUL(a.t.c[k], v.y);
          acc.x = (k == 0) ? p0 : DADD(acc.x, p0);
          acc.y = (k == 0) ? p1 : DADD(acc.y, p1);
        }
      double* zp = a.z + off;
      if (REGION == 2 && (wedge || eedge))
      { // ring cells belong to the ring kernel
        if (!wedge) zp[0] = acc.x;
        if (!eedge) zp[1] = acc.y;
        if (has_f)
        {
          if (!wedge) a.f_out[off] = L0;
          if (!eedge) a.f_out[off + 1] = L1;
        }
      }
      else
      {
        *reinterpret_cast<double2*>(zp) = acc;
        if (has_f) *reinterpret_cast<double2*>(a.f_out + off) = make_double2(L0, L1);
      }
      if (REGION == 0)
      {
        if (do_sw) a.send_w[j] = acc.x;
        if (do_se) a.send_e[j] = acc.y;
        if (do_ss && j == 0) *reinterpret_cast<double2*>(a.send_s + ic) = acc;
        if (do_sn && j == jlast) *reinterpret_cast<double2*>(a.send_n + ic) = acc;
      }
      if (HAS_RED)
      {
        const double2 w = ld_stream2(a.rw + off);
        const double q0 = DMUL(acc.x, w.x), q1 = DMUL(acc.y, w.y);
        wr = DADD(wr, DADD(DMUL(q0, q0), DMUL(q1, q1)));
      }
      if (HAS_RED == 2)
      {
        const double e0 = __ddiv_rn(1.0, DADD(DMUL(a.ewt_rtol, fabs(xc.x)), a.ewt_atol));
        const double e1 = __ddiv_rn(1.0, DADD(DMUL(a.ewt_rtol, fabs(xc.y)), a.ewt_atol));
        *reinterpret_cast<double2*>(a.ewt_out + off) = make_double2(e0, e1);
        const double s0 = DMUL(xc.x, e0), s1 = DMUL(xc.y, e1);
        we = DADD(we, DADD(DMUL(s0, s0), DMUL(s1, s1)));
      }
    }
    xm = xc;
    xc = xp;
    xrow += nx;
    off += nx;
    wptr += wstep;
    eptr += estep;
  }
#undef SRC_OF
  if (HAS_RED == 2)
  {
    double v  = block_reduce<RED_SUM>(wr, smem);
    double v2 = block_reduce<RED_SUM>(we, smem);
    grid_finish2(v, v2, gridDim.x * gridDim.y, blockIdx.y * gridDim.x + blockIdx.x, a.partials, a.ticket, a.result,
                 a.result2, smem);
  }
  else if (HAS_RED)
  {
    double v = block_reduce<RED_SUM>(wr, smem);
    grid_finish<RED_SUM>(v, gridDim.x * gridDim.y, blockIdx.y * gridDim.x + blockIdx.x,
                         a.partials, a.ticket, a.result, smem);
  }
}

