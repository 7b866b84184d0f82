// halo_kernels.cuh -- deep-halo strip packing (k_pack_strips), the one-deep edge pack (k_pack) and the
// Jacobi diagonal (k_jacobi).  Included by b200_kernels.cu (nvcc) and, under B200_HOST_EMU, by the host
// emulation harness tests/emu.
#pragma once
#include "reduce_prims.cuh"

// W / E strips of one field: columns [0, g2) and [nx-g2, nx) over rows -g..ny+g-1, the rows
// outside the field taken from the S / N halo (so corners travel with the second phase).
__global__ void __launch_bounds__(kThreads)
  k_pack_strips(const double* __restrict__ field, const double* __restrict__ halo, int64_t nx, int64_t ny,
                int g, int g2, double* __restrict__ wstrip, double* __restrict__ estrip)
{
  const int64_t t     = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nrows = ny + 2 * g;
  if (t >= nrows * g2) return;
  const int64_t rr = t / g2; // 0 .. ny+2g-1  <->  row rr - g
  const int cc     = (int)(t - rr * g2);
  const int64_t r  = rr - g;
  const double* row;
  if (r < 0) row = halo + (r + g) * nx;
  else if (r >= ny) row = halo + (g + (r - ny)) * nx;
  else row = field + r * nx;
  wstrip[t] = row[cc];
  estrip[t] = row[nx - g2 + cc];
}


__global__ void __launch_bounds__(kThreads)
  k_pack(const double* __restrict__ u, int64_t nx, int64_t ny, double* sw, double* se,
         double* ss, double* sn)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ny)
  {
    if (sw) sw[t] = u[t * nx];
    if (se) se[t] = u[t * nx + nx - 1];
  }
  if (t < nx)
  {
    if (ss) ss[t] = u[t];
    if (sn) sn[t] = u[(ny - 1) * nx + t];
  }
}


__global__ void __launch_bounds__(kThreads)
  k_jacobi(int64_t nx, int64_t ny, const double* __restrict__ pxw, const double* __restrict__ pxe,
           const double* __restrict__ pys, const double* __restrict__ pyn, double gamma, double* diag)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= nx) return;
  // preconditioner_jacobi.cpp:41-42: diag = -((Dx_w+Dx_e)+(Dy_s+Dy_n)); 1/(1 - gamma*diag)
  const double d   = -DADD(DADD(pxw[i], pxe[i]), DADD(pys[j], pyn[j]));
  diag[j * nx + i] = __ddiv_rn(1.0, DSUB(1.0, DMUL(gamma, d)));
}

