// react_kernels.cuh -- the implicit reaction partition of the adr 2-D driver: Jacobian of the Brusselator
// reaction and the direct solver of its Newton systems.  Included by b200_kernels.cu (nvcc) and, under
// B200_HOST_EMU, by the host emulation harness tests/emu.
//
// The reference stores J_reaction (adr/advection_diffusion_reaction_2d.cpp:1523-1551) in a SUNBandMatrix(neq, 2, 2)
// and factors I - gamma*J with SUNLinSol_Band (LAPACK-style banded LU with partial pivoting,
// SUN/src/sundials/sundials_band.c bandGBTRF / bandGBTRS).  With the two species of a grid point interleaved the
// matrix is BLOCK DIAGONAL, one 2x2 block per grid point: every entry outside the blocks is a structural zero of the
// band, so the banded elimination never leaves a block -- the pivot search of column 2p can only pick row 2p or
// 2p + 1 (the third candidate is zero), column 2p + 1 keeps its diagonal, and all the multipliers that reach into
// the next block are zero.  The kernels below run exactly the operations bandGBTRF / bandGBTRS perform inside one
// block, in their order and with their roundings (separate multiply and add), one thread per grid point.
//
// Storage: 4 doubles per grid point { a00, a10, a01, a11 } = d(row)/d(col) with row / col in {u, v}:
// a10 = dV/du is the sub-diagonal entry of column u, a01 = dU/dv the super-diagonal entry of column v.
#pragma once
#include "adr_kernels.cuh"

// J_reaction, ...2d.cpp:1540-1546
__global__ void __launch_bounds__(kThreads) k_adr_jac_reaction(int64_t npts, double B, double Bp1, const double* __restrict__ y,
                                                                double* __restrict__ J)
{
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const double2 w  = ld_keep2(y + 2 * p);
  const double u = w.x, v = w.y;
  const double tuv = DMUL(DMUL(2.0, u), v); // TWO * u * v, left to right
  const double uu  = DMUL(u, u);
  double2 c0, c1;
  c0.x = DSUB(tuv, Bp1); // dU/du = 2uv - (B + 1)
  c0.y = DSUB(B, tuv);   // dV/du = B - 2uv
  c1.x = uu;             // dU/dv = u^2
  c1.y = -uu;            // dV/dv = -u^2
  *reinterpret_cast<double2*>(J + 4 * p)     = c0;
  *reinterpret_cast<double2*>(J + 4 * p + 2) = c1;
}

// SUNMatScaleAddI_Band (SUN/src/sunmatrix/band/sunmatrix_band.c): every stored entry *= c, then diagonal += 1
__global__ void __launch_bounds__(kThreads) k_blk2_scale_add_i(int64_t npts, double c, double* __restrict__ A)
{
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  double2 c0 = ld_keep2(A + 4 * p), c1 = ld_keep2(A + 4 * p + 2);
  c0.x = DADD(DMUL(c, c0.x), 1.0);
  c0.y = DMUL(c, c0.y);
  c1.x = DMUL(c, c1.x);
  c1.y = DADD(DMUL(c, c1.y), 1.0);
  *reinterpret_cast<double2*>(A + 4 * p)     = c0;
  *reinterpret_cast<double2*>(A + 4 * p + 2) = c1;
}

// bandGBTRF restricted to one block (elimination steps k = 2p and k = 2p + 1).  piv[p] = 1: rows u and v were swapped.
// fail: smallest 1-based column with a zero pivot (0 = none), as bandGBTRF returns it.
__global__ void __launch_bounds__(kThreads) k_blk2_factor(int64_t npts, double* __restrict__ A, double* __restrict__ piv,
                                                           unsigned long long* fail)
{
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  double2 c0 = ld_keep2(A + 4 * p), c1 = ld_keep2(A + 4 * p + 2);
  double a00 = c0.x, a10 = c0.y, a01 = c1.x, a11 = c1.y;
  // step k = 2p: pivot = the larger of |a00|, |a10| (strictly larger wins, ties keep the diagonal)
  const bool swap = fabs(a10) > fabs(a00);
  const double pivot = swap ? a10 : a00;
  unsigned long long bad = 0;
  if (pivot == 0.0) bad = (unsigned long long)(2 * p + 1);
  else
  {
    if (swap) { a10 = a00; a00 = pivot; }
    const double mult = __ddiv_rn(-1.0, a00);
    a10 = DMUL(a10, mult); // multiplier -a(i,k)/a(k,k)
    // column j = 2p + 1: a_kj = a(l, j); swap a(k,j) <-> a(l,j); a(i,j) += a_kj * multiplier
    const double akj = swap ? a11 : a01;
    if (swap) { a11 = a01; a01 = akj; }
    if (akj != 0.0) a11 = DADD(a11, DMUL(akj, a10));
    // step k = 2p + 1: the diagonal is the only non-zero candidate
    if (a11 == 0.0) bad = (unsigned long long)(2 * p + 2);
  }
  *reinterpret_cast<double2*>(A + 4 * p)     = make_double2(a00, a10);
  *reinterpret_cast<double2*>(A + 4 * p + 2) = make_double2(a01, a11);
  piv[p] = swap ? 1.0 : 0.0;
  if (bad) atomicMin(fail, bad);
}

// bandGBTRS restricted to one block: x = b; forward (L y = P b), backward (U x = y)
__global__ void __launch_bounds__(kThreads) k_blk2_solve(int64_t npts, const double* __restrict__ A, const double* __restrict__ piv,
                                                          const double* __restrict__ b, double* __restrict__ x)
{
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const double2 c0 = ld_keep2(A + 4 * p), c1 = ld_keep2(A + 4 * p + 2);
  double2 r = ld_keep2(b + 2 * p);
  double b0 = r.x, b1 = r.y;
  if (piv[p] != 0.0) { const double t = b0; b0 = b1; b1 = t; }
  b1 = DADD(b1, DMUL(b0, c0.y));        // b[i] += mult * a(i,k), k = 2p
  b1 = __ddiv_rn(b1, c1.y);             // k = 2p + 1: b[k] /= a(k,k)
  b0 = DADD(b0, DMUL(-b1, c1.x));       //             b[i] += (-b[k]) * a(i,k)
  b0 = __ddiv_rn(b0, c0.x);             // k = 2p
  *reinterpret_cast<double2*>(x + 2 * p) = make_double2(b0, b1);
}
