// adr_chain.cuh -- k_adr_chain: K temporally blocked STS stages of the adr 2-D diffusion partition
// (f_diffusion, adr/advection_diffusion_reaction_2d.cpp:1448-1491) per launch.
//
//   z_l = c[l][0] F(z_{l-1}) + c[l][1] z_{l-2} + c[l][2] yn + c[l][3] z_{l-1} + c[l][4] fn ,  l = 1..K,
//   z_0 = x, z_{-1} = prev2,  F = adr_point<2> (both species of a grid point, the reference's order),
// each value produced by exactly the instruction sequence of k_adr_march<2> with the term pattern
// [F(y), v, v, y, v], so the result is bit-identical to K separate b200_adr_lincomb launches.
//
// Structure = k_chain_quad's (chain_quad.cuh) with "two cells" replaced by "one grid point = (u, v)":
// a warp owns a window of 64 grid points made of two halves of 32; lane i owns point i of each half, so
// every global access is 16 B per lane / 512 contiguous bytes per warp instruction; K halo points at each
// end of the window (K = 4: 56 of 64 points stored); operands are staged through a thread-private
// cp.async ring; west/east neighbours (whole points) come from rotating warp shuffles with one select per
// side at the seam between the halves.  The problem is periodic on one rank (the reference's adr driver is
// serial), so there is no halo flavour; the coefficients are the two scalars d/dx^2, d/dy^2.
// Why: at BASELINE configs[3] (2048^2 x 2 species, 64 MiB per vector) a stage is ~60 us of kernel and the
// run is launch-bound (2458 launches in 0.21 s); K stages per launch cut both launches and traffic.
// Status: bit-exact on the host emulator (tests/test_kernel_emulation.py); behind `--sts_chain K` of the
// adr driver, default 1 = off until it has been measured on the GPU.
#pragma once
#include "adr_kernels.cuh"
#include "chain_march.cuh" // cp.async ring conventions, B200_MAX_CHAIN

struct AdrChainArgs
{
  int64_t nx, ny; // grid points
  AdrConsts k;
  const double* x;
  const double* prev2;
  const double* yn;
  const double* fn;
  double c[B200_MAX_CHAIN][5];
  double* out[B200_MAX_CHAIN];
  int rows;
};

static const int kAdrChainThreads = 128;
static const int kAdrChainPF      = 3;

struct AdrHalf
{
  const double *px, *pp, *py, *pf; // next group: x at row ir+1 ; prev2 / yn / fn at row ir
  int64_t col;                     // wrapped grid-point column of this lane's point
  int64_t soff;                    // 2*(r1*nx + col) with the unwrapped row r1
};

struct AdrChainState
{
  AdrHalf h[2];
  int ir;
  int sx_issue, sy_issue, sx_use, sy_use;
};

__device__ __forceinline__ const double* adr_row_ptr(const double* field, int r, int64_t col, int64_t nx, int ny)
{
  const int rw = (r < 0) ? r + ny : ((r >= ny) ? r - ny : r);
  return field + 2 * ((int64_t)rw * nx + col);
}

// WRAP = false (steady state of the row loop): neither row of the next group is row 0 or row ny
template <bool WRAP>
__device__ __forceinline__ void adr_half_advance(const AdrChainArgs& a, AdrHalf& q, int r, int64_t nx, int ny)
{
  if (WRAP && (r == 0 || r == ny))
  {
    q.pp = adr_row_ptr(a.prev2, r, q.col, nx, ny);
    q.py = adr_row_ptr(a.yn, r, q.col, nx, ny);
    q.pf = adr_row_ptr(a.fn, r, q.col, nx, ny);
  }
  else { q.pp += 2 * nx; q.py += 2 * nx; q.pf += 2 * nx; }
  if (WRAP && (r + 1 == 0 || r + 1 == ny)) q.px = adr_row_ptr(a.x, r + 1, q.col, nx, ny);
  else q.px += 2 * nx;
}

template <int K, int PF, bool WRAP = true>
__device__ __forceinline__ void adr_chain_issue(const AdrChainArgs& a, AdrChainState& st, double2* rx, double2* rp,
                                                double2* ry, double2* rf, int64_t nx, int ny, bool issue)
{
  constexpr int DX = PF + 1, DY = PF + K;
  if (issue)
  {
    double2* sx = rx + 2 * st.sx_issue * kAdrChainThreads;
    double2* sp = rp + 2 * st.sx_issue * kAdrChainThreads;
    double2* sy = ry + 2 * st.sy_issue * kAdrChainThreads;
    double2* sf = rf + 2 * st.sy_issue * kAdrChainThreads;
    cp_async16(sx, st.h[0].px);           cp_async16(sx + kAdrChainThreads, st.h[1].px);
    cp_async16(sp, st.h[0].pp);           cp_async16(sp + kAdrChainThreads, st.h[1].pp);
    cp_async16(sy, st.h[0].py);           cp_async16(sy + kAdrChainThreads, st.h[1].py);
    cp_async16(sf, st.h[0].pf);           cp_async16(sf + kAdrChainThreads, st.h[1].pf);
  }
  cp_async_commit();
  st.sx_issue = ring_next<DX>(st.sx_issue);
  st.sy_issue = ring_next<DY>(st.sy_issue);
  const int r = ++st.ir;
  adr_half_advance<WRAP>(a, st.h[0], r, nx, ny);
  adr_half_advance<WRAP>(a, st.h[1], r, nx, ny);
}

__device__ __forceinline__ double2 shfl_point(double2 v, int src_lane)
{
  return make_double2(__shfl_sync(0xffffffffu, v.x, src_lane), __shfl_sync(0xffffffffu, v.y, src_lane));
}

// the stage combination in k_adr_march's order for the pattern [F, v, v, y, v] (left to right)
__device__ __forceinline__ double2 adr_stage_point(const double* cf, double2 F, double2 p2, double2 yv, double2 uc, double2 fv)
{
  double2 z;
  z.x = DMUL(cf[0], F.x);               z.y = DMUL(cf[0], F.y);
  z.x = DADD(z.x, DMUL(cf[1], p2.x));   z.y = DADD(z.y, DMUL(cf[1], p2.y));
  z.x = DADD(z.x, DMUL(cf[2], yv.x));   z.y = DADD(z.y, DMUL(cf[2], yv.y));
  z.x = DADD(z.x, DMUL(cf[3], uc.x));   z.y = DADD(z.y, DMUL(cf[3], uc.y));
  z.x = DADD(z.x, DMUL(cf[4], fv.x));   z.y = DADD(z.y, DMUL(cf[4], fv.y));
  return z;
}

template <int K, int PF, int PH, bool CHECK>
__device__ __forceinline__ void adr_chain_row(const AdrChainArgs& a, AdrChainState& st, double2 (&Wa)[K][3],
                                              double2 (&Wb)[K][3], double2* rx, double2* rp, double2* ry, double2* rf,
                                              int64_t nx, unsigned smask_a, unsigned smask_b, int lane, int r1, int j0,
                                              int j1, int ny, bool issue)
{
  constexpr int DX = PF + 1, DY = PF + K;
  constexpr int IO = PH % 3, IM = (PH + 1) % 3, IC = (PH + 2) % 3; // oldest (overwritten), then south, centre ; north = IO
  adr_chain_issue<K, PF, CHECK>(a, st, rx, rp, ry, rf, nx, ny, issue); // group(r1 + PF)
  cp_async_wait<PF>();

  Wa[0][IO]        = rx[(2 * st.sx_use) * kAdrChainThreads]; // x row r1+1 replaces the oldest row
  Wb[0][IO]        = rx[(2 * st.sx_use + 1) * kAdrChainThreads];
  const double2 Pa = rp[(2 * st.sx_use) * kAdrChainThreads];
  const double2 Pb = rp[(2 * st.sx_use + 1) * kAdrChainThreads];
  int64_t soa = st.h[0].soff, sob = st.h[1].soff;
  const int lw = (lane + 31) & 31, le = (lane + 1) & 31;
#pragma unroll
  for (int l = 1; l <= K; l++)
  {
    const double2 sa = Wa[l - 1][IM], ca = Wa[l - 1][IC], na = Wa[l - 1][IO];
    const double2 sb = Wb[l - 1][IM], cb = Wb[l - 1][IC], nb = Wb[l - 1][IO];
    // rotating shuffles of whole points; seam: west of half b's lane 0 is half a's lane 31 and vice versa
    const double2 wa = shfl_point(ca, lw), wb = shfl_point(cb, lw);
    const double2 ea = shfl_point(ca, le), eb = shfl_point(cb, le);
    const double2 west_a = wa, west_b = (lane == 0) ? wa : wb;
    const double2 east_a = (lane == 31) ? eb : ea, east_b = eb;
    const double2 p2a = (l == 1) ? Pa : Wa[(l >= 2) ? l - 2 : 0][IM];
    const double2 p2b = (l == 1) ? Pb : Wb[(l >= 2) ? l - 2 : 0][IM];
    int sl = st.sy_use - (l - 1); // yn / fn of row r1-(l-1)
    if (sl < 0) sl += DY;
    const double2 yva = ry[(2 * sl) * kAdrChainThreads], yvb = ry[(2 * sl + 1) * kAdrChainThreads];
    const double2 fva = rf[(2 * sl) * kAdrChainThreads], fvb = rf[(2 * sl + 1) * kAdrChainThreads];
    const double2 Fa = adr_point<2>(a.k, ca, west_a, east_a, sa, na);
    const double2 Fb = adr_point<2>(a.k, cb, west_b, east_b, sb, nb);
    const double2 za = adr_stage_point(a.c[l - 1], Fa, p2a, yva, ca, fva);
    const double2 zb = adr_stage_point(a.c[l - 1], Fb, p2b, yvb, cb, fvb);
    bool row_ok = true;
    if (CHECK)
    {
      const int rl = r1 - (l - 1);
      row_ok       = rl >= j0 && rl < j1;
    }
    if (row_ok && ((smask_a >> (l - 1)) & 1u)) *reinterpret_cast<double2*>(a.out[l - 1] + soa) = za;
    if (row_ok && ((smask_b >> (l - 1)) & 1u)) *reinterpret_cast<double2*>(a.out[l - 1] + sob) = zb;
    soa -= 2 * nx;
    sob -= 2 * nx;
    if (l < K) { Wa[l][IO] = za; Wb[l][IO] = zb; }
  }
  st.h[0].soff += 2 * nx;
  st.h[1].soff += 2 * nx;
  st.sx_use = (st.sx_use + 1 == DX) ? 0 : st.sx_use + 1;
  st.sy_use = (st.sy_use + 1 == DY) ? 0 : st.sy_use + 1;
}

__device__ __forceinline__ void adr_half_setup(const AdrChainArgs& a, AdrHalf& q, int64_t col_u, int64_t nx, int ny, int rstart)
{
  int64_t c = col_u;
  if (c < 0) c += nx;
  else if (c >= nx) c -= nx;
  q.col  = c;
  q.soff = 2 * ((int64_t)rstart * nx + c);
  q.px   = adr_row_ptr(a.x, rstart + 1, c, nx, ny);
  q.pp   = adr_row_ptr(a.prev2, rstart, c, nx, ny);
  q.py   = adr_row_ptr(a.yn, rstart, c, nx, ny);
  q.pf   = adr_row_ptr(a.fn, rstart, c, nx, ny);
}

template <int K, int PF>
__global__ void __launch_bounds__(kAdrChainThreads, 2) k_adr_chain(const AdrChainArgs a)
{
  constexpr int WUSE = 64 - 2 * K; // grid points a warp stores per row (K halo points at each end)
  constexpr int DX   = PF + 1;
  constexpr int DY   = PF + K;
  B200_DYN_SMEM(double2, ring);
  double2* rx = ring + threadIdx.x;                                                // [DX][2][threads]
  double2* rp = ring + (size_t)2 * DX * kAdrChainThreads + threadIdx.x;            // [DX][2][threads]
  double2* ry = ring + (size_t)4 * DX * kAdrChainThreads + threadIdx.x;            // [DY][2][threads]
  double2* rf = ring + (size_t)(4 * DX + 2 * DY) * kAdrChainThreads + threadIdx.x; // [DY][2][threads]

  const int lane   = threadIdx.x & 31;
  const int64_t nx = a.nx;
  const int ny     = (int)a.ny;
  const int j0     = (int)blockIdx.y * a.rows;
  int j1           = j0 + a.rows;
  if (j1 > ny) j1 = ny;
  const int rstart = j0 - (K - 1), rend = j1 + (K - 1); // level-1 rows [rstart, rend)

  const int64_t wg = (int64_t)blockIdx.x * (kAdrChainThreads / 32) + (threadIdx.x >> 5);
  if (wg * WUSE >= nx) return; // window entirely outside the field (no block-level sync in this kernel)
  const int64_t col_a = wg * WUSE - K + lane; // unwrapped column of my point, left half
  const int64_t col_b = col_a + 32;           // right half
  const bool ok_a     = (lane >= K) && (col_a < nx);
  const bool ok_b     = (lane < 32 - K) && (col_b < nx);
  unsigned smask_a = 0, smask_b = 0;
#pragma unroll
  for (int l = 0; l < K; l++)
    if (a.out[l])
    {
      if (ok_a) smask_a |= 1u << l;
      if (ok_b) smask_b |= 1u << l;
    }

  AdrChainState st;
  adr_half_setup(a, st.h[0], col_a, nx, ny, rstart);
  adr_half_setup(a, st.h[1], col_b, nx, ny, rstart);
  st.sx_issue = st.sy_issue = st.sx_use = st.sy_use = 0;
  st.ir       = rstart;

  double2 Wa[K][3], Wb[K][3];
#pragma unroll
  for (int l = 0; l < K; l++)
    Wa[l][0] = Wa[l][1] = Wa[l][2] = Wb[l][0] = Wb[l][1] = Wb[l][2] = make_double2(0.0, 0.0);
  // canonical layout at phase 0: index 0 oldest (about to be overwritten), 1 = south row, 2 = centre row
  Wa[0][1] = ld_keep2(adr_row_ptr(a.x, rstart - 1, st.h[0].col, nx, ny));
  Wa[0][2] = ld_keep2(adr_row_ptr(a.x, rstart, st.h[0].col, nx, ny));
  Wb[0][1] = ld_keep2(adr_row_ptr(a.x, rstart - 1, st.h[1].col, nx, ny));
  Wb[0][2] = ld_keep2(adr_row_ptr(a.x, rstart, st.h[1].col, nx, ny));

#pragma unroll
  for (int q = 0; q < PF; q++) adr_chain_issue<K, PF>(a, st, rx, rp, ry, rf, nx, ny, true);

#define AROW(PH, CHECK, R1) \
  adr_chain_row<K, PF, PH, CHECK>(a, st, Wa, Wb, rx, rp, ry, rf, nx, smask_a, smask_b, lane, R1, j0, j1, ny, (R1) + PF < rend)

  // phases as in k_chain_march: checked warm-up in whole triples, unchecked steady state, checked drain
  const int total3 = ((rend - rstart + 2) / 3) * 3;
  int warm         = 2 * (K - 1);
  warm             = ((warm + 2) / 3) * 3;
  // (as in k_chain_march: the steady state stays clear of the rows whose next group touches row ny, so that it carries
  // no wrap-around logic)
  const int lim    = (j1 < ny - (PF + 2)) ? j1 : ny - (PF + 2);
  int steady       = (lim - (rstart + warm)) / 3 * 3;
  if (steady < 0) steady = 0;
  int r1 = rstart;
#pragma unroll 1
  for (; r1 < rstart + warm && r1 < rstart + total3; r1 += 3)
  {
    AROW(0, true, r1);
    AROW(1, true, r1 + 1);
    AROW(2, true, r1 + 2);
  }
  const int s1 = r1 + steady;
#pragma unroll 1
  for (; r1 < s1; r1 += 3)
  {
    AROW(0, false, r1);
    AROW(1, false, r1 + 1);
    AROW(2, false, r1 + 2);
  }
#pragma unroll 1
  for (; r1 < rstart + total3; r1 += 3)
  {
    AROW(0, true, r1);
    AROW(1, true, r1 + 1);
    AROW(2, true, r1 + 2);
  }
  cp_async_wait<0>();
#undef AROW
}

// ---- launch geometry (host side)
static inline size_t adr_chain_smem(int K, int PF) { return (size_t)(4 * (PF + 1) + 4 * (PF + K)) * kAdrChainThreads * sizeof(double2); }
static inline bool adr_chain_supported(int64_t nx, int64_t ny, int K) { return K >= 2 && K <= B200_MAX_CHAIN && nx >= 64 && ny >= 16; }
static inline dim3 adr_chain_grid(int64_t nx, int64_t ny, int K, int rows)
{
  const int use = 64 - 2 * K;
  int64_t warps = (nx + use - 1) / use;
  int64_t gx    = (warps + kAdrChainThreads / 32 - 1) / (kAdrChainThreads / 32);
  return dim3((unsigned)gx, (unsigned)((ny + rows - 1) / rows));
}
