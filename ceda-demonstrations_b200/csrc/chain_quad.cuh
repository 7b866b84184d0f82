// chain_quad.cuh -- k_chain_quad: K temporally blocked STS stages per launch, FOUR cells per thread.
//
// Same computation, tiling idea and bit-for-bit arithmetic as k_chain_march (chain_march.cuh); what
// changes is the width of a thread.  A warp owns a 128-cell window made of two 64-cell halves; lane i
// owns cells {2i, 2i+1} of the left half ("a") and cells {64+2i, 64+2i+1} of the right half ("b"):
//   * every global access is, per half, exactly k_chain_march's: 16 bytes per lane, 512 contiguous
//     bytes per warp instruction (a first version gave each lane four ADJACENT cells: its two 16-byte
//     accesses then touched half of each 32-byte sector and doubled the L2 tag requests -- ncu:
//     lts__t_tag_requests at 66 % with DRAM at 52 %; profiles/r01_quad_v1_*);
//   * the halo is only at the two ends of the 128-cell window: 2*ceil(K/2) cells per side, so for
//     K = 4, 120 of 128 cells are stored instead of 56 of 64;
//   * the per-row and per-level bookkeeping (ring slots, row pointers, y-coefficient loads, store
//     addresses, loop control) is paid once for four cells instead of two;
//   * the two halves are independent dependency chains, which doubles the FP64 work in flight per
//     warp (8 warps per SM: 2 blocks x 128 threads; the thread-private cp.async ring limits residency).
// West/east neighbours come from rotating warp shuffles; the seam between the halves (lane 31 of
// "a" <-> lane 0 of "b") is one select per side.  Shape requirements are k_chain_march's: nx even,
// nx >= 128, ny >= 16; halo flavour: g2 even >= 2*ceil(K/2).
#pragma once
#include "chain_march.cuh"

static const int kQuadThreads = 128;
static const int kQuadPF      = 3;

// per-half source state: where this lane's two cells of a row come from
struct QuadHalf
{
  const double *px, *pp, *py, *pf; // next group: x at row ir+1 ; prev2 / yn / fn at row ir
  int64_t pstep;                   // row stride of the source (nx, or g2 in a W/E halo strip)
  int64_t lane_col;                // column (halo mode: strip offset + column) of the first cell
  int64_t soff;                    // r1*nx + column (unwrapped row; valid whenever a store can happen)
  bool we;                         // halo mode: this half of the lane reads the W/E halo strips
};

struct QuadState
{
  QuadHalf h[2];
  int ir;                                 // unwrapped row of the next group
  int sx_issue, sy_issue, sx_use, sy_use; // ring slots
  int trow;                               // index of row r1 in the y-coefficient table
};

template <bool HALO>
__device__ __forceinline__ void quad_half_advance(const ChainArgs& a, QuadHalf& q, int r, int64_t nx, int ny)
{
  if (r == 0 || r == ny)
  {
    q.pp = row_ptr<HALO>(a.prev2, a.hp, r, q.we, q.lane_col, nx, ny, a.g, a.g2);
    q.py = row_ptr<HALO>(a.yn, a.hy, r, q.we, q.lane_col, nx, ny, a.g, a.g2);
    q.pf = row_ptr<HALO>(a.fn, a.hf, r, q.we, q.lane_col, nx, ny, a.g, a.g2);
  }
  else { q.pp += q.pstep; q.py += q.pstep; q.pf += q.pstep; }
  if (r + 1 == 0 || r + 1 == ny) q.px = row_ptr<HALO>(a.x, a.hx, r + 1, q.we, q.lane_col, nx, ny, a.g, a.g2);
  else q.px += q.pstep;
}

// ring addressing: slot s, half h of this thread = base[(2*s + h) * kQuadThreads]
template <int K, int PF, bool HALO>
__device__ __forceinline__ void quad_issue(const ChainArgs& a, QuadState& st, double2* rx, double2* rp,
                                           double2* ry, double2* rf, int64_t nx, int ny, bool issue)
{
  constexpr int DX = PF + 1, DY = PF + K;
  if (issue)
  {
    double2* sx = rx + 2 * st.sx_issue * kQuadThreads;
    double2* sp = rp + 2 * st.sx_issue * kQuadThreads;
    double2* sy = ry + 2 * st.sy_issue * kQuadThreads;
    double2* sf = rf + 2 * st.sy_issue * kQuadThreads;
    cp_async16(sx, st.h[0].px);           cp_async16(sx + kQuadThreads, st.h[1].px);
    cp_async16(sp, st.h[0].pp);           cp_async16(sp + kQuadThreads, st.h[1].pp);
    cp_async16(sy, st.h[0].py);           cp_async16(sy + kQuadThreads, st.h[1].py);
    cp_async16(sf, st.h[0].pf);           cp_async16(sf + kQuadThreads, st.h[1].pf);
  }
  cp_async_commit();
  st.sx_issue = (st.sx_issue + 1 == DX) ? 0 : st.sx_issue + 1;
  st.sy_issue = (st.sy_issue + 1 == DY) ? 0 : st.sy_issue + 1;
  const int r = ++st.ir;
  quad_half_advance<HALO>(a, st.h[0], r, nx, ny);
  quad_half_advance<HALO>(a, st.h[1], r, nx, ny);
}

// x-direction face coefficients of this lane's four columns and their sums (diffusion.cpp:48)
struct QuadXCoef
{
  double2 cwa, cwb, cea, ceb, sxa, sxb;
};

// one cell of one level: diffusion.cpp:48-53 and the stage combination, in k_stage_march's order
template <bool FMA>
__device__ __forceinline__ double quad_cell(double sx, double sy, double cw, double ce, double dys, double dyn,
                                            double uc, double uw, double ue, double um, double up,
                                            const double* cf, double p2, double yv, double fv)
{
  double L = DMUL(-DADD(sx, sy), uc);
  L        = mad<FMA>(cw, uw, L);
  L        = mad<FMA>(ce, ue, L);
  L        = mad<FMA>(dys, um, L);
  L        = mad<FMA>(dyn, up, L);
  // (the reference's "f = 0; f += L" is dropped as in k_chain_march: it can only change the sign of an
  // exactly-zero L, which is never stored)
  double z = DMUL(cf[0], L);
  z        = mad<FMA>(cf[1], p2, z);
  z        = mad<FMA>(cf[2], yv, z);
  z        = mad<FMA>(cf[3], uc, z);
  z        = mad<FMA>(cf[4], fv, z);
  return z;
}

template <int K, int PF, int PH, bool CHECK, bool HALO, bool FMA>
__device__ __forceinline__ void quad_row(const ChainArgs& a, QuadState& st, double2 (&Wa)[K][3], double2 (&Wb)[K][3],
                                         double2* rx, double2* rp, double2* ry, double2* rf,
                                         const double2* ytab, const double* stab, int64_t nx, int ny,
                                         const QuadXCoef& xc, unsigned smask_a, unsigned smask_b, int lane,
                                         int r1, int j0, int j1, bool issue)
{
  constexpr int DX = PF + 1, DY = PF + K;
  constexpr int IO = PH % 3, IM = (PH + 1) % 3, IC = (PH + 2) % 3; // oldest (overwritten), then um, uc ; up = IO
  quad_issue<K, PF, HALO>(a, st, rx, rp, ry, rf, nx, ny, issue); // group(r1 + PF)
  cp_async_wait<PF>(); // all but the PF newest groups have landed: group(r1) is ready

  Wa[0][IO]        = rx[(2 * st.sx_use) * kQuadThreads]; // x row r1+1 replaces the oldest row
  Wb[0][IO]        = rx[(2 * st.sx_use + 1) * kQuadThreads];
  const double2 Pa = rp[(2 * st.sx_use) * kQuadThreads];
  const double2 Pb = rp[(2 * st.sx_use + 1) * kQuadThreads];
  int64_t soa = st.h[0].soff, sob = st.h[1].soff;
  const int lw = (lane + 31) & 31, le = (lane + 1) & 31;
#pragma unroll
  for (int l = 1; l <= K; l++)
  {
    const double2 dy = ytab[st.trow - (l - 1)]; // (Dy_s, Dy_n) of row r1-(l-1)
    const double sy  = stab[st.trow - (l - 1)]; // Dy_s + Dy_n
    const double2 uma = Wa[l - 1][IM], uca = Wa[l - 1][IC], upa = Wa[l - 1][IO];
    const double2 umb = Wb[l - 1][IM], ucb = Wb[l - 1][IC], upb = Wb[l - 1][IO];
    // rotating shuffles; the seam: west of half b's lane 0 is half a's lane 31, east of half a's
    // lane 31 is half b's lane 0.  (West of a's lane 0 / east of b's lane 31 lie outside the
    // window: whatever arrives there only reaches halo cells.)
    const double wa = __shfl_sync(0xffffffffu, uca.y, lw);
    const double wb = __shfl_sync(0xffffffffu, ucb.y, lw);
    const double ea = __shfl_sync(0xffffffffu, uca.x, le);
    const double eb = __shfl_sync(0xffffffffu, ucb.x, le);
    const double uw_a = wa, uw_b = (lane == 0) ? wa : wb;
    const double ue_a = (lane == 31) ? eb : ea, ue_b = eb;
    // z_{l-2} at this row: prev2 for the first stage, else the oldest row of level l-2's window
    const double2 p2a = (l == 1) ? Pa : Wa[(l >= 2) ? l - 2 : 0][IM];
    const double2 p2b = (l == 1) ? Pb : Wb[(l >= 2) ? l - 2 : 0][IM];
    int sl = st.sy_use - (l - 1); // yn / fn of row r1-(l-1)
    if (sl < 0) sl += DY;
    const double2 yva = ry[(2 * sl) * kQuadThreads], yvb = ry[(2 * sl + 1) * kQuadThreads];
    const double2 fva = rf[(2 * sl) * kQuadThreads], fvb = rf[(2 * sl + 1) * kQuadThreads];
    const double* cf = a.c[l - 1];
    double2 za, zb;
    za.x = quad_cell<FMA>(xc.sxa.x, sy, xc.cwa.x, xc.cea.x, dy.x, dy.y, uca.x, uw_a, uca.y, uma.x, upa.x, cf, p2a.x, yva.x, fva.x);
    za.y = quad_cell<FMA>(xc.sxa.y, sy, xc.cwa.y, xc.cea.y, dy.x, dy.y, uca.y, uca.x, ue_a, uma.y, upa.y, cf, p2a.y, yva.y, fva.y);
    zb.x = quad_cell<FMA>(xc.sxb.x, sy, xc.cwb.x, xc.ceb.x, dy.x, dy.y, ucb.x, uw_b, ucb.y, umb.x, upb.x, cf, p2b.x, yvb.x, fvb.x);
    zb.y = quad_cell<FMA>(xc.sxb.y, sy, xc.cwb.y, xc.ceb.y, dy.x, dy.y, ucb.y, ucb.x, ue_b, umb.y, upb.y, cf, p2b.y, yvb.y, fvb.y);
    bool row_ok = true;
    if (CHECK)
    {
      const int rl = r1 - (l - 1);
      row_ok       = rl >= j0 && rl < j1;
    }
    if (row_ok && ((smask_a >> (l - 1)) & 1u)) *reinterpret_cast<double2*>(a.out[l - 1] + soa) = za;
    if (row_ok && ((smask_b >> (l - 1)) & 1u)) *reinterpret_cast<double2*>(a.out[l - 1] + sob) = zb;
    soa -= nx;
    sob -= nx;
    if (l < K) { Wa[l][IO] = za; Wb[l][IO] = zb; } // newest row of level l replaces its oldest
  }
  st.h[0].soff += nx;
  st.h[1].soff += nx;
  st.trow += 1;
  st.sx_use = (st.sx_use + 1 == DX) ? 0 : st.sx_use + 1;
  st.sy_use = (st.sy_use + 1 == DY) ? 0 : st.sy_use + 1;
}

// where one half of a lane reads and stores: col_u = unwrapped column of its first cell
template <bool HALO>
__device__ __forceinline__ void quad_half_setup(const ChainArgs& a, QuadHalf& q, int64_t col_u, int64_t nx, int ny,
                                                int rstart, int64_t* xi)
{
  int64_t ic = col_u; // column used for stores and (wrap mode) loads
  *xi        = col_u; // column index into the x-direction coefficient tables
  q.we       = false;
  q.pstep    = nx;
  if (HALO)
  { // columns outside [0, nx) come from the W / E halo strips; beyond the strips: clamp (never used)
    const int64_t strip = (int64_t)(ny + 2 * a.g) * a.g2;
    if (col_u < 0)
    {
      int64_t c = col_u + a.g2;
      if (c < 0) { c = 0; *xi = -(int64_t)a.g2; }
      q.we       = true;
      q.lane_col = 2 * a.g * nx + c;
    }
    else if (col_u >= nx)
    {
      int64_t c = col_u - nx;
      if (c > a.g2 - 2) { c = a.g2 - 2; *xi = nx + c; }
      q.we       = true;
      q.lane_col = 2 * a.g * nx + strip + c;
    }
    else q.lane_col = col_u;
    if (q.we) q.pstep = a.g2;
  }
  else
  {
    if (ic < 0) ic += nx;
    else if (ic >= nx) ic -= nx;
    *xi        = ic;
    q.lane_col = ic;
  }
  q.soff = (int64_t)rstart * nx + ic;
  q.px   = row_ptr<HALO>(a.x, a.hx, rstart + 1, q.we, q.lane_col, nx, ny, a.g, a.g2);
  q.pp   = row_ptr<HALO>(a.prev2, a.hp, rstart, q.we, q.lane_col, nx, ny, a.g, a.g2);
  q.py   = row_ptr<HALO>(a.yn, a.hy, rstart, q.we, q.lane_col, nx, ny, a.g, a.g2);
  q.pf   = row_ptr<HALO>(a.fn, a.hf, rstart, q.we, q.lane_col, nx, ny, a.g, a.g2);
}

// MINB = resident blocks per SM the register allocation is held to.  (Holding K <= 4 to 3 blocks,
// 168 registers and PF = 2, was measured: no faster, small spills -- 8 warps per SM are enough.)
template <int K, int PF, bool HALO, bool FMA, int MINB = 2>
__global__ void __launch_bounds__(kQuadThreads, MINB) k_chain_quad(const ChainArgs a)
{
  constexpr int HC   = (K + 1) / 2;   // halo lanes at each end of the window (2 cells each): 2*HC >= K
  constexpr int WUSE = 128 - 4 * HC;  // cells a warp stores per row
  constexpr int DX   = PF + 1;        // ring depth of x and prev2
  constexpr int DY   = PF + K;        // ring depth of yn and fn
  B200_DYN_SMEM(double2, ring);
  double2* rx   = ring + threadIdx.x;                                            // [DX][2][threads]
  double2* rp   = ring + (size_t)2 * DX * kQuadThreads + threadIdx.x;            // [DX][2][threads]
  double2* ry   = ring + (size_t)4 * DX * kQuadThreads + threadIdx.x;            // [DY][2][threads]
  double2* rf   = ring + (size_t)(4 * DX + 2 * DY) * kQuadThreads + threadIdx.x; // [DY][2][threads]
  double2* ytab = ring + (size_t)(4 * DX + 4 * DY) * kQuadThreads;               // [rows + 3(K-1) + 2]
  double* stab  = reinterpret_cast<double*>(ytab + (a.rows + 3 * (K - 1) + 2));  // [rows + 3(K-1) + 2]

  const int lane   = threadIdx.x & 31;
  const int64_t nx = a.nx;
  const int ny     = (int)a.ny;
  const int j0     = (int)blockIdx.y * a.rows;
  int j1           = j0 + a.rows;
  if (j1 > ny) j1 = ny;
  const int rstart = j0 - (K - 1), rend = j1 + (K - 1); // level-1 rows [rstart, rend)
  // y-direction face coefficients of rows rstart-(K-1) .. rend+1 (table index 0 = row rstart-(K-1))
  for (int t = threadIdx.x; t < (rend - rstart) + (K - 1) + 2; t += kQuadThreads)
  { // HALO: the tables are extended by the caller (global periodic index), negative rows are valid
    const int r  = rstart - (K - 1) + t;
    const int rw = HALO ? r : (r < 0 ? r + ny : (r >= ny ? r - ny : r));
    const double ds = a.cys[rw], dn = a.cyn[rw];
    ytab[t]         = make_double2(ds, dn);
    stab[t]         = DADD(ds, dn); // diffusion.cpp:48: (Dys + Dyn)
  }
  __syncthreads();

  const int64_t wg = (int64_t)blockIdx.x * (kQuadThreads / 32) + (threadIdx.x >> 5);
  if (wg * WUSE >= nx) return; // window entirely outside the field (no block-level sync below)
  const int64_t col_a = wg * WUSE - 2 * HC + 2 * lane; // unwrapped column of my first cell, left half
  const int64_t col_b = col_a + 64;                    // right half
  const bool ok_a     = (lane >= HC) && (col_a < nx);
  const bool ok_b     = (lane < 32 - HC) && (col_b < nx);
  unsigned smask_a = 0, smask_b = 0;
#pragma unroll
  for (int l = 0; l < K; l++)
    if (a.out[l])
    {
      if (ok_a) smask_a |= 1u << l;
      if (ok_b) smask_b |= 1u << l;
    }

  QuadState st;
  int64_t xia, xib;
  quad_half_setup<HALO>(a, st.h[0], col_a, nx, ny, rstart, &xia);
  quad_half_setup<HALO>(a, st.h[1], col_b, nx, ny, rstart, &xib);
  QuadXCoef xc;
  xc.cwa = ld_keep2(a.cxw + xia); xc.cwb = ld_keep2(a.cxw + xib);
  xc.cea = ld_keep2(a.cxe + xia); xc.ceb = ld_keep2(a.cxe + xib);
  xc.sxa = make_double2(DADD(xc.cwa.x, xc.cea.x), DADD(xc.cwa.y, xc.cea.y));
  xc.sxb = make_double2(DADD(xc.cwb.x, xc.ceb.x), DADD(xc.cwb.y, xc.ceb.y));

  st.sx_issue = st.sy_issue = st.sx_use = st.sy_use = 0;
  st.trow     = K - 1;
  st.ir       = rstart;

  double2 Wa[K][3], Wb[K][3];
#pragma unroll
  for (int l = 0; l < K; l++)
    Wa[l][0] = Wa[l][1] = Wa[l][2] = Wb[l][0] = Wb[l][1] = Wb[l][2] = make_double2(0.0, 0.0);
  // canonical layout at phase 0: index 0 oldest (about to be overwritten), 1 = um, 2 = uc
  Wa[0][1] = ld_keep2(row_ptr<HALO>(a.x, a.hx, rstart - 1, st.h[0].we, st.h[0].lane_col, nx, ny, a.g, a.g2));
  Wa[0][2] = ld_keep2(row_ptr<HALO>(a.x, a.hx, rstart, st.h[0].we, st.h[0].lane_col, nx, ny, a.g, a.g2));
  Wb[0][1] = ld_keep2(row_ptr<HALO>(a.x, a.hx, rstart - 1, st.h[1].we, st.h[1].lane_col, nx, ny, a.g, a.g2));
  Wb[0][2] = ld_keep2(row_ptr<HALO>(a.x, a.hx, rstart, st.h[1].we, st.h[1].lane_col, nx, ny, a.g, a.g2));

  // prologue of the pipeline: groups rstart .. rstart+PF-1
#pragma unroll
  for (int q = 0; q < PF; q++) quad_issue<K, PF, HALO>(a, st, rx, rp, ry, rf, nx, ny, true);

#define QROW(PH, CHECK, R1) \
  quad_row<K, PF, PH, CHECK, HALO, FMA>(a, st, Wa, Wb, rx, rp, ry, rf, ytab, stab, nx, ny, xc, smask_a, smask_b, lane, R1, j0, j1, (R1) + PF < rend)

  // phases as in k_chain_march: checked warm-up in whole triples, unchecked steady state, checked
  // drain; the trip count is rounded up to a multiple of 3 (the extra rows compute values that are
  // never stored and load wrapped / halo rows that exist).
  const int total3 = ((rend - rstart + 2) / 3) * 3;
  int warm         = 2 * (K - 1);
  warm             = ((warm + 2) / 3) * 3;
  int steady       = (j1 - (rstart + warm)) / 3 * 3;
  if (steady < 0) steady = 0;
  int r1 = rstart;
#pragma unroll 1
  for (; r1 < rstart + warm && r1 < rstart + total3; r1 += 3)
  {
    QROW(0, true, r1);
    QROW(1, true, r1 + 1);
    QROW(2, true, r1 + 2);
  }
  const int s1 = r1 + steady;
#pragma unroll 1
  for (; r1 < s1; r1 += 3)
  {
    QROW(0, false, r1);
    QROW(1, false, r1 + 1);
    QROW(2, false, r1 + 2);
  }
#pragma unroll 1
  for (; r1 < rstart + total3; r1 += 3)
  {
    QROW(0, true, r1);
    QROW(1, true, r1 + 1);
    QROW(2, true, r1 + 2);
  }
  cp_async_wait<0>();
#undef QROW
}

// ---- launch geometry (host side)
static inline size_t chain_quad_smem(int K, int PF, int rows)
{
  return (size_t)(4 * (PF + 1) + 4 * (PF + K)) * kQuadThreads * sizeof(double2) +
         (size_t)(rows + 3 * (K - 1) + 2) * (sizeof(double2) + sizeof(double));
}
// halo_cols: deep-halo columns g2 of the halo flavour, or -1 for the periodic-wrap flavour
static inline bool chain_quad_supported(int64_t nx, int64_t ny, int K, int halo_cols)
{
  if (K < 2 || K > B200_MAX_CHAIN) return false;
  if ((nx & 1) || nx < 128 || ny < 16) return false;
  if (halo_cols >= 0 && ((halo_cols & 1) || halo_cols < 2 * ((K + 1) / 2))) return false;
  return true;
}
static inline dim3 chain_quad_grid(int64_t nx, int64_t ny, int K, int* rows)
{
  const int hc  = (K + 1) / 2;
  const int use = 128 - 4 * hc;
  int64_t warps = (nx + use - 1) / use;
  int64_t gx    = (warps + kQuadThreads / 32 - 1) / (kQuadThreads / 32);
  int64_t gy    = (ny + *rows - 1) / *rows;
  if (gy > 65535)
  {
    *rows = (int)((ny + 65534) / 65535);
    gy    = (ny + *rows - 1) / *rows;
  }
  return dim3((unsigned)gx, (unsigned)gy);
}
