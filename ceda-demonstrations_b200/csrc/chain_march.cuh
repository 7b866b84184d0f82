// chain_march.cuh -- k_chain_march: K temporally blocked STS stages per launch, two cells per thread.
//
// Included by b200_kernels.cu (nvcc, sm_100a) and -- with B200_HOST_EMU defined -- by the host
// emulation harness tests/emu (g++), which runs the same source lane by lane on CPU threads so the
// indexing, tiling, ring and halo logic is tested without a GPU.  The harness is test
// infrastructure: nothing in the product includes this file with B200_HOST_EMU.
#pragma once
#include "kernel_prims.cuh"

// ------------------------------------------- temporally blocked STS stages
// K consecutive RKC/RKL stages (arkode_lsrkstep.c:674-750 / :960-1050) in ONE pass:
//   z_1 = c1[0] L(x)   + c1[1] p   + c1[2] yn + c1[3] x   + c1[4] fn      (x = z_{j-1}, p = z_{j-2})
//   z_2 = c2[0] L(z_1) + c2[1] x   + c2[2] yn + c2[3] z_1 + c2[4] fn
//   z_l = cl[0] L(z_{l-1}) + cl[1] z_{l-2} + cl[2] yn + cl[3] z_{l-1} + cl[4] fn
// Every cell value is produced by exactly the instruction sequence of the one-stage kernel, so
// the result is bit-identical; only the traffic changes: 4 streamed reads + (usually) 2 writes
// per K cell-updates instead of per one (48/K bytes instead of 40).
//
// Overlapped tiling, no block-level sync: a warp owns a 64-cell window of which the outer HL
// lanes on each side are halo (level l is valid on cells [l, 63-l] of the window; halo lanes
// never store); a block marches down `rows` output rows and starts K-1 rows early.  Level l lags
// level l-1 by one row; each level keeps a 3-row window of the level below in registers and gets
// west/east neighbours by warp shuffle.  One periodic rank (index wrap) only.
struct ChainArgs
{
  int64_t nx, ny;
  const double *cxw, *cxe, *cys, *cyn;
  const double* x;
  const double* prev2;
  const double* yn;
  const double* fn;
  double c[B200_MAX_CHAIN][5];
  double* out[B200_MAX_CHAIN];
  double* f_out; // HEAD flavour: where L(x) = f(t_n, y_n) goes (the fn of all later stages); NULL = it is stored already
  int head;      // the chain begins the step (HEAD flavour)
  int rows;
  // multi-rank (HALO = true): per-operand deep-halo buffers, layout of b200_deep_halo_exchange
  const double *hx, *hp, *hy, *hf;
  int g, g2; // halo depth in rows / in columns (g >= K, g2 even >= 2*ceil(K/2))
  // uniform coefficients (b200_stencil_geom.uniform): the four face coefficients and
  // u_ndc = -((cxw + cxe) + (cys + cyn)), summed on the host in the reference's order (IEEE add)
  double u_cxw, u_cxe, u_cys, u_cyn, u_ndc;
};

static const int kChainThreads = 256;

// address of (row r, this lane's column), r in [-g, ny+g):
//   wrap mode: rows outside [0, ny) wrap periodically onto the field itself;
//   halo mode: they come from the field's deep halo
//     halo = [ S: g rows x nx | N: g rows x nx | W: (ny+2g) rows x g2 | E: (ny+2g) rows x g2 ]
//     (S = rows -g..-1, N = rows ny..ny+g-1, W / E = columns -g2..-1 / nx..nx+g2-1 of rows
//     -g..ny+g-1); we = this lane lies in a W/E strip (lane_col then includes the strip offset).
template <bool HALO>
__device__ __forceinline__ const double* row_ptr(const double* field, const double* halo, int r, bool we,
                                                 int64_t lane_col, int64_t nx, int ny, int g, int g2)
{
  if (!HALO)
  {
    const int rw = (r < 0) ? r + ny : ((r >= ny) ? r - ny : r);
    return field + (int64_t)rw * nx + lane_col;
  }
  if (we) return halo + lane_col + (int64_t)(r + g) * g2;
  if (r >= 0 && r < ny) return field + (int64_t)r * nx + lane_col;
  const int hr = (r < 0) ? r + g : g + (r - ny);
  return halo + (int64_t)hr * nx + lane_col;
}

// Operands are staged through a shared-memory ring PF rows ahead: the bytes in flight that keep HBM
// busy cost no registers, and the FP64 pipe works on row r while rows r+1..r+PF stream in.  Every
// thread reads back its own 16 bytes of a ring row.  Two ways of filling it:
//   plain -- every thread copies its own 16 bytes with cp.async (LDGSTS), so cp.async.wait_group is
//            the only synchronisation in the row loop;
//   BULK  -- a warp's ring row is 512 contiguous bytes on both sides, so one elected lane issues one
//            cp.async.bulk (UBLKCP, the TMA unit's 1-D path) per operand and row from warp-uniform
//            pointers, completed on an mbarrier per warp and ring slot which all 32 lanes wait for
//            (see ChainState / chain_issue); the windows on the block's first and last columns,
//            which are not one contiguous run, stay with cp.async.  Ring depths: x and prev2 PF+1 rows, yn and fn PF+K rows
// (level l consumes yn/fn of row r-(l-1)).  The y-direction coefficients of the block's rows
// sit in a small shared table (one __syncthreads before the loop).
//
// The row loop is issue-bound once HBM is no longer the limit, so the steady state is unrolled
// by 3 with the 3-row register windows addressed by a compile-time phase (no rotation moves),
// carries no row-range predicates and no wrap-around / field-to-halo logic (the rows whose next
// group touches row 0 or row ny belong to the checked phases), and uses running offsets instead of
// index multiplies; the 2(K-1) warm-up rows and the drain rows run through the same body with
// CHECK = true.
// SPLIT flavour (round 2).  In the plain order level l + 1 consumes, as its newest row, what level l produced a moment
// ago in the SAME row step: the K levels of a step form one dependent chain of ~10 FP64 operations each, and with four
// warps per SM sub-partition the FP64 pipe (the busiest unit at the sustained clock) runs out of independent work
// whenever a warp or two are in their load / store phases.  SPLIT gives the upper half of the levels (H+1..K,
// H = ceil(K/2)) ONE extra row of lag and computes it FIRST in a row step, from the windows of levels H-1 and H as the
// PREVIOUS step left them: the two halves of a step are independent instruction streams (twice the ILP), the 3-row
// register windows still hold every row that is needed (the newest rows of levels H-1, H are simply read one step
// later), every cell is still produced by the same instruction sequence -> the same bits.  Costs one more yn / fn ring
// row and one more row step per block.  chain_lag(l) = rows level l lags behind level 1.
template <int K, bool SPLIT>
__host__ __device__ constexpr int chain_half() { return SPLIT ? (K + 1) / 2 : K; }
template <int K, bool SPLIT>
__host__ __device__ constexpr int chain_lag(int l) { return (l - 1) + ((SPLIT && l > chain_half<K, SPLIT>()) ? 1 : 0); }

// ring slot arithmetic: a depth that is a power of two costs one AND
template <int D>
__device__ __forceinline__ int ring_next(int s)
{
  if constexpr ((D & (D - 1)) == 0) return (s + 1) & (D - 1);
  else return (s + 1 == D) ? 0 : s + 1;
}
template <int D>
__device__ __forceinline__ int ring_back(int s, int lag)
{
  if constexpr ((D & (D - 1)) == 0) return (s - lag) & (D - 1);
  else
  {
    const int t = s - lag;
    return t < 0 ? t + D : t;
  }
}

struct ChainState
{
  int64_t soff;      // r1*nx + ic (unwrapped; valid whenever a store can happen)
  const double *px, *pp, *py, *pf; // next group: x at row ir+1 ; prev2 / yn / fn at row ir
  int64_t pstep;     // row stride of this lane's source (nx, or g2 in a W/E halo strip)
  int ir;            // unwrapped row of the next group
  int sx_issue, sy_issue, sx_use, sy_use; // ring slots
  int trow;          // index of row r1 in the y-coefficient table
  bool we;           // halo mode: this lane reads the W/E halo strips
  int64_t lane_col;  // column (halo mode: strip offset + column) of this lane
  // BULK flavour.  A warp whose 64-cell window is ONE contiguous run of the field (every window but those that touch
  // the first or last column of the block) has it copied by bulk copies: `bulk` is set and the source offsets below
  // belong to the WINDOW (the same in every lane -- they live in uniform registers); the edge windows keep the
  // per-thread cp.async of the plain flavour, with their addresses computed from (row, lane_col) when needed.
  // The four operands share one layout, and row ir+1 of this group is row ir of the next: ONE offset is computed per
  // row step (uox), the other is handed down (uoy); uhx / uhy say whether the row lies in the S / N halo (halo mode).
  bool bulk;
  int64_t wcol;     // first column of the warp's window
  int64_t uox, uoy; // element offset of the window in row ir+1 / row ir (from the field's or the halo's base)
  bool uhx, uhy;
  unsigned par;     // parity of the barrier phase the next wait is for
};

// issue group(ir) = { x row ir+1, prev2 / yn / fn row ir } into the ring and advance the running
// source pointers by one row; the pointers are recomputed only where the source changes
// (rows 0 and ny: wrap-around, or field <-> S/N halo)
// HEAD flavour (the chain starts with stage 1 of the step, z_1 = y_n + c L(y_n)): only x (= y_n) is streamed -- there is
// no z_{-1}, and f_n = L(y_n) is produced by level 1 of this very launch, which writes it into the fn ring itself.
// BULK flavour: the same group as ONE bulk copy per operand (cp.async.bulk, the TMA unit's 1-D path: 512 bytes, the
// warp's whole window), issued by one elected lane from warp-uniform pointers and completed on the warp's mbarrier of
// this ring slot; rxw .. rfw are the ring rows of the WARP (lane 0's slot).
// WRAP = false (the steady state of the row loop, see the kernel): neither row of the NEXT group is row 0 or row ny, so
// the running pointers just advance by one row.
template <int K, int PF, bool HALO, bool HEAD, bool SPLIT, bool BULK, bool WRAP = true>
__device__ __forceinline__ void chain_issue(const ChainArgs& a, ChainState& st, double2* rx, double2* rp,
                                            double2* ry, double2* rf, double2* rxw, double2* rpw, double2* ryw,
                                            double2* rfw, unsigned long long* bars, int64_t nx, int ny, bool issue)
{
  constexpr int DX = PF + 1, DY = PF + K + (SPLIT ? 1 : 0);
  if (BULK && st.bulk)
  {
    if (issue && elect_one())
    {
      unsigned long long* bar = bars + st.sx_issue;
      mbar_arrive_expect_tx(bar, 512u * (HEAD ? 2u : 4u));
      bulk_g2s(rxw + st.sx_issue * kChainThreads, ((HALO && st.uhx) ? a.hx : a.x) + st.uox, 512u, bar);
      if (!HEAD) bulk_g2s(rpw + st.sx_issue * kChainThreads, ((HALO && st.uhy) ? a.hp : a.prev2) + st.uoy, 512u, bar);
      bulk_g2s(ryw + st.sy_issue * kChainThreads, ((HALO && st.uhy) ? a.hy : a.yn) + st.uoy, 512u, bar);
      if (!HEAD) bulk_g2s(rfw + st.sy_issue * kChainThreads, ((HALO && st.uhy) ? a.hf : a.fn) + st.uoy, 512u, bar);
    }
    st.sx_issue = ring_next<DX>(st.sx_issue);
    st.sy_issue = ring_next<DY>(st.sy_issue);
    const int r = ++st.ir; // the next group: rows r and r + 1
    st.uoy = st.uox;
    st.uhy = st.uhx;
    if (WRAP && (r + 1 == 0 || r + 1 == ny))
    { // row 0 of the field (wrap mode: from either side; halo mode: coming out of the S halo), or into the N halo
      st.uhx = HALO && (r + 1 == ny);
      st.uox = st.wcol + (st.uhx ? (int64_t)a.g * nx : 0);
    }
    else st.uox += nx;
    return;
  }
  if (BULK)
  { // edge window of the BULK flavour: per-thread cp.async, addresses from (row, lane_col) -- no running pointers
    if (issue)
    {
      const int r = st.ir;
      cp_async16(rx + st.sx_issue * kChainThreads, row_ptr<HALO>(a.x, a.hx, r + 1, st.we, st.lane_col, nx, ny, a.g, a.g2));
      if (!HEAD) cp_async16(rp + st.sx_issue * kChainThreads, row_ptr<HALO>(a.prev2, a.hp, r, st.we, st.lane_col, nx, ny, a.g, a.g2));
      cp_async16(ry + st.sy_issue * kChainThreads, row_ptr<HALO>(a.yn, a.hy, r, st.we, st.lane_col, nx, ny, a.g, a.g2));
      if (!HEAD) cp_async16(rf + st.sy_issue * kChainThreads, row_ptr<HALO>(a.fn, a.hf, r, st.we, st.lane_col, nx, ny, a.g, a.g2));
    }
    cp_async_commit();
    st.sx_issue = ring_next<DX>(st.sx_issue);
    st.sy_issue = ring_next<DY>(st.sy_issue);
    ++st.ir;
    return;
  }
  if (issue)
  {
    cp_async16(rx + st.sx_issue * kChainThreads, st.px);
    if (!HEAD) cp_async16(rp + st.sx_issue * kChainThreads, st.pp);
    cp_async16(ry + st.sy_issue * kChainThreads, st.py);
    if (!HEAD) cp_async16(rf + st.sy_issue * kChainThreads, st.pf);
  }
  cp_async_commit();
  st.sx_issue = ring_next<DX>(st.sx_issue);
  st.sy_issue = ring_next<DY>(st.sy_issue);
  const int r = ++st.ir;
  if (WRAP && (r == 0 || r == ny))
  {
    if (!HEAD) st.pp = row_ptr<HALO>(a.prev2, a.hp, r, st.we, st.lane_col, nx, ny, a.g, a.g2);
    st.py = row_ptr<HALO>(a.yn, a.hy, r, st.we, st.lane_col, nx, ny, a.g, a.g2);
    if (!HEAD) st.pf = row_ptr<HALO>(a.fn, a.hf, r, st.we, st.lane_col, nx, ny, a.g, a.g2);
  }
  else
  {
    st.py += st.pstep;
    if (!HEAD) { st.pp += st.pstep; st.pf += st.pstep; }
  }
  if (WRAP && (r + 1 == 0 || r + 1 == ny)) st.px = row_ptr<HALO>(a.x, a.hx, r + 1, st.we, st.lane_col, nx, ny, a.g, a.g2);
  else st.px += st.pstep;
}

// one level of one row step (l = 1..K); BEFORE: the window of level l-1 has not been updated in this row step yet
// (SPLIT: level H+1 reads level H's rows one step late), so its three rows sit one slot further round
template <int K, int PF, int PH, int L, bool CHECK, bool HALO, bool FMA, bool UNI, bool HEAD, bool SPLIT>
__device__ __forceinline__ void chain_level(const ChainArgs& a, const ChainState& st, double2 (&W)[K][3], double2 P,
                                            double2* ry, double2* rf, const double2* ytab, const double* stab,
                                            int64_t nx, double2 cw, double2 ce, double sx0, double sx1,
                                            unsigned smask, int r1, int j0, int j1)
{
  constexpr int DY = PF + K + (SPLIT ? 1 : 0);
  constexpr int IO = PH % 3, IM = (PH + 1) % 3, IC = (PH + 2) % 3; // oldest (overwritten), then um, uc ; up = IO
  constexpr int H   = chain_half<K, SPLIT>();
  constexpr int LAG = chain_lag<K, SPLIT>(L);
  constexpr bool BEFORE  = SPLIT && (L == H + 1);                 // level L-1 is updated AFTER this level in a row step
  constexpr bool P2_LATE = SPLIT && (L == H + 1 || L == H + 2);   // level L-2 likewise
  // UNI: the coefficients are kernel parameters (constant bank): no table loads, and the centre
  // coefficient -((Dxw+Dxe)+(Dys+Dyn)) is one number for the whole field
  const double2 dy = UNI ? make_double2(a.u_cys, a.u_cyn) : ytab[st.trow - LAG]; // (Dy_s, Dy_n) of row r1-LAG
  const double sy  = UNI ? 0.0 : stab[st.trow - LAG]; // Dy_s + Dy_n, summed once per block when the table is filled
  const double2 um = W[L - 1][BEFORE ? IO : IM], uc = W[L - 1][BEFORE ? IM : IC], up = W[L - 1][BEFORE ? IC : IO];
  // diffusion.cpp:48-53, same association as k_stage_march
  double L0 = DMUL(UNI ? a.u_ndc : -DADD(sx0, sy), uc.x);
  double L1 = DMUL(UNI ? a.u_ndc : -DADD(sx1, sy), uc.y);
#ifndef B200_NO_XSHARE // (A/B builds only)
  constexpr bool XSHARE = UNI && !FMA;
#else
  constexpr bool XSHARE = false;
#endif
  if constexpr (XSHARE)
  { // Uniform coefficients with Dx_w == Dx_e (the launcher checks the bits; it is one number, kx / dx^2, in the
    // reference): the product Dx * u of a cell is what BOTH its x-neighbours add, so each thread rounds the products
    // of its own two cells once and the neighbours' come through the shuffles -- two multiplies per level and thread
    // less, the same bits (identical operands), the sums in the reference's order (west, then east).
    const double pa = DMUL(cw.x, uc.x), pb = DMUL(cw.x, uc.y);
    const double pw = __shfl_up_sync(0xffffffffu, pb, 1);   // Dx_w * u(west of my first cell)
    const double pe = __shfl_down_sync(0xffffffffu, pa, 1); // Dx_e * u(east of my second cell)
    L0 = DADD(L0, pw); L1 = DADD(L1, pa);
    L0 = DADD(L0, pb); L1 = DADD(L1, pe);
  }
  else
  {
    const double uw0 = __shfl_up_sync(0xffffffffu, uc.y, 1);
    const double ue1 = __shfl_down_sync(0xffffffffu, uc.x, 1);
    L0 = mad<FMA>(cw.x, uw0, L0);  L1 = mad<FMA>(cw.y, uc.x, L1);
    L0 = mad<FMA>(ce.x, uc.y, L0); L1 = mad<FMA>(ce.y, ue1, L1);
  }
  L0 = mad<FMA>(dy.x, um.x, L0); L1 = mad<FMA>(dy.x, um.y, L1);
  L0 = mad<FMA>(dy.y, up.x, L0); L1 = mad<FMA>(dy.y, up.y, L1);
  // The reference's "f = 0; f += ..." (k_stage_march: DADD(0.0, L)) is dropped here: 0 + L differs from L only
  // for L = -0.0, and L is consumed by z = c0*L + ... below and never stored, so the only trace it could
  // leave is the sign of an exactly-zero z (all five terms zero) -- equal as a number, and the FP64 pipe is
  // what bounds this kernel.
  // z_{l-2} at this row: prev2 for the first stage, else the row of level l-2's window that holds it
  const double2 p2 = (L == 1) ? P : W[(L >= 2) ? L - 2 : 0][P2_LATE ? IO : IM];
  const int sl = ring_back<DY>(st.sy_use, LAG); // yn / fn of row r1-LAG
  const double* cf = a.c[L - 1];
  const int64_t so = st.soff - (int64_t)LAG * nx;
  double2 z;
  bool doit = (smask >> (L - 1)) & 1u;
  if (CHECK)
  {
    const int rl = r1 - LAG;
    doit         = doit && rl >= j0 && rl < j1;
  }
  if (HEAD && L == 1)
  { // stage 1 of the step, z_1 = 1*y_n + (h mu~_1)*f_n (N_VLinearSum, arkode_lsrkstep.c:640 / :930), in the order of
    // the one-stage kernel's PAT2(C,S): acc = 1*x ; acc += c*L -- and f_n = L(y_n) itself, with the reference's
    // "f = 0; f += ..." (0 + L: a -0 becomes +0, and unlike further down this value is stored)
    L0 = DADD(0.0, L0);
    L1 = DADD(0.0, L1);
    rf[sl * kChainThreads] = make_double2(L0, L1); // the later levels read f_n of the rows behind from this ring
    z.x = mad<FMA>(cf[0], L0, DMUL(1.0, uc.x));
    z.y = mad<FMA>(cf[0], L1, DMUL(1.0, uc.y));
    bool dof = (smask >> K) & 1u; // store_ok of this lane
    if (CHECK) dof = dof && r1 >= j0 && r1 < j1;
    if (dof) *reinterpret_cast<double2*>(a.f_out + so) = make_double2(L0, L1);
  }
  else
  {
    const double2 yv = ry[sl * kChainThreads], fv = rf[sl * kChainThreads];
    z.x = DMUL(cf[0], L0);               z.y = DMUL(cf[0], L1);
    z.x = mad<FMA>(cf[1], p2.x, z.x);  z.y = mad<FMA>(cf[1], p2.y, z.y);
    z.x = mad<FMA>(cf[2], yv.x, z.x);  z.y = mad<FMA>(cf[2], yv.y, z.y);
    z.x = mad<FMA>(cf[3], uc.x, z.x);  z.y = mad<FMA>(cf[3], uc.y, z.y);
    z.x = mad<FMA>(cf[4], fv.x, z.x);  z.y = mad<FMA>(cf[4], fv.y, z.y);
  }
  if (doit) *reinterpret_cast<double2*>(a.out[L - 1] + so) = z;
  if (L < K) W[L < K ? L : 0][IO] = z; // newest row of level l replaces its oldest
}

template <int K, int PF, int PH, int LO, int HI, bool CHECK, bool HALO, bool FMA, bool UNI, bool HEAD, bool SPLIT>
__device__ __forceinline__ void chain_levels(const ChainArgs& a, const ChainState& st, double2 (&W)[K][3], double2 P,
                                             double2* ry, double2* rf, const double2* ytab, const double* stab,
                                             int64_t nx, double2 cw, double2 ce, double sx0, double sx1,
                                             unsigned smask, int r1, int j0, int j1)
{
  if constexpr (LO <= HI)
  {
    chain_level<K, PF, PH, LO, CHECK, HALO, FMA, UNI, HEAD, SPLIT>(a, st, W, P, ry, rf, ytab, stab, nx, cw, ce, sx0, sx1,
                                                                   smask, r1, j0, j1);
    chain_levels<K, PF, PH, LO + 1, HI, CHECK, HALO, FMA, UNI, HEAD, SPLIT>(a, st, W, P, ry, rf, ytab, stab, nx, cw, ce,
                                                                            sx0, sx1, smask, r1, j0, j1);
  }
}

template <int K, int PF, int PH, bool CHECK, bool HALO, bool FMA, bool UNI, bool HEAD, bool SPLIT, bool BULK>
__device__ __forceinline__ void chain_row(const ChainArgs& a, ChainState& st, double2 (&W)[K][3],
                                          double2* rx, double2* rp, double2* ry, double2* rf, double2* rxw,
                                          double2* rpw, double2* ryw, double2* rfw, unsigned long long* bars,
                                          const double2* ytab, const double* stab, int64_t nx, int ny,
                                          double2 cw, double2 ce, double sx0, double sx1,
                                          unsigned smask, int r1, int j0, int j1, bool issue)
{
  constexpr int DX = PF + 1, DY = PF + K + (SPLIT ? 1 : 0);
  constexpr int IO = PH % 3;
  constexpr int H  = chain_half<K, SPLIT>();
  // BULK: the slots group(r1 + PF) goes into were last read in the previous row step, and every lane must have seen
  // that step's barrier phase before the barrier is armed again
  if (BULK && st.bulk) __syncwarp();
  chain_issue<K, PF, HALO, HEAD, SPLIT, BULK, CHECK>(a, st, rx, rp, ry, rf, rxw, rpw, ryw, rfw, bars, nx, ny, issue); // group(r1 + PF)
  if (BULK && st.bulk)
  { // group(r1) has landed when its barrier phase completes (rows at and beyond rend = j1 + K - 1 have no group)
    if (!CHECK || r1 < j1 + (K - 1)) mbar_wait(bars + st.sx_use, st.par);
  }
  else cp_async_wait<PF>(); // all but the PF newest groups have landed: group(r1) is ready

  const double2 xnew = rx[st.sx_use * kChainThreads]; // x row r1+1
  const double2 P    = HEAD ? make_double2(0.0, 0.0) : rp[st.sx_use * kChainThreads];
  // SPLIT: the upper half first, from the windows as the previous row step left them ...
  if constexpr (SPLIT)
    chain_levels<K, PF, PH, H + 1, K, CHECK, HALO, FMA, UNI, HEAD, SPLIT>(a, st, W, P, ry, rf, ytab, stab, nx, cw, ce, sx0,
                                                                         sx1, smask, r1, j0, j1);
  W[0][IO] = xnew; // x row r1+1 replaces the oldest row
  // ... then levels 1..H (all of them without SPLIT), each from the row the level below produced a moment ago
  chain_levels<K, PF, PH, 1, H, CHECK, HALO, FMA, UNI, HEAD, SPLIT>(a, st, W, P, ry, rf, ytab, stab, nx, cw, ce, sx0, sx1,
                                                                    smask, r1, j0, j1);
  st.soff += nx;
  st.trow += 1;
  st.sx_use = ring_next<DX>(st.sx_use);
  st.sy_use = ring_next<DY>(st.sy_use);
  if constexpr (BULK)
    if (st.sx_use == 0) st.par ^= 1u; // every barrier has been through one more phase
}

// BULK flavour: element offset of a window (first column w0, inside the field) in row r, r in [-g, ny+g), from the base
// of the field or -- *in_halo -- of its deep halo (layout at row_ptr)
template <bool HALO>
__device__ __forceinline__ int64_t chain_window_off(int r, int64_t w0, int64_t nx, int ny, int g, bool* in_halo)
{
  *in_halo = HALO && (r < 0 || r >= ny);
  if (!HALO) return (int64_t)((r < 0) ? r + ny : ((r >= ny) ? r - ny : r)) * nx + w0;
  if (r >= 0 && r < ny) return (int64_t)r * nx + w0;
  return (int64_t)((r < 0) ? r + g : g + (r - ny)) * nx + w0;
}

// where the cell at unwrapped column col_u of the block comes from: wrap mode -- the field itself, periodically;
// halo mode -- the field, or the W / E strips of the deep halo (beyond the strips: clamped, never used)
template <bool HALO>
__device__ __forceinline__ void chain_source(const ChainArgs& a, int64_t col_u, int64_t nx, int ny, bool* we,
                                             int64_t* lane_col, int64_t* pstep, int64_t* ic, int64_t* xc)
{
  *ic    = col_u;
  *xc    = col_u;
  *we    = false;
  *pstep = nx;
  if (HALO)
  {
    const int64_t strip = (int64_t)(ny + 2 * a.g) * a.g2;
    if (col_u < 0) { *we = true; *lane_col = 2 * a.g * nx + (col_u + a.g2); }
    else if (col_u >= nx)
    {
      int64_t c = col_u - nx;
      if (c > a.g2 - 2) { c = a.g2 - 2; *xc = nx + c; }
      *we       = true;
      *lane_col = 2 * a.g * nx + strip + c;
    }
    else *lane_col = col_u;
    if (*we) *pstep = a.g2;
  }
  else
  {
    if (*ic < 0) *ic += nx;
    else if (*ic >= nx) *ic -= nx;
    *xc       = *ic;
    *lane_col = *ic;
  }
}

template <int K, int PF, bool HALO, bool FMA, bool UNI = false, bool HEAD = false, bool SPLIT = false, bool BULK = false>
__global__ void __launch_bounds__(kChainThreads, 2) k_chain_march(const ChainArgs a)
{
  constexpr int HL   = (K + 1) / 2;  // halo lanes per side (2 cells each): 2*HL >= K
  constexpr int WUSE = 64 - 4 * HL;  // cells a warp stores per row
  constexpr int DX   = PF + 1;       // ring depth of x and prev2
  constexpr int XL   = SPLIT ? 1 : 0; // extra row of lag of the upper half of the levels
  constexpr int DY   = PF + K + XL;  // ring depth of yn and fn
  B200_DYN_SMEM(double2, ring);
  double2* rx   = ring + threadIdx.x;                                         // [DX][threads]
  double2* rp   = ring + (size_t)DX * kChainThreads + threadIdx.x;            // [DX][threads]
  double2* ry   = ring + (size_t)2 * DX * kChainThreads + threadIdx.x;        // [DY][threads]
  double2* rf   = ring + (size_t)(2 * DX + DY) * kChainThreads + threadIdx.x; // [DY][threads]
  double2* ytab = ring + (size_t)(2 * DX + 2 * DY) * kChainThreads;           // [rows + 3(K-1) + 2 + 2 XL]
  double* stab  = reinterpret_cast<double*>(ytab + (a.rows + 3 * (K - 1) + 2 + 2 * XL)); // [rows + 3(K-1) + 2 + 2 XL]
  unsigned long long* bars = nullptr; // BULK: one mbarrier per warp and x-ring slot, behind stab: [warps][DX]

  const int lane   = threadIdx.x & 31;
  const int64_t nx = a.nx;
  const int ny     = (int)a.ny;
  const int j0     = (int)blockIdx.y * a.rows;
  int j1           = j0 + a.rows;
  if (j1 > ny) j1 = ny;
  const int rstart = j0 - (K - 1), rend = j1 + (K - 1); // level-1 rows [rstart, rend)
#define WROW(r) ((r) < 0 ? (r) + ny : ((r) >= ny ? (r) - ny : (r)))
  // y-direction face coefficients of rows rstart-(K-1)-XL .. rend+1+XL (table index 0 = row rstart-(K-1)-XL)
  if (!UNI)
  {
    for (int t = threadIdx.x; t < (rend - rstart) + (K - 1) + 2 + 2 * XL; t += kChainThreads)
    { // HALO: the tables are extended by the caller (global periodic index), negative rows are valid
      const int rw = HALO ? (rstart - (K - 1) - XL + t) : WROW(rstart - (K - 1) - XL + t);
      const double ds = a.cys[rw], dn = a.cyn[rw];
      ytab[t]         = make_double2(ds, dn);
      stab[t]         = DADD(ds, dn); // diffusion.cpp:48: (Dys + Dyn)
    }
    __syncthreads();
  }

  const int64_t wg = (int64_t)blockIdx.x * (kChainThreads / 32) + (threadIdx.x >> 5);
  if (wg * WUSE >= nx) return; // window entirely outside the field (no block-level sync below)
  const int64_t col_u = wg * WUSE - 2 * HL + 2 * lane; // unwrapped column of my first cell
  const bool store_ok = (lane >= HL) && (lane < 32 - HL) && (col_u < nx);
  unsigned smask      = 0;
#pragma unroll
  for (int l = 0; l < K; l++)
    if (store_ok && a.out[l]) smask |= 1u << l;
  if (HEAD && store_ok && a.f_out) smask |= 1u << K; // f_out

  ChainState st;
  int64_t ic; // column used for stores and (wrap mode) loads
  int64_t xc; // column index into the x-direction coefficient tables
  chain_source<HALO>(a, col_u, nx, ny, &st.we, &st.lane_col, &st.pstep, &ic, &xc);
  const double2 cw = UNI ? make_double2(a.u_cxw, a.u_cxw) : ld_keep2(a.cxw + xc);
  const double2 ce = UNI ? make_double2(a.u_cxe, a.u_cxe) : ld_keep2(a.cxe + xc);
  const double sx0 = DADD(cw.x, ce.x), sx1 = DADD(cw.y, ce.y);

  st.soff     = (int64_t)rstart * nx + ic;
  st.sx_issue = st.sy_issue = st.sx_use = st.sy_use = 0;
  st.trow     = K - 1 + XL;
  st.ir       = rstart;

  double2 W[K][3];
#pragma unroll
  for (int l = 0; l < K; l++) W[l][0] = W[l][1] = W[l][2] = make_double2(0.0, 0.0);
  // canonical layout at phase 0: index 0 oldest (about to be overwritten), 1 = um, 2 = uc
  W[0][1] = ld_keep2(row_ptr<HALO>(a.x, a.hx, rstart - 1, st.we, st.lane_col, nx, ny, a.g, a.g2));
  W[0][2] = ld_keep2(row_ptr<HALO>(a.x, a.hx, rstart, st.we, st.lane_col, nx, ny, a.g, a.g2));

  st.bulk = false;
  st.par  = 0;
  st.wcol = 0;
  st.uox = st.uoy = 0;
  st.uhx = st.uhy = false;
  double2 *rxw = rx, *rpw = rp, *ryw = ry, *rfw = rf;
  if constexpr (BULK)
  { // the warp index once more, as a value the compiler knows to be the same in every lane: everything derived from
    // it (window column, source pointers, ring rows, barriers) stays in uniform registers
    const int wu     = warp_uniform(threadIdx.x >> 5);
    const int64_t w0 = ((int64_t)blockIdx.x * (kChainThreads / 32) + wu) * WUSE - 2 * HL;
    st.bulk          = (w0 >= 0) && (w0 + 64 <= nx);
    st.wcol          = w0;
    rxw  = ring + wu * 32;
    rpw  = rxw + (size_t)DX * kChainThreads;
    ryw  = rxw + (size_t)2 * DX * kChainThreads;
    rfw  = rxw + (size_t)(2 * DX + DY) * kChainThreads;
    bars = reinterpret_cast<unsigned long long*>(stab + (a.rows + 3 * (K - 1) + 2 + 2 * XL)) + wu * DX;
    if (st.bulk)
    {
      if (elect_one())
      {
#pragma unroll
        for (int q = 0; q < DX; q++) mbar_init(bars + q, 1u);
        mbar_fence_init();
      }
      __syncwarp();
      st.uoy = chain_window_off<HALO>(rstart, w0, nx, ny, a.g, &st.uhy);
      st.uox = chain_window_off<HALO>(rstart + 1, w0, nx, ny, a.g, &st.uhx);
    }
  }
  else
  {
    st.px = row_ptr<HALO>(a.x, a.hx, rstart + 1, st.we, st.lane_col, nx, ny, a.g, a.g2);
    st.py = row_ptr<HALO>(a.yn, a.hy, rstart, st.we, st.lane_col, nx, ny, a.g, a.g2);
    st.pp = st.pf = st.py; // (not streamed in the HEAD flavour)
    if (!HEAD)
    {
      st.pp = row_ptr<HALO>(a.prev2, a.hp, rstart, st.we, st.lane_col, nx, ny, a.g, a.g2);
      st.pf = row_ptr<HALO>(a.fn, a.hf, rstart, st.we, st.lane_col, nx, ny, a.g, a.g2);
    }
  }

  // prologue of the pipeline: groups rstart .. rstart+PF-1
#pragma unroll
  for (int q = 0; q < PF; q++) chain_issue<K, PF, HALO, HEAD, SPLIT, BULK>(a, st, rx, rp, ry, rf, rxw, rpw, ryw, rfw, bars, nx, ny, true);

#define ROW(PH, CHECK, R1) \
  chain_row<K, PF, PH, CHECK, HALO, FMA, UNI, HEAD, SPLIT, BULK>(a, st, W, rx, rp, ry, rf, rxw, rpw, ryw, rfw, bars, ytab, stab, nx, ny, cw, ce, sx0, sx1, smask, R1, j0, j1, (R1) + PF < rend)

  // phases: [rstart, s0) checked warm-up in whole triples, [s0, s1) unchecked steady state in
  // triples, [s1, rend3) checked drain; rend3 rounds the trip count up to a multiple of 3 (the
  // extra rows compute garbage that is never stored and load wrapped, in-range rows).  SPLIT: the last level lags
  // one row more, so the loop runs one row step longer (nothing is loaded for it: `rend` still bounds the issue).
  const int total3 = ((rend + XL - rstart + 2) / 3) * 3;
  int warm         = 2 * (K - 1) + XL;
  warm             = ((warm + 2) / 3) * 3;
  // (the steady state also stays clear of the rows whose NEXT group touches row ny -- the group issued in row step r1
  // is group(r1 + PF), the one after it holds rows r1 + PF + 1 and r1 + PF + 2 -- so that it needs no wrap-around /
  // field-to-halo logic at all: only the last block rows of the field hand a few more rows to the checked drain)
  const int lim    = (j1 < ny - (PF + 2)) ? j1 : ny - (PF + 2);
  int steady       = (lim - (rstart + warm)) / 3 * 3;
  if (steady < 0) steady = 0;
  int r1 = rstart;
#pragma unroll 1
  for (; r1 < rstart + warm && r1 < rstart + total3; r1 += 3)
  {
    ROW(0, true, r1);
    ROW(1, true, r1 + 1);
    ROW(2, true, r1 + 2);
  }
  const int s1 = r1 + steady;
#pragma unroll 1
  for (; r1 < s1; r1 += 3)
  {
    ROW(0, false, r1);
    ROW(1, false, r1 + 1);
    ROW(2, false, r1 + 2);
  }
#pragma unroll 1
  for (; r1 < rstart + total3; r1 += 3)
  {
    ROW(0, true, r1);
    ROW(1, true, r1 + 1);
    ROW(2, true, r1 + 2);
  }
  cp_async_wait<0>();
#undef ROW
#undef WROW
}


// ---- launch geometry (host side; shared by b200_kernels.cu and the emulation harness)
static inline size_t chain_march_smem(int K, int PF, int rows, bool split = false, bool bulk = false)
{
  const int xl = split ? 1 : 0;
  return (size_t)(2 * (PF + 1) + 2 * (PF + K + xl)) * kChainThreads * sizeof(double2) +
         (size_t)(rows + 3 * (K - 1) + 2 + 2 * xl) * (sizeof(double2) + sizeof(double)) +
         (bulk ? (size_t)(kChainThreads / 32) * (PF + 1) * sizeof(unsigned long long) : 0);
}
static inline int chain_march_pf(int K) { return K <= 3 ? 4 : 3; } // prefetch depth instantiated per K
// rows may be raised so that grid.y fits 65535
static inline dim3 chain_march_grid(int64_t nx, int64_t ny, int K, int* rows)
{
  const int hl  = (K + 1) / 2;
  const int use = 64 - 4 * hl;
  int64_t warps = (nx + use - 1) / use;
  int64_t gx    = (warps + kChainThreads / 32 - 1) / (kChainThreads / 32);
  int64_t gy    = (ny + *rows - 1) / *rows;
  if (gy > 65535)
  {
    *rows = (int)((ny + 65534) / 65535);
    gy    = (ny + *rows - 1) / *rows;
  }
  return dim3((unsigned)gx, (unsigned)gy);
}
