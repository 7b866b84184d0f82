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


// ------------------------------------------------------------ peer-mapped deep-halo exchange
// One kernel per exchange, no NCCL: every rank WRITES the edge bands of its fields straight into the deep-halo
// slots of its (up to) eight neighbours -- peer-mapped device memory, i.e. stores that travel over NVLink /
// NVSwitch -- then raises one flag per neighbour and waits until its own eight flags show that every neighbour
// has done the same for this exchange.  What arrives where (receiver's slot layout [S | N | W | E], W / E strips
// cover rows -g .. ny+g-1 so the corners live in them):
//   my rows 0..g-1        -> S neighbour's N block        my rows ny-g..ny-1   -> N neighbour's S block
//   my cols 0..g2-1       -> W neighbour's E strip        my cols nx-g2..nx-1  -> E neighbour's W strip
//   my four g x g2 corners -> the E / W strips of the diagonal neighbours (rows below 0 / above ny-1 there)
// dst[f][d] already points at the first double this rank owns in neighbour d's slot (the host adds the block /
// strip / corner-row offsets, which depend on the NEIGHBOUR's height for the strips of the S-side diagonal ones).
enum PeerDir { PD_W = 0, PD_E, PD_S, PD_N, PD_SW, PD_SE, PD_NW, PD_NE };
struct PeerXArgs
{
  int nf;                        // fields in this exchange (<= 4)
  const double* field[4];
  double* dst[4][8];
  int64_t nx, ny;
  int g, g2;
  unsigned long long epoch;      // number of this exchange (1, 2, ...), the same on every rank
  unsigned long long* peer_flag[8]; // neighbour d's arrival counter for the direction I am in, seen from it
  unsigned long long* my_flag;   // [8] my arrival counters, written by the neighbours
  unsigned* ticket;
  int* err;                      // mapped host int: set to 1 if a neighbour did not arrive within the time limit
  unsigned long long timeout_ns;
};

// Host side: where this rank's data starts in each neighbour's slot.  nbr_slot[d] = start of the slot in
// neighbour d's ring; the receiver's layout is [S | N | W strip | E strip] and the W / E strips of a block of
// height h have (h + 2g) rows, so the strips of the diagonal neighbours are addressed with THEIR heights
// (ny_s / ny_n: the rows of blocks south / north of this one; W / E neighbours are as tall as this block).
static inline void peer_dst_pointers(double* const nbr_slot[8], int64_t nx, int64_t ny, int64_t ny_s, int64_t ny_n,
                                     int g, int g2, double* dst[8])
{
  const int64_t srow = (int64_t)g * nx;
  const int64_t strip_me = (ny + 2 * (int64_t)g) * g2, strip_s = (ny_s + 2 * (int64_t)g) * g2, strip_n = (ny_n + 2 * (int64_t)g) * g2;
  dst[PD_S]  = nbr_slot[PD_S] + srow;                                             // its N block
  dst[PD_N]  = nbr_slot[PD_N];                                                    // its S block
  dst[PD_W]  = nbr_slot[PD_W] + 2 * srow + strip_me;                              // its E strip
  dst[PD_E]  = nbr_slot[PD_E] + 2 * srow;                                         // its W strip
  dst[PD_SW] = nbr_slot[PD_SW] + 2 * srow + strip_s + (ny_s + (int64_t)g) * g2;   // its E strip, rows ny..ny+g-1
  dst[PD_SE] = nbr_slot[PD_SE] + 2 * srow + (ny_s + (int64_t)g) * g2;             // its W strip, rows ny..ny+g-1
  dst[PD_NW] = nbr_slot[PD_NW] + 2 * srow + strip_n;                              // its E strip, rows -g..-1
  dst[PD_NE] = nbr_slot[PD_NE] + 2 * srow;                                        // its W strip, rows -g..-1
}
// I am my W neighbour's E neighbour, and so on
static const int kPeerOpposite[8] = {PD_E, PD_W, PD_N, PD_S, PD_NE, PD_NW, PD_SE, PD_SW};

__global__ void __launch_bounds__(kThreads) k_peer_exchange(const PeerXArgs a)
{
  const int f      = blockIdx.y;
  const int64_t nx = a.nx, ny = a.ny;
  const int g = a.g, g2 = a.g2;
  const int64_t n_sn = (int64_t)g * nx, n_we = ny * g2, n_c = (int64_t)g * g2;
  const int64_t total = 2 * n_sn + 2 * n_we + 4 * n_c;
  const double* u     = a.field[f];
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    int64_t q = t;
    if (q < n_sn) { a.dst[f][PD_S][q] = u[q]; continue; }                       // rows 0..g-1, all columns
    q -= n_sn;
    if (q < n_sn) { a.dst[f][PD_N][q] = u[(ny - g) * nx + q]; continue; }       // rows ny-g..ny-1
    q -= n_sn;
    if (q < n_we) { const int64_t r = q / g2; const int c = (int)(q - r * g2); a.dst[f][PD_W][(r + g) * g2 + c] = u[r * nx + c]; continue; }
    q -= n_we;
    if (q < n_we) { const int64_t r = q / g2; const int c = (int)(q - r * g2); a.dst[f][PD_E][(r + g) * g2 + c] = u[r * nx + nx - g2 + c]; continue; }
    q -= n_we;
    const int corner = (int)(q / n_c); // 0 SW, 1 SE, 2 NW, 3 NE
    q -= corner * n_c;
    const int r = (int)(q / g2), c = (int)(q - (int64_t)r * g2);
    const int64_t sr = (corner < 2) ? r : ny - g + r;
    const int64_t sc = (corner & 1) ? nx - g2 + c : c;
    a.dst[f][PD_SW + corner][(int64_t)r * g2 + c] = u[sr * nx + sc];
  }
  // all blocks done -> flags.  One system-scope fence per BLOCK (thread 0, after the block barrier: the barrier orders
  // the other threads' stores before it and the fence is cumulative -- the pattern of a grid barrier), not per thread:
  // with every thread fencing, 150 000 MEMBAR.SYS per exchange cost 0.36 ms (2-GPU measurement, round 2).
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    fence_sys();
    const unsigned done = atomicAdd(a.ticket, 1u);
    is_last             = (done == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x < 8)
  {
    fence_sys();
    st_release_sys_u64(a.peer_flag[threadIdx.x], a.epoch);
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys_u64(a.my_flag + threadIdx.x) < a.epoch)
    {
      nap_ns(100);
      if (global_timer_ns() - t0 > a.timeout_ns) { *a.err = 1; break; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.ticket = 0;
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

