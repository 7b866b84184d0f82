// emu_kernels.cpp -- TEST INFRASTRUCTURE: the elementwise / reduction / halo / Jacobi / adr kernels of
// ceda-demonstrations_b200/csrc/*.cuh on the host emulator (cuda_emu.h), one C entry point per
// b200_* launcher they stand behind.  Grids are deliberately small and odd-sized so that the
// grid-stride loops, the partial-block tails and the last-ticket reduction are exercised.
#include "b200_sts.h"
#include "vector_kernels.cuh"
#include "halo_kernels.cuh"
#include "adr_kernels.cuh"

#include <vector>

#define EMU_API extern "C" __attribute__((visibility("default")))

namespace
{
template <class F>
void launch_fn(F f, dim3 grid)
{
  emu::launch([&](int) { f(); }, grid, kThreads, 0, 0);
}

template <int OP>
void run_ew(const EwArgs& a, unsigned blocks)
{
  emu::launch(k_elementwise<OP>, dim3(blocks), kThreads, 0, a);
}

template <int KIND, int ROP>
double run_reduce(const double* x, const double* y, int64_t n, unsigned blocks, double ys = 0.0)
{
  std::vector<double> partials(blocks + 8);
  unsigned ticket = 0;
  double result   = 0.0;
  launch_fn([&]() { k_reduce<KIND, ROP>(x, y, ys, n, partials.data(), &ticket, &result); }, dim3(blocks));
  return result;
}
} // namespace

// op: EwOp of vector_kernels.cuh (0 lincomb, 1 a(x+y), 2 a(x-y), 3 const, 4 prod, 5 div, 6 abs, 7 inv,
// 8 addconst, 9 ewt = 1/(a|x|+b))
EMU_API int emu_elementwise(int op, int64_t n, const double* x, const double* y, double a, double b, int nterms,
                            const double* cf, const double* const* v, double* z, int blocks)
{
  EwArgs e;
  memset(&e, 0, sizeof(e));
  e.x = x; e.y = y; e.a = a; e.b = b; e.z = z; e.n = n;
  e.t.n = nterms;
  for (int k = 0; k < nterms; k++) { e.t.c[k] = cf[k]; e.t.v[k] = v[k]; }
  switch (op)
  {
  case EW_LINCOMB: run_ew<EW_LINCOMB>(e, blocks); break;
  case EW_SCALESUM: run_ew<EW_SCALESUM>(e, blocks); break;
  case EW_SCALEDIFF: run_ew<EW_SCALEDIFF>(e, blocks); break;
  case EW_CONST: run_ew<EW_CONST>(e, blocks); break;
  case EW_PROD: run_ew<EW_PROD>(e, blocks); break;
  case EW_DIV: run_ew<EW_DIV>(e, blocks); break;
  case EW_ABS: run_ew<EW_ABS>(e, blocks); break;
  case EW_INV: run_ew<EW_INV>(e, blocks); break;
  case EW_ADDCONST: run_ew<EW_ADDCONST>(e, blocks); break;
  case EW_EWT: run_ew<EW_EWT>(e, blocks); break;
  default: return -1;
  }
  return 0;
}

// kind: RdKind (0 dot, 1 sum((x*w)^2), 2 max|x|, 3 min, 4 sum|x|, 5 sum((x*ys)^2) with the scalar weight y[0])
EMU_API double emu_reduce(int kind, int64_t n, const double* x, const double* y, int blocks)
{
  switch (kind)
  {
  case RD_WSQRC: return run_reduce<RD_WSQRC, RED_SUM>(x, nullptr, n, blocks, y[0]);
  case RD_DOT: return run_reduce<RD_DOT, RED_SUM>(x, y, n, blocks);
  case RD_WSQR: return run_reduce<RD_WSQR, RED_SUM>(x, y, n, blocks);
  case RD_MAXNORM: return run_reduce<RD_MAXNORM, RED_MAX>(x, nullptr, n, blocks);
  case RD_MIN: return run_reduce<RD_MIN, RED_MIN>(x, nullptr, n, blocks);
  default: return run_reduce<RD_L1, RED_SUM>(x, nullptr, n, blocks);
  }
}

EMU_API void emu_pack(const double* u, int64_t nx, int64_t ny, double* sw, double* se, double* ss, double* sn)
{
  const int64_t m = nx > ny ? nx : ny;
  launch_fn([&]() { k_pack(u, nx, ny, sw, se, ss, sn); }, dim3((unsigned)((m + kThreads - 1) / kThreads)));
}

EMU_API void emu_pack_strips(const double* field, const double* halo, int64_t nx, int64_t ny, int g, int g2,
                             double* wstrip, double* estrip)
{
  const int64_t cells = (ny + 2 * g) * g2;
  launch_fn([&]() { k_pack_strips(field, halo, nx, ny, g, g2, wstrip, estrip); },
            dim3((unsigned)((cells + kThreads - 1) / kThreads)));
}

EMU_API void emu_jacobi(int64_t nx, int64_t ny, const double* pxw, const double* pxe, const double* pys,
                        const double* pyn, double gamma, double* diag)
{
  launch_fn([&]() { k_jacobi(nx, ny, pxw, pxe, pys, pyn, gamma, diag); },
            dim3((unsigned)((nx + kThreads - 1) / kThreads), (unsigned)ny));
}

// b200_adr_rhs (nterms = 0: f = F_mode(y)) / b200_adr_lincomb (z = sum c_k T_k, optionally f_out = F(y))
EMU_API int emu_adr(const b200_adr_params* p, int mode, const double* y, int nterms, const double* cf, const int* src,
                    const double* const* v, double* z, double* f_out, int rows)
{
  if (mode < 1 || mode > 7) return -1;
  AdrArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = p->nx; a.ny = p->ny; a.k = adr_consts(*p); a.y = y; a.f = f_out; a.z = z;
  a.t.n = nterms;
  for (int k = 0; k < nterms; k++)
  {
    a.t.c[k] = cf[k]; a.t.src[k] = src[k];
    a.t.v[k] = (src[k] == B200_SRC_VECTOR) ? v[k] : nullptr;
  }
  a.rows = rows;
  dim3 grid((unsigned)((a.nx + kThreads - 1) / kThreads), (unsigned)((a.ny + rows - 1) / rows));
  switch (mode)
  {
  case 1: emu::launch(k_adr_march<1>, grid, kThreads, 0, a); break;
  case 2: emu::launch(k_adr_march<2>, grid, kThreads, 0, a); break;
  case 3: emu::launch(k_adr_march<3>, grid, kThreads, 0, a); break;
  case 4: emu::launch(k_adr_march<4>, grid, kThreads, 0, a); break;
  case 5: emu::launch(k_adr_march<5>, grid, kThreads, 0, a); break;
  case 6: emu::launch(k_adr_march<6>, grid, kThreads, 0, a); break;
  default: emu::launch(k_adr_march<7>, grid, kThreads, 0, a); break;
  }
  return 0;
}

// b200_adr_chain on the emulator: K temporally blocked stages of the adr diffusion partition
#include "adr_chain.cuh"

namespace
{
template <int K>
void run_adr_chain(const AdrChainArgs& a)
{
  emu::launch(k_adr_chain<K, kAdrChainPF>, adr_chain_grid(a.nx, a.ny, K, a.rows), kAdrChainThreads,
              adr_chain_smem(K, kAdrChainPF), a);
}
} // namespace

EMU_API int emu_adr_chain(const b200_adr_params* p, int K, const double* x, const double* prev2, const double* yn,
                          const double* fn, const double* coeffs, double* const* out, int rows, int lazy_cp_async)
{
  if (!adr_chain_supported(p->nx, p->ny, K)) return -1;
  AdrChainArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = p->nx; a.ny = p->ny; a.k = adr_consts(*p);
  a.x = x; a.prev2 = prev2; a.yn = yn; a.fn = fn;
  for (int l = 0; l < K; l++)
  {
    for (int q = 0; q < 5; q++) a.c[l][q] = coeffs[5 * l + q];
    a.out[l] = out[l];
  }
  a.rows             = rows;
  emu::cp_async_lazy = lazy_cp_async != 0;
  switch (K)
  {
  case 2: run_adr_chain<2>(a); break;
  case 3: run_adr_chain<3>(a); break;
  case 4: run_adr_chain<4>(a); break;
  case 5: run_adr_chain<5>(a); break;
  case 6: run_adr_chain<6>(a); break;
  default: return -1;
  }
  return 0;
}

// A periodic npx x npy process grid simulated in one process: every "rank" pushes the edge bands and corners of its
// block into its neighbours' deep-halo slots with k_peer_exchange, addressed by peer_dst_pointers exactly as
// b200_peer_halo_exchange does.  Ranks run one after the other here, so the arrival flags are raised beforehand.
// blocks[r] / slots[r]: rank r's nx_loc x ny_loc[r] block and its [S | N | W | E] deep halo; rank = idx * npy + idy
// (diffusion_2D.cpp:243-317), all blocks nx_loc wide, heights ny_loc[idy].
EMU_API int emu_peer_exchange_grid(int npx, int npy, int64_t nx_loc, const int64_t* ny_loc, int g, int g2,
                                   const double* const* blocks, double* const* slots)
{
  const int np = npx * npy;
  std::vector<unsigned long long> flags((size_t)np * 8, 1ull), sink(8);
  std::vector<unsigned> ticket((size_t)np, 0u);
  int err = 0;
  auto rank_of = [&](int cx, int cy) { return ((cx % npx + npx) % npx) * npy + ((cy % npy + npy) % npy); };
  for (int r = 0; r < np; r++)
  {
    const int cx = r / npy, cy = r % npy;
    const int nbr[8] = {rank_of(cx - 1, cy),     rank_of(cx + 1, cy),     rank_of(cx, cy - 1),     rank_of(cx, cy + 1),
                        rank_of(cx - 1, cy - 1), rank_of(cx + 1, cy - 1), rank_of(cx - 1, cy + 1), rank_of(cx + 1, cy + 1)};
    PeerXArgs a;
    memset(&a, 0, sizeof(a));
    a.nf = 1; a.nx = nx_loc; a.ny = ny_loc[cy]; a.g = g; a.g2 = g2; a.epoch = 1;
    a.field[0] = blocks[r];
    double* nbr_slot[8];
    for (int d = 0; d < 8; d++) nbr_slot[d] = slots[nbr[d]];
    peer_dst_pointers(nbr_slot, nx_loc, ny_loc[cy], ny_loc[((cy - 1) % npy + npy) % npy], ny_loc[(cy + 1) % npy], g, g2, a.dst[0]);
    for (int d = 0; d < 8; d++) a.peer_flag[d] = &sink[(size_t)d];
    a.my_flag    = &flags[(size_t)r * 8];
    a.ticket     = &ticket[(size_t)r];
    a.err        = &err;
    a.timeout_ns = 1000000000ull;
    emu::launch(k_peer_exchange, dim3(3, 1), kThreads, 0, a);
  }
  return err;
}
