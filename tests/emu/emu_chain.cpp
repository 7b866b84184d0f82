// emu_chain.cpp -- TEST INFRASTRUCTURE: runs the temporally blocked stage kernels
// (csrc/chain_march.cuh, csrc/chain_quad.cuh) on CPU threads through cuda_emu.h and exposes one
// C entry point for the pytest side (tests/test_kernel_emulation.py).  Built by tests/emu/Makefile
// with -DB200_HOST_EMU; the product never links this.
#define EMU_DEFINE_SWITCH
#include "b200_sts.h"
#include "chain_march.cuh"
#include "chain_quad.cuh"
#include "stage_kernels.cuh"

#include <vector>

namespace
{
template <int NT, uint32_t PAT>
void run_stage(const StageArgs& a, dim3 grid)
{
  if (a.region == 2) emu::launch(k_stage_march<NT, PAT, 2, 0>, grid, kThreads, 0, a);
  else if (a.rw && a.ewt_out) emu::launch(k_stage_march<NT, PAT, 0, 2>, grid, kThreads, 0, a);
  else if (a.rw) emu::launch(k_stage_march<NT, PAT, 0, 1>, grid, kThreads, 0, a);
  else emu::launch(k_stage_march<NT, PAT, 0, 0>, grid, kThreads, 0, a);
}

int g_emu_bulk = 0; // emu_set_chain_bulk: the BULK flavour of k_chain_march (exact arithmetic)

template <int K, int PF, bool HALO, bool FMA>
void run_march(const ChainArgs& a, dim3 grid, bool uni)
{
  if constexpr (!FMA)
    if (g_emu_bulk)
    {
      if (g_emu_bulk == 2)
      { // one row more in flight
        const size_t smem = chain_march_smem(K, PF + 1, a.rows, false, true);
        if (uni) emu::launch(k_chain_march<K, PF + 1, HALO, false, true, false, false, true>, grid, kChainThreads, smem, a);
        else emu::launch(k_chain_march<K, PF + 1, HALO, false, false, false, false, true>, grid, kChainThreads, smem, a);
        return;
      }
      const size_t smem = chain_march_smem(K, PF, a.rows, false, true);
      if (uni) emu::launch(k_chain_march<K, PF, HALO, false, true, false, false, true>, grid, kChainThreads, smem, a);
      else emu::launch(k_chain_march<K, PF, HALO, false, false, false, false, true>, grid, kChainThreads, smem, a);
      return;
    }
  if (uni) emu::launch(k_chain_march<K, PF, HALO, FMA, true>, grid, kChainThreads, chain_march_smem(K, PF, a.rows), a);
  else emu::launch(k_chain_march<K, PF, HALO, FMA, false>, grid, kChainThreads, chain_march_smem(K, PF, a.rows), a);
}
template <int K, int PF>
void run_march_k(const ChainArgs& a, dim3 grid, bool halo, bool fma, bool uni)
{
  if (halo) fma ? run_march<K, PF, true, true>(a, grid, uni) : run_march<K, PF, true, false>(a, grid, uni);
  else fma ? run_march<K, PF, false, true>(a, grid, uni) : run_march<K, PF, false, false>(a, grid, uni);
}

template <int K, int PF, bool HALO, bool FMA>
void run_quad(const ChainArgs& a, dim3 grid)
{
  emu::launch(k_chain_quad<K, PF, HALO, FMA>, grid, kQuadThreads, chain_quad_smem(K, PF, a.rows), a);
}
template <int K, int PF>
void run_quad_k(const ChainArgs& a, dim3 grid, bool halo, bool fma)
{
  if (halo) fma ? run_quad<K, PF, true, true>(a, grid) : run_quad<K, PF, true, false>(a, grid);
  else fma ? run_quad<K, PF, false, true>(a, grid) : run_quad<K, PF, false, false>(a, grid);
}
} // namespace

extern "C" __attribute__((visibility("default"))) void emu_set_chain_bulk(int on) { g_emu_bulk = on; }

// variant: 0 = k_chain_march (2 cells / thread), 1 = k_chain_quad (4 cells / thread)
// halos: NULL (periodic wrap) or the four deep-halo buffers {x, prev2, yn, fn} (g rows, g2 columns)
// lazy_cp_async: see cuda_emu.h.  uniform4: NULL, or {cxw, cxe, cys, cyn} = the uniform-coefficient
// flavour of k_chain_march (the tables are then not read).  Returns 0, or -1 if unsupported.
extern "C" __attribute__((visibility("default"))) int emu_stencil_chain(int variant, int K, int fma, int lazy_cp_async, int64_t nx, int64_t ny,
                                 const double* cxw, const double* cxe, const double* cys, const double* cyn,
                                 const double* x, const double* prev2, const double* yn, const double* fn,
                                 const double* coeffs, double* const* out, int rows,
                                 const double* const* halos, int g, int g2, const double* uniform4)
{
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = nx; a.ny = ny;
  a.cxw = cxw; a.cxe = cxe; a.cys = cys; a.cyn = cyn;
  a.x = x; a.prev2 = prev2; a.yn = yn; a.fn = fn;
  for (int l = 0; l < K; l++)
  {
    for (int k = 0; k < 5; k++) a.c[l][k] = coeffs[5 * l + k];
    a.out[l] = out[l];
  }
  a.rows = rows;
  if (halos) { a.hx = halos[0]; a.hp = halos[1]; a.hy = halos[2]; a.hf = halos[3]; a.g = g; a.g2 = g2; }
  const bool uni = uniform4 != nullptr;
  if (uni && memcmp(&uniform4[0], &uniform4[1], sizeof(double)) != 0) return -1; // the uniform flavour needs Dx_w == Dx_e
  if (uni)
  {
    a.u_cxw = uniform4[0]; a.u_cxe = uniform4[1]; a.u_cys = uniform4[2]; a.u_cyn = uniform4[3];
    a.u_ndc = -((uniform4[0] + uniform4[1]) + (uniform4[2] + uniform4[3]));
  }
  emu::cp_async_lazy = lazy_cp_async != 0;
  const bool h = halos != nullptr, f = fma != 0;
  if (variant == 0)
  {
    dim3 grid = chain_march_grid(nx, ny, K, &a.rows);
    switch (K)
    {
    case 2: run_march_k<2, 4>(a, grid, h, f, uni); break;
    case 3: run_march_k<3, 4>(a, grid, h, f, uni); break;
    case 4: run_march_k<4, 3>(a, grid, h, f, uni); break;
    case 5: run_march_k<5, 3>(a, grid, h, f, uni); break;
    case 6: run_march_k<6, 3>(a, grid, h, f, uni); break;
    default: return -1;
    }
    return 0;
  }
  if (variant == 1)
  {
    if (uni) return -1;
    if (!chain_quad_supported(nx, ny, K, h ? g2 : -1)) return -1;
    dim3 grid = chain_quad_grid(nx, ny, K, &a.rows);
    switch (K)
    {
    case 2: run_quad_k<2, kQuadPF>(a, grid, h, f); break;
    case 3: run_quad_k<3, kQuadPF>(a, grid, h, f); break;
    case 4: run_quad_k<4, kQuadPF>(a, grid, h, f); break;
    case 5: run_quad_k<5, kQuadPF>(a, grid, h, f); break;
    case 6: run_quad_k<6, kQuadPF>(a, grid, h, f); break;
    default: return -1;
    }
    return 0;
  }
  return -1;
}


// b200_stencil_lincomb on the emulator: the same kernel choice as the launcher in b200_kernels.cu
// (region 1 -> k_stage_ring; even nx -> k_stage_march with the compiled pattern for the LSRKStep
// sequences, else the runtime pattern; odd nx or force_generic -> k_stage_generic).
// wrms_w / wrms_result: fused sum((z*w)^2) or NULL.  Returns 0, or -1 for a bad argument.
// b200_stage_extras.ewt_* for the NEXT emu_stencil_lincomb call (which must carry wrms_w)
static double* g_ewt_out    = nullptr;
static double* g_ewt_result = nullptr;
static double g_ewt_rtol = 0.0, g_ewt_atol = 0.0;
extern "C" __attribute__((visibility("default"))) void emu_set_next_ewt(double* ewt_out, double rtol, double atol, double* result)
{
  g_ewt_out = ewt_out; g_ewt_rtol = rtol; g_ewt_atol = atol; g_ewt_result = result;
}

extern "C" __attribute__((visibility("default"))) int emu_stencil_lincomb(
  int64_t nx, int64_t ny, const double* cxw, const double* cxe, const double* cys, const double* cyn,
  const double* hw, const double* he, const double* hs, const double* hn, const double* x, int nterms,
  const double* cf, const int* src, const double* const* v, double* z, double* f_out, double* send_w,
  double* send_e, double* send_s, double* send_n, const double* wrms_w, double* wrms_result, int rows,
  int region, int force_generic)
{
  if (nterms < 1 || nterms > B200_MAX_TERMS) return -1;
  StageArgs a;
  memset(&a, 0, sizeof(a));
  a.nx = nx; a.ny = ny;
  a.cxw = cxw; a.cxe = cxe; a.cys = cys; a.cyn = cyn;
  a.hw = hw; a.he = he; a.hs = hs; a.hn = hn;
  a.x = x; a.z = z;
  a.t.n = nterms;
  for (int k = 0; k < nterms; k++)
  {
    a.t.c[k]   = cf[k];
    a.t.src[k] = src[k];
    a.t.v[k]   = (src[k] == B200_SRC_VECTOR) ? v[k] : nullptr;
  }
  a.f_out = f_out;
  a.send_w = send_w; a.send_e = send_e; a.send_s = send_s; a.send_n = send_n;
  a.rw = wrms_w; a.result = wrms_result;
  if (g_ewt_out && wrms_w)
  {
    a.ewt_out = g_ewt_out; a.ewt_rtol = g_ewt_rtol; a.ewt_atol = g_ewt_atol; a.result2 = g_ewt_result;
  }
  g_ewt_out = nullptr;
  a.region = region;
  std::vector<double> partials(1 << 16);
  unsigned ticket = 0;
  a.partials = partials.data();
  a.ticket   = &ticket;
  if (region == 1)
  {
    const int64_t cells = 2 * nx + 2 * (ny - 2);
    emu::launch(k_stage_ring, dim3((unsigned)((cells + kThreads - 1) / kThreads)), kThreads, 0, a);
    return 0;
  }
  if ((nx % 2 == 0) && !force_generic)
  {
    a.rows = rows;
    dim3 grid((unsigned)((nx / 2 + kThreads - 1) / kThreads), (unsigned)((ny + rows - 1) / rows));
    uint32_t pat = 0;
    for (int k = 0; k < nterms; k++) pat |= (uint32_t)src[k] << (2 * k);
    const int V = B200_SRC_VECTOR, C = B200_SRC_CENTRE, S = B200_SRC_STENCIL;
    if (nterms == 1 && pat == PAT1(S)) run_stage<1, PAT1(S)>(a, grid);
    else if (nterms == 2 && pat == PAT2(C, S)) run_stage<2, PAT2(C, S)>(a, grid);
    else if (nterms == 4 && pat == PAT4(V, C, V, S)) run_stage<4, PAT4(V, C, V, S)>(a, grid);
    else if (nterms == 5 && pat == PAT5(S, V, V, C, V)) run_stage<5, PAT5(S, V, V, C, V)>(a, grid);
    else run_stage<B200_MAX_TERMS, PAT_RUNTIME>(a, grid);
    return 0;
  }
  emu::launch(k_stage_generic, dim3((unsigned)((nx + kThreads - 1) / kThreads), (unsigned)ny), kThreads, 0, a);
  return 0;
}
