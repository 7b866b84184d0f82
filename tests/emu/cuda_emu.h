// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: a minimal host emulation of the CUDA execution model,
// just large enough to run the stage kernels of ceda-demonstrations_b200/csrc/*.cuh on CPU threads.
//
// One OS thread per CUDA thread of a block, blocks one after the other.  What is emulated:
//   threadIdx / blockIdx / blockDim / gridDim, dynamic and static shared memory, __syncthreads (block
//   barrier), atomicAdd(unsigned) / __threadfence (the last-block ticket of the reductions),
//   __shfl_{up,down,xor}_sync on doubles (per-warp exchange + barrier), exactly rounded FP64
//   (__dmul_rn ... compile with -ffp-contract=off; DFMA = std::fma), and cp.async groups.
// cp.async has two modes (emu::cp_async_lazy): eager = the copy happens at issue time; lazy = the
// copy happens at the last legal moment (the cp.async.wait_group that forces its group), so a
// kernel that reads a ring slot before waiting for it, or overwrites one before it was consumed,
// produces wrong numbers in at least one of the two modes.
//
// Nothing in the product includes this file: kernel_prims.cuh pulls it in only under
// B200_HOST_EMU, which only tests/emu/Makefile defines.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct alignas(16) double2
{
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

struct emu_uint3
{
  unsigned x, y, z;
};
struct dim3
{
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu
{
struct Warp
{
  std::unique_ptr<std::barrier<>> bar;
  double xch[32];
};
struct Block
{
  std::unique_ptr<std::barrier<>> bar;
  std::vector<Warp> warps;
  char* smem = nullptr;
};
struct PendingCopy
{
  void* dst;
  const void* src;
};
struct ThreadState
{
  Block* blk = nullptr;
  int lane = 0, warp = 0;
  std::deque<std::vector<PendingCopy>> groups; // committed cp.async groups, oldest first
  std::vector<PendingCopy> open;               // copies issued since the last commit
};
inline thread_local ThreadState ts;
inline bool cp_async_lazy = false;

inline void run_copies(const std::vector<PendingCopy>& g)
{
  for (const PendingCopy& c : g) memcpy(c.dst, c.src, 16);
}
} // namespace emu

inline thread_local emu_uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

#define __global__
#define __shared__ static /* blocks run one after the other, so a function-level static is block-shared */
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

#define DMUL(a, b) ((double)(a) * (double)(b))
#define DADD(a, b) ((double)(a) + (double)(b))
#define DSUB(a, b) ((double)(a) - (double)(b))
#define DFMA(a, b, c) std::fma((double)(a), (double)(b), (double)(c))
static inline double __ddiv_rn(double a, double b) { return a / b; }
#define B200_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::ts.blk->smem)

static inline void __syncthreads() { emu::ts.blk->bar->arrive_and_wait(); }

static inline double emu_shfl(double v, int src_lane)
{
  emu::Warp& w       = emu::ts.blk->warps[emu::ts.warp];
  w.xch[emu::ts.lane] = v;
  w.bar->arrive_and_wait();
  const double r = (src_lane >= 0 && src_lane < 32) ? w.xch[src_lane] : v;
  w.bar->arrive_and_wait();
  return r;
}
static inline double __shfl_up_sync(unsigned, double v, int d) { return emu_shfl(v, emu::ts.lane - d); }
static inline double __shfl_down_sync(unsigned, double v, int d) { return emu_shfl(v, emu::ts.lane + d); }
static inline double __shfl_xor_sync(unsigned, double v, int m) { return emu_shfl(v, emu::ts.lane ^ m); }
static inline double __shfl_sync(unsigned, double v, int src) { return emu_shfl(v, src & 31); }

static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double __longlong_as_double(long long v)
{
  double d;
  memcpy(&d, &v, sizeof(d));
  return d;
}
using std::fmax;
using std::fmin;

static inline double2 ld_stream2(const double* p) { return *reinterpret_cast<const double2*>(p); }
static inline double2 ld_keep2(const double* p) { return *reinterpret_cast<const double2*>(p); }

static inline void cp_async16(void* smem, const void* gmem)
{
  if (((uintptr_t)smem & 15) || ((uintptr_t)gmem & 15)) abort(); // cp.async 16 needs 16-byte alignment
  if (emu::cp_async_lazy) emu::ts.open.push_back({smem, gmem});
  else memcpy(smem, gmem, 16);
}
static inline void cp_async_commit()
{
  emu::ts.groups.push_back(std::move(emu::ts.open));
  emu::ts.open.clear();
}
template <int N>
static inline void cp_async_wait()
{
  while ((int)emu::ts.groups.size() > N)
  {
    emu::run_copies(emu::ts.groups.front());
    emu::ts.groups.pop_front();
  }
}

namespace emu
{
// launch<<<grid, block, smem>>>: blocks sequentially, one OS thread per CUDA thread
template <class Kernel, class Args>
void launch(Kernel kernel, dim3 grid, unsigned nthreads, size_t smem_bytes, const Args& args)
{
  if (nthreads % 32) abort();
  std::vector<char> smem_store(smem_bytes + 64);
  char* smem = smem_store.data();
  smem += (64 - ((uintptr_t)smem & 63)) & 63;
  for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++)
    {
      Block blk;
      blk.smem = smem;
      memset(smem, 0xff, smem_bytes); // NaN-poison: a slot read before it was filled shows up
      blk.bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)nthreads);
      blk.warps.resize(nthreads / 32);
      for (Warp& w : blk.warps) w.bar = std::make_unique<std::barrier<>>(32);
      std::vector<std::thread> threads;
      threads.reserve(nthreads);
      for (unsigned t = 0; t < nthreads; t++)
        threads.emplace_back(
          [&, t]()
          {
            threadIdx = {t, 0, 0};
            blockIdx  = {bx, by, 0};
            blockDim  = dim3(nthreads);
            gridDim   = grid;
            ts        = ThreadState();
            ts.blk    = &blk;
            ts.lane   = (int)(t & 31);
            ts.warp   = (int)(t >> 5);
            kernel(args);
            // an exited thread no longer takes part in barriers (CUDA semantics)
            blk.warps[ts.warp].bar->arrive_and_drop();
            blk.bar->arrive_and_drop();
          });
      for (std::thread& th : threads) th.join();
    }
}
} // namespace emu
