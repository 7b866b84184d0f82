// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: a minimal host emulation of the CUDA execution model,
// just large enough to run the stage kernels of ceda-demonstrations_b200/csrc/*.cuh on CPU threads.
//
// One fiber per CUDA thread, a block on one OS thread, blocks one after the other: deterministic and
// fast (a context switch costs a fraction of a microsecond).  What is emulated:
//   threadIdx / blockIdx / blockDim / gridDim, dynamic and static shared memory, __syncthreads (block
//   barrier), atomicAdd(unsigned) / __threadfence (the last-block ticket of the reductions),
//   __shfl_{up,down,xor}_sync on doubles (per-warp exchange + barrier), exactly rounded FP64
//   (__dmul_rn ... compile with -ffp-contract=off; DFMA = std::fma), cp.async groups, and shared-memory mbarriers
//   with transaction counts + bulk copies (cp.async.bulk), __syncwarp, elect.
// cp.async has two modes (emu::cp_async_lazy): eager = the copy happens at issue time; lazy = the
// copy happens at the last legal moment (the cp.async.wait_group that forces its group), so a
// kernel that reads a ring slot before waiting for it, or overwrites one before it was consumed,
// produces wrong numbers in at least one of the two modes.
//
// Nothing in the product includes this file: kernel_prims.cuh pulls it in only under
// B200_HOST_EMU, which only tests/emu/Makefile defines.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <functional>
#include <memory>
#include <unordered_map>
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
// One CUDA thread = one fiber; a block runs on ONE OS thread, blocks one after the other.  Fibers
// are resumed warp by warp, lane by lane, and give the processor back only at barriers (a shuffle is two warp
// barriers), so execution is deterministic and a context switch costs a fraction of a microsecond.
struct Block;
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
// Context switch: callee-saved registers + stack pointer, no signal mask (glibc's swapcontext makes two
// system calls per switch, which dominated the run time).  x86-64 System V only; defined once, in the
// translation unit that sets EMU_DEFINE_SWITCH before including this header.
#if !defined(__x86_64__)
#error "tests/emu: the fiber switch is written for x86-64"
#endif
extern "C" void emu_switch(void** save_sp, void* load_sp);
#ifdef EMU_DEFINE_SWITCH
asm(R"(
.text
.globl emu_switch
.hidden emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
#endif

struct Fiber
{
  void* sp = nullptr; // saved stack pointer while the fiber is not running
  char* stack = nullptr; // from stack_pool(): uninitialised, kept across launches (pages are touched on demand)
  ThreadState st;
  unsigned tid = 0;
  bool done    = false;
};
struct Warp
{
  double xch[32];
  int arrived = 0, alive = 32;
  unsigned gen = 0;
};
struct BulkCopy
{
  void* dst;
  const void* src;
  unsigned bytes;
};
// shared-memory mbarrier with a transaction count (mbar_* / bulk_g2s below); the state lives beside the block, keyed by
// the barrier's address
struct MBar
{
  unsigned count = 0;  // arrivals per phase
  int pending    = 0;  // arrivals still missing in the current phase
  long tx        = 0;  // bytes announced and not yet delivered
  unsigned phase = 0;  // completed phases
  std::vector<BulkCopy> copies; // lazy mode: bulk copies issued and not yet performed
};
struct Block
{
  std::unordered_map<const void*, MBar> bars;
  std::vector<Warp> warps;
  std::vector<Fiber> fibers;
  int arrived = 0, alive = 0;
  unsigned gen = 0;
  unsigned long progress = 0; // barrier arrivals + finished threads: a scheduler pass without any is a deadlock
  char* smem   = nullptr;
  void* sched_sp = nullptr; // the scheduler's saved stack pointer
  Fiber* cur     = nullptr;
  std::function<void()> body; // kernel(args)
};
constexpr size_t kStackBytes = 256 * 1024;
inline std::vector<std::unique_ptr<char[]>>& stack_pool()
{
  static thread_local std::vector<std::unique_ptr<char[]>> pool;
  return pool;
}
inline thread_local Block* cur_block = nullptr;
inline ThreadState& cur() { return cur_block->cur->st; }
inline bool cp_async_lazy = false;

inline void yield_to_scheduler()
{
  Block* b = cur_block;
  emu_switch(&b->cur->sp, b->sched_sp);
}
inline void warp_barrier()
{
  Warp& w          = cur_block->warps[cur().warp];
  const unsigned g = w.gen;
  cur_block->progress++;
  if (++w.arrived == w.alive) { w.arrived = 0; w.gen++; }
  else
    while (w.gen == g) yield_to_scheduler();
}
inline void block_barrier()
{
  Block* b         = cur_block;
  const unsigned g = b->gen;
  b->progress++;
  if (++b->arrived == b->alive) { b->arrived = 0; b->gen++; }
  else
    while (b->gen == g) yield_to_scheduler();
}
inline void fiber_main()
{
  Block* b = cur_block;
  Fiber* f = b->cur;
  b->body();
  // an exited thread no longer takes part in barriers (CUDA semantics): release whoever waits for it
  f->done = true;
  b->progress++;
  Warp& w = b->warps[f->st.warp];
  if (--w.alive > 0 && w.arrived == w.alive) { w.arrived = 0; w.gen++; }
  if (--b->alive > 0 && b->arrived == b->alive) { b->arrived = 0; b->gen++; }
  emu_switch(&f->sp, b->sched_sp); // never resumed
  abort();
}

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
#define B200_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::cur().blk->smem)

static inline void __syncthreads() { emu::block_barrier(); }

static inline double emu_shfl(double v, int src_lane)
{
  emu::ThreadState& t = emu::cur();
  emu::Warp& w        = t.blk->warps[t.warp];
  w.xch[t.lane]       = v;
  emu::warp_barrier();
  const double r = (src_lane >= 0 && src_lane < 32) ? w.xch[src_lane] : v;
  emu::warp_barrier();
  return r;
}
static inline double __shfl_up_sync(unsigned, double v, int d) { return emu_shfl(v, emu::cur().lane - d); }
static inline double __shfl_down_sync(unsigned, double v, int d) { return emu_shfl(v, emu::cur().lane + d); }
static inline double __shfl_xor_sync(unsigned, double v, int m) { return emu_shfl(v, emu::cur().lane ^ m); }
static inline double __shfl_sync(unsigned, double v, int src) { return emu_shfl(v, src & 31); }

static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }

static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v)
{ // (blocks of one launch run on one OS thread at a time in the emulator; the CAS loop is for form)
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double __longlong_as_double(long long v)
{
  double d;
  memcpy(&d, &v, sizeof(d));
  return d;
}
using std::fmax;
using std::fmin;

// system-scope flag traffic of the peer-mapped halo exchange: one process, one "device" here
static inline void st_release_sys_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned long long ld_acquire_sys_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void fence_sys() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned long long global_timer_ns()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
static inline void nap_ns(unsigned)
{ // a spinning thread lets the other fibers of its block run (they may be the ones that raise the flag)
  emu::cur_block->progress++;
  emu::yield_to_scheduler();
}

static inline double2 ld_stream2(const double* p) { return *reinterpret_cast<const double2*>(p); }
static inline double2 ld_keep2(const double* p) { return *reinterpret_cast<const double2*>(p); }

static inline void cp_async16(void* smem, const void* gmem)
{
  if (((uintptr_t)smem & 15) || ((uintptr_t)gmem & 15)) abort(); // cp.async 16 needs 16-byte alignment
  if (emu::cp_async_lazy) emu::cur().open.push_back({smem, gmem});
  else memcpy(smem, gmem, 16);
}
static inline void cp_async_commit()
{
  emu::cur().groups.push_back(std::move(emu::cur().open));
  emu::cur().open.clear();
}
template <int N>
static inline void cp_async_wait()
{
  while ((int)emu::cur().groups.size() > N)
  {
    emu::run_copies(emu::cur().groups.front());
    emu::cur().groups.pop_front();
  }
}

// mbarrier + bulk copies (kernel_prims.cuh).  Eager mode: a bulk copy happens when it is issued; lazy mode
// (emu::cp_async_lazy): at the last legal moment, when somebody waits for the barrier it completes on -- so a slot read
// before its wait, or refilled before it was consumed, gives wrong numbers in at least one of the two modes.
namespace emu
{
inline void mbar_settle(MBar& b)
{
  if (b.pending == 0 && b.tx == 0)
  {
    b.phase++;
    b.pending = (int)b.count;
  }
}
inline MBar& mbar_of(const void* bar)
{
  auto it = cur_block->bars.find(bar);
  if (it == cur_block->bars.end())
  {
    fprintf(stderr, "cuda_emu: mbarrier %p used before mbar_init\n", bar);
    abort();
  }
  return it->second;
}
} // namespace emu
static inline void mbar_init(unsigned long long* bar, unsigned arrivals)
{
  emu::MBar& b = emu::cur_block->bars[bar];
  b            = emu::MBar();
  b.count      = arrivals;
  b.pending    = (int)arrivals;
}
static inline void mbar_fence_init() {}
static inline void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
  emu::MBar& b = emu::mbar_of(bar);
  if (b.pending <= 0) { fprintf(stderr, "cuda_emu: more arrivals than the mbarrier was initialised for\n"); abort(); }
  b.tx += bytes;
  b.pending--;
  emu::mbar_settle(b);
}
static inline void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar)
{
  if (((uintptr_t)smem & 15) || ((uintptr_t)gmem & 15) || (bytes & 15) || bytes == 0) abort(); // cp.async.bulk: 16-byte granules
  emu::MBar& b = emu::mbar_of(bar);
  if (emu::cp_async_lazy) { b.copies.push_back({smem, gmem, bytes}); return; }
  memcpy(smem, gmem, bytes);
  b.tx -= bytes;
  emu::mbar_settle(b);
}
static inline bool elect_one() { return emu::cur().lane == 0; }
static inline int warp_uniform(int v) { return v; }
static inline void mbar_wait(unsigned long long* bar, unsigned parity)
{
  emu::MBar& b = emu::mbar_of(bar);
  for (long spins = 0; (b.phase & 1u) == (parity & 1u); spins++)
  {
    if (!b.copies.empty())
    {
      for (const emu::BulkCopy& c : b.copies)
      {
        memcpy(c.dst, c.src, c.bytes);
        b.tx -= c.bytes;
      }
      b.copies.clear();
      emu::mbar_settle(b);
      continue;
    }
    if (spins > 100000) { fprintf(stderr, "cuda_emu: mbar_wait never completes (missing arrival or bytes)\n"); abort(); }
    emu::cur_block->progress++; // the other lanes of the warp may be the ones that still have to arrive
    emu::yield_to_scheduler();
  }
}

namespace emu
{
// launch<<<grid, block, smem>>>: blocks one after the other, each on the calling OS thread; `body` is what every
// CUDA thread executes (the kernel bound to its arguments)
inline void launch_body(dim3 grid, unsigned nthreads, size_t smem_bytes, const std::function<void()>& body)
{
  if (nthreads % 32) abort();
  auto& pool = stack_pool();
  while (pool.size() < nthreads) pool.emplace_back(new char[kStackBytes]);
  std::vector<char> smem_store(smem_bytes + 64);
  char* smem = smem_store.data();
  smem += (64 - ((uintptr_t)smem & 63)) & 63;
  for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++)
    {
      Block blk;
      blk.smem = smem;
      memset(smem, 0xff, smem_bytes); // NaN-poison: a slot read before it was filled shows up
      blk.warps.resize(nthreads / 32);
      blk.fibers.resize(nthreads);
      blk.alive = (int)nthreads;
      blk.body  = body;
      cur_block = &blk;
      for (unsigned t = 0; t < nthreads; t++)
      {
        Fiber& f  = blk.fibers[t];
        f.tid     = t;
        f.st.blk  = &blk;
        f.st.lane = (int)(t & 31);
        f.st.warp = (int)(t >> 5);
        f.stack = pool[t].get();
        // initial frame: six callee-saved registers, then the entry point as the return address; after the
        // `ret` the stack pointer must be 8 modulo 16, as it is right after a call instruction
        uintptr_t top = ((uintptr_t)(f.stack + kStackBytes) & ~(uintptr_t)15) - 8;
        void** frame  = reinterpret_cast<void**>(top) - 7;
        for (int q = 0; q < 6; q++) frame[q] = nullptr;
        frame[6] = reinterpret_cast<void*>(&fiber_main);
        f.sp     = frame;
      }
      blockIdx = {bx, by, 0};
      blockDim = dim3(nthreads);
      gridDim  = grid;
      int remaining = (int)nthreads;
      while (remaining > 0)
      {
        const unsigned long before = blk.progress;
        for (Fiber& f : blk.fibers)
        {
          if (f.done) continue;
          blk.cur   = &f;
          threadIdx = {f.tid, 0, 0};
          emu_switch(&blk.sched_sp, f.sp);
          if (f.done) remaining--;
        }
        if (remaining > 0 && blk.progress == before)
        {
          fprintf(stderr, "cuda_emu: deadlock in block (%u, %u): %d threads wait at a barrier the others never reach\n", bx, by, remaining);
          abort();
        }
      }
      cur_block = nullptr;
    }
}
template <class Kernel, class Args>
void launch(Kernel kernel, dim3 grid, unsigned nthreads, size_t smem_bytes, const Args& args)
{
  launch_body(grid, nthreads, smem_bytes, [&]() { kernel(args); });
}
} // namespace emu
