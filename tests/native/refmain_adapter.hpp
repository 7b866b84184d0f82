// refmain_adapter.hpp -- TEST INFRASTRUCTURE: the few lines a maintainer adds next to the reference's own
// diffusion_2D/main.cpp to run it on N_Vector_B200.  tests/native/patch_reference_main.py
// includes this header from the patched copy of main.cpp; nothing else of the reference driver changes: its UserData,
// UserOptions, UserOutput, Initial(), the ARKODE call sequence and the statistics output are the reference's.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "b200_callbacks.h"
#include "b200_sts.h"
#include "nvector_b200.h"

struct B200Adapter
{
  b200_ctx* ctx          = nullptr;
  b200_d2d_problem* prob = nullptr;
  ~B200Adapter() { /* the process is about to exit; device memory goes with the context */ }
};
static B200Adapter g_b200;

// one process per GPU: device = LOCAL_RANK / B200_DEVICE, else 0
static b200_ctx* b200_adapter_ctx()
{
  if (!g_b200.ctx)
  {
    const char* d = getenv("LOCAL_RANK");
    if (!d) d = getenv("B200_DEVICE");
    if (b200_ctx_create(d ? atoi(d) : 0, nullptr, &g_b200.ctx))
    {
      fprintf(stderr, "b200_ctx_create: %s\n", b200_last_error());
      exit(1);
    }
  }
  return g_b200.ctx;
}

// user_data for the b200_diffusion_* callbacks from the reference's UserData (after udata.setup(), with udata.diag
// already cloned when preconditioning is on: main.cpp:161, :227, :266)
template <class RefUserData>
static void* b200_adapter_user_data(RefUserData& u)
{
  if (!g_b200.prob &&
      b200_d2d_problem_create(b200_adapter_ctx(), u.nx, u.ny, u.xl, u.xu, u.yl, u.yu, u.kx, u.ky, u.inhomogeneous ? 1 : 0, u.npx,
                              u.npy, u.myid_c, u.np, u.diag, &g_b200.prob))
  {
    fprintf(stderr, "b200_d2d_problem_create failed: %s\n", b200_last_error());
    exit(1);
  }
  return b200_d2d_problem_user_data(g_b200.prob);
}
