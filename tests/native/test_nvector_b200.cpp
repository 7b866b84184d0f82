// test_nvector_b200.cpp -- SUNDIALS' own N_Vector unit-test suite
// (deps/sundials/test/unit_tests/nvector/test_nvector.c: closed-form answers for every vector
// operation) run against N_Vector_B200.  Only this driver is ours: the Test_N_V*
// functions are compiled from the reference tree where they lie (oracle/Makefile, target nvsuite);
// it mirrors serial/test_nvector_serial.c:28-240 for the operations the B200 ops table provides
// (nvector_b200.cpp, N_VNew_B200) and implements the backend hooks the suite asks for.
//
//   oracle/_ref/test_nvector_b200 <length> <print timing 0|1>      (needs a CUDA device)
#include <cstdio>
#include <cstdlib>

#include "b200_sts.h"
#include "nvector_b200.h"

extern "C" {
#include <sundials/sundials_math.h>
#include "test_nvector.h"
}

int main(int argc, char* argv[])
{
  int fails = 0;
  if (argc < 3)
  {
    printf("ERROR: TWO (2) inputs required: vector length, print timing\n");
    return -1;
  }
  const sunindextype length = (sunindextype)atol(argv[1]);
  if (length <= 0)
  {
    printf("ERROR: length of vector must be a positive integer\n");
    return -1;
  }
  Test_Init(SUN_COMM_NULL);
  SetTiming(atoi(argv[2]), 0);
  printf("Testing the B200 N_Vector\nVector length %ld\n", (long)length);

  b200_ctx* dev = nullptr;
  if (b200_ctx_create(0, nullptr, &dev))
  {
    printf("FAIL: %s\n", b200_last_error());
    Test_Finalize();
    return 1;
  }
  N_Vector X = N_VNew_B200(dev, length, length, sunctx);
  if (!X)
  {
    printf("FAIL: Unable to create a new vector\n");
    Test_Finalize();
    return 1;
  }

  fails += Test_N_VGetLength(X, 0);
  fails += Test_N_VGetCommunicator(X, SUN_COMM_NULL, 0);
  fails += Test_N_VClone(X, length, 0);
  fails += Test_N_VCloneVectorArray(5, X, length, 0);
  fails += Test_N_VGetArrayPointer(X, length, 0);

  N_Vector Y = N_VClone(X), Z = N_VClone(X);
  if (!Y || !Z)
  {
    printf("FAIL: Unable to clone\n");
    Test_Finalize();
    return 1;
  }

  /* standard vector operations (test_nvector_serial.c:122-146) */
  fails += Test_N_VConst(X, length, 0);
  fails += Test_N_VLinearSum(X, Y, Z, length, 0);
  fails += Test_N_VProd(X, Y, Z, length, 0);
  fails += Test_N_VDiv(X, Y, Z, length, 0);
  fails += Test_N_VScale(X, Z, length, 0);
  fails += Test_N_VAbs(X, Z, length, 0);
  fails += Test_N_VInv(X, Z, length, 0);
  fails += Test_N_VAddConst(X, Z, length, 0);
  fails += Test_N_VDotProd(X, Y, length, 0);
  fails += Test_N_VMaxNorm(X, length, 0);
  fails += Test_N_VWrmsNorm(X, Y, length, 0);
  fails += Test_N_VMin(X, length, 0);
  fails += Test_N_VWL2Norm(X, Y, length, 0);
  fails += Test_N_VL1Norm(X, length, 0);
  /* the fused operation LSRKStep relies on (:152-160: with the fused op enabled) */
  fails += Test_N_VLinearCombination(X, length, 0);

  N_VDestroy(X);
  N_VDestroy(Y);
  N_VDestroy(Z);
  b200_ctx_destroy(dev);

  if (fails) printf("FAIL: NVector module failed %i tests \n\n", fails);
  else printf("SUCCESS: NVector module passed all tests \n\n");
  Test_Finalize();
  return fails;
}

/* ---- backend hooks the suite declares (test_nvector.h:40-47).  N_Vector_B200 hands out a pinned host
   mirror: N_VGetArrayPointer brings it up to date and marks it dirty, the next device operation uploads it
   again, so every hook is a view on that mirror. */
namespace
{
sunrealtype* host_view(N_Vector v) { return N_VGetArrayPointer(v); }
} // namespace

extern "C" {
int check_ans(sunrealtype expected, N_Vector v, sunindextype n)
{
  const sunrealtype* h = host_view(v);
  sunindextype bad     = 0;
  for (sunindextype k = 0; k < n; ++k)
    if (SUNRCompare(h[k], expected)) ++bad;
  return bad ? 1 : 0;
}

sunbooleantype has_data(N_Vector v) { return host_view(v) ? SUNTRUE : SUNFALSE; }

void set_element_range(N_Vector v, sunindextype first, sunindextype last, sunrealtype value)
{
  sunrealtype* h = host_view(v);
  for (sunindextype k = first; k <= last; ++k) h[k] = value;
}

void set_element(N_Vector v, sunindextype k, sunrealtype value) { set_element_range(v, k, k, value); }

sunrealtype get_element(N_Vector v, sunindextype k) { return host_view(v)[k]; }

double max_time(N_Vector, double seconds) { return seconds; } /* one rank */

void sync_device(N_Vector v) { b200_ctx_sync(N_VGetContext_B200(v)); }
}
