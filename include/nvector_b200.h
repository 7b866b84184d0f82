/* nvector_b200.h -- a SUNDIALS N_Vector whose data lives in B200 HBM.
 *
 * Drop-in for the one constructor call that selects the vector backend in the
 * reference drivers:
 *     N_VNew_Parallel(comm, local_length, global_length, ctx)   diffusion_2D/main.cpp:176
 *     N_VNew_Serial(neq, ctx)                                    adr/advection_diffusion_reaction_2d.cpp:83
 * Every other vector (ARKODE's ewt/yn/fn/tempv*, the Lagrange history, the power
 * iteration's V/q, PCG's r/p/z/Ap, udata.diag) is an N_VClone and inherits the ops
 * table (struct _generic_N_Vector_Ops, SUN/include/sundials/sundials_nvector.h:98-192).
 *
 * Host code only: no CUDA headers.  The ops call the C-ABI of b200_sts.h.
 *
 * Lazy fusion.  A vector's value is an immutable, reference
 * counted object that is either materialised (a device buffer) or DEFERRED
 * ("the RHS operator applied to that other value").  An ARKRhsFn written for this
 * vector does not launch anything: it calls N_VSetDeferredRhs_B200(f, op, y).  The
 * next N_VLinearCombination / N_VLinearSum / N_VScale that consumes f launches ONE
 * fused kernel (stencil + recurrence), so an STS stage makes one pass over HBM;
 * N_VScale(1, x, z) only shares the value.  Any other consumer materialises first.
 */
#ifndef NVECTOR_B200_H
#define NVECTOR_B200_H

#include <sundials/sundials_nvector.h>

#include "b200_sts.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A deferred right-hand-side operator F(y).  `fused` must enqueue, on the context's
   stream, one kernel (group) that computes
       z = sum_k c[k] * T_k   (left to right),  T_k in { v[k], y, F(y) } per src[k]
   and, if f_out != NULL, also stores F(y) there.  If wrms_w != NULL it may fuse
   sum_i (z_i w_i)^2 into wrms_result (device) and set *wrms_done = 1.
   y, v[k], z, f_out are device pointers of the vector's local length. */
typedef struct B200RhsOp
{
  void* self;
  int (*fused)(void* self, b200_ctx* ctx, const double* y, int nterms, const double* c,
               const int* src, const double* const* v, double* z, double* f_out,
               const double* wrms_w, double* wrms_result, int* wrms_done);
  /* Optional (NULL = not available): `fused` with a fused norm that ALSO produces the error weights of y itself,
     ewt_i = 1/(rtol*|y_i| + atol) (arkEwtSetSS, arkode.c:2932-2944) into ewt_out and sum_i (y_i ewt_i)^2 into the device
     double ewt_result -- requested speculatively when the launch looks like the closing stage of an adaptive step,
     whose y is the candidate y_{n+1}.  Sets *ewt_done = 1 if it did so (only together with *wrms_done = 1). */
  int (*fused_ewt)(void* self, b200_ctx* ctx, const double* y, int nterms, const double* c,
                   const int* src, const double* const* v, double* z, double* f_out,
                   const double* wrms_w, double* wrms_result, int* wrms_done,
                   double rtol, double atol, double* ewt_out, double* ewt_result, int* ewt_done);
  /* Optional (NULL / 0 = not available): temporal blocking of `nstages` <= chain_max
     consecutive STS stages  z_l = c[l][0] F(z_{l-1}) + c[l][1] z_{l-2} + c[l][2] yn +
     c[l][3] z_{l-1} + c[l][4] fn  (z_0 = x, z_{-1} = prev2) in one kernel; coeffs is
     [nstages][5], z_out[l] == NULL means "stage l+1 need not be stored". */
  int (*chain)(void* self, b200_ctx* ctx, int nstages, const double* x, const double* prev2,
               const double* yn, const double* fn, const double* coeffs, double* const* z_out,
               double* const* halos, const int* halo_valid);
  int chain_max;
  /* Optional (NULL = not available): the chain that BEGINS a step.  Stage 1 of the STS methods is
     z_1 = x + c_1 F(x) with x = y_n (arkode_lsrkstep.c:640 / :930) and F(x) is the f_n of every later stage, so the
     kernel produces f_n itself (stored to f_out) and streams only x: coeffs row 0 = { c_1, -, -, -, - }, rows
     1..nstages-1 as for `chain` with z_0 = y_n = x.  f_out == NULL: F(x) is stored already (the vector knows that the
     f_n it holds is F of exactly this x), the kernel only recomputes it.  halo_x / halo_x_valid: as halos[0] /
     halo_valid[0] of `chain`. */
  int (*chain_head)(void* self, b200_ctx* ctx, int nstages, const double* x, const double* coeffs,
                    double* const* z_out, double* f_out, double* halo_x, int halo_x_valid);
  /* > 0: the operator runs on one rank of a decomposition and `chain` needs a deep halo
     (this many doubles) for each of x, prev2, yn, fn.  The vector owns and caches the
     buffers per value; `chain` receives halos[4] and halo_valid[4] and must fill (exchange)
     the stale ones before launching.  0: halos == NULL. */
  int64_t halo_doubles;
  /* Optional (NULL = the vector allocates halo_doubles per value from its own pool): where a value's deep-halo
     buffer comes from and goes back to -- e.g. the slots of a peer-mapped ring that the neighbouring ranks write
     into (b200_peer_halo_*).  Every rank must see the same sequence of calls. */
  double* (*halo_alloc)(void* self);
  void (*halo_free)(void* self, double* halo);
  /* Optional (NULL = not available): the difference-quotient Jacobian-vector product around F in one pass,
       outer = 1:  z = ca*v + cb*( siginv*( F(sigma*v + y) - fy ) )      (arkLsATimes o arkLsDQJtimes)
       outer = 0:  z = siginv*( F(sigma*v + y) - fy )                     (arkLsDQJtimes, lsrkStep_DQJtimes)
     and, if dot_result != NULL, sum_i z_i v_i.  Return 0 = done, 1 = not handled here (the vector then evaluates
     the pieces one by one), < 0 = error. */
  int (*dq)(void* self, b200_ctx* ctx, const double* v, const double* y, const double* fy, double sigma,
            double siginv, int outer, double ca, double cb, double* z, double* dot_result);
} B200RhsOp;

/* Create a vector: local_length entries on this rank's GPU, global_length overall
   (N_VWrmsNorm divides by it, nvector_parallel.c:729; N_VGetLength returns it). */
SUNDIALS_EXPORT N_Vector N_VNew_B200(b200_ctx* ctx, sunindextype local_length,
                                     sunindextype global_length, SUNContext sunctx);

SUNDIALS_EXPORT b200_ctx* N_VGetContext_B200(N_Vector v);
SUNDIALS_EXPORT sunindextype N_VGetLocalLength_B200(N_Vector v);

/* Read access to the (materialised) device data. */
/* 1 once a device operation behind any vector of this process has failed.  The failure is printed once; from then on
   every vector operation returns immediately (fused operations SUN_ERR_EXT_FAIL, which LSRKStep maps to
   ARK_VECTOROP_ERR; reductions NaN; accessors NULL / -1), so the integrator comes back with an error code instead of
   the process being aborted.  There is no host fallback. */
SUNDIALS_EXPORT int N_VDeviceFailed_B200(void);
SUNDIALS_EXPORT const double* N_VGetDeviceArrayPointer_B200(N_Vector v);
/* Give v a fresh, writable device buffer (previous value is dropped, not copied). */
SUNDIALS_EXPORT double* N_VGetDeviceArrayPointerForWrite_B200(N_Vector v);
/* Explicit host <-> device copies (n = local length). */
SUNDIALS_EXPORT int N_VCopyFromHost_B200(N_Vector v, const double* host);
SUNDIALS_EXPORT int N_VCopyToHost_B200(N_Vector v, double* host);

/* Called from an ARKRhsFn: f := F(y), deferred.  op must outlive the vectors. */
SUNDIALS_EXPORT int N_VSetDeferredRhs_B200(N_Vector f, const B200RhsOp* op, N_Vector y);
/* 1 if v currently holds a deferred value (tests / diagnostics) */
SUNDIALS_EXPORT int N_VIsDeferred_B200(N_Vector v);
/* Turn lazy fusion off (every RHS materialises immediately) / on; default on. */
SUNDIALS_EXPORT void N_VSetLazyFusion_B200(int on);
/* Temporal blocking depth: how many consecutive STS stages one kernel launch may cover
   (1 = off, default 4, at most B200_MAX_CHAIN and what the RHS operator supports).
   Results do not depend on it (bit-identical). */
SUNDIALS_EXPORT void N_VSetStageChain_B200(int depth);
SUNDIALS_EXPORT int N_VGetStageChain_B200(void);

/* statistics since process start: fused launches, aliased copies, device buffers */
typedef struct B200VecStats
{
  long fused_launches;    /* RHS evaluations realised inside a fused (stencil+combination) kernel */
  long plain_rhs_launches;/* deferred values materialised on their own */
  long aliased_copies;    /* N_VScale(1,x,z) turned into a handle share */
  long buffers_allocated; /* cudaMalloc'd vector buffers */
  long wrms_fused;        /* WRMS norms answered from a fused partial */
  long chain_launches;    /* temporally blocked launches (>= 2 stages each) */
  long chain_stages;      /* stages covered by those launches */
  long dq_fused;          /* difference-quotient matvecs evaluated in one stencil pass (B200RhsOp::dq) */
  long ew_fused;          /* elementwise results produced inside a reduction kernel (lin2+wsqr, prod+dot) */
} B200VecStats;
SUNDIALS_EXPORT void N_VGetStats_B200(B200VecStats* s);

/* Lazy fusion, the chain depth, the arithmetic flavour and the chain-kernel variant are PROCESS-wide switches.  A
   session (b200_d2d_create / b200_adr_create) registers what it runs with; a second session created while the first is
   alive must ask for the same values (-1 = "whatever is set"), else N_VAcquireSettings_B200 returns -1 and the
   constructor fails instead of silently changing the other session's arithmetic or fusion.  Returns 0 and applies the
   values otherwise.  Every successful acquire is paired with one N_VReleaseSettings_B200. */
SUNDIALS_EXPORT int N_VAcquireSettings_B200(int lazy, int chain_depth, int fma_arithmetic, int chain_variant);
SUNDIALS_EXPORT void N_VReleaseSettings_B200(void);

#ifdef __cplusplus
}
#endif
#endif
