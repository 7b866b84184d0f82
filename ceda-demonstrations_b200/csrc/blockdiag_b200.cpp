// blockdiag_b200.cpp -- SUNMatrix / SUNLinearSolver pair for the implicit reaction partition of the adr 2-D
// driver, on device data.
//
// The reference attaches SUNBandMatrix(neq, 2, 2) + SUNLinSol_Band to the ARKStep / MRIStep integrator that treats
// the Brusselator reaction implicitly (adr/advection_diffusion_reaction_2d.cpp:820-835, :1213-1226) and fills the
// matrix in J_reaction (:1523-1551).  ARKODE's direct-solver interface (SUN/src/arkode/arkode_ls.c, arkLsLinSys /
// arkLsSetup / arkLsSolve) touches the pair only through the generic operations
//     SUNMatClone / SUNMatZero / SUNMatCopy / SUNMatScaleAddI,  SUNLinSolSetup / SUNLinSolSolve,
// so a custom pair drops in without touching ARKODE.  With the two species interleaved the matrix is block diagonal
// (2 x 2 per grid point) and the band LU never leaves a block: the device kernels (csrc/react_kernels.cuh) repeat
// its operations per block with the same roundings, so the Newton iterates equal the reference's.
//
// Host C++ only (no CUDA headers): everything goes through the C-ABI of b200_sts.h.

#include <sundials/sundials_core.h>
#include <sundials/sundials_linearsolver.h>
#include <sundials/sundials_matrix.h>

#include <cstdio>
#include <cstdlib>

#include "b200_blockdiag.h"
#include "b200_sts.h"
#include "nvector_b200.h"

namespace {

struct Blk2Mat
{
  b200_ctx* ctx;
  int64_t npts;
  double* d; // 4 doubles per grid point
};
inline Blk2Mat* M(SUNMatrix A) { return static_cast<Blk2Mat*>(A->content); }

struct Blk2Sol
{
  b200_ctx* ctx;
  int64_t npts;
  double* piv; // one double per grid point (1 = rows swapped)
  long long last_flag;
};
inline Blk2Sol* S(SUNLinearSolver L) { return static_cast<Blk2Sol*>(L->content); }

SUNMatrix_ID mat_getid(SUNMatrix) { return SUNMATRIX_CUSTOM; }

void mat_destroy(SUNMatrix A)
{
  if (!A) return;
  if (A->content)
  {
    if (M(A)->d) b200_free(M(A)->ctx, M(A)->d);
    delete M(A);
  }
  SUNMatFreeEmpty(A);
}

SUNMatrix mat_clone(SUNMatrix A) { return SUNMatrix_B200Block2(M(A)->ctx, M(A)->npts, A->sunctx); }

SUNErrCode mat_zero(SUNMatrix A)
{
  return b200_const(M(A)->ctx, 0.0, M(A)->d, 4 * M(A)->npts) ? SUN_ERR_EXT_FAIL : SUN_SUCCESS;
}

SUNErrCode mat_copy(SUNMatrix A, SUNMatrix B)
{ // B = A (1.0 * a is exact)
  const double one    = 1.0;
  const double* vp[1] = {M(A)->d};
  return b200_lincomb(M(A)->ctx, 1, &one, vp, M(B)->d, 4 * M(A)->npts) ? SUN_ERR_EXT_FAIL : SUN_SUCCESS;
}

SUNErrCode mat_scaleaddi(sunrealtype c, SUNMatrix A)
{
  return b200_blk2_scale_add_i(M(A)->ctx, c, M(A)->d, M(A)->npts) ? SUN_ERR_EXT_FAIL : SUN_SUCCESS;
}

SUNErrCode mat_space(SUNMatrix A, long int* lenrw, long int* leniw)
{
  *lenrw = (long int)(4 * M(A)->npts);
  *leniw = 2;
  return SUN_SUCCESS;
}

SUNLinearSolver_Type ls_gettype(SUNLinearSolver) { return SUNLINEARSOLVER_DIRECT; }
SUNLinearSolver_ID ls_getid(SUNLinearSolver) { return SUNLINEARSOLVER_CUSTOM; }
SUNErrCode ls_initialize(SUNLinearSolver L)
{
  S(L)->last_flag = 0;
  return SUN_SUCCESS;
}

// SUNLinSolSetup_Band (SUN/src/sunlinsol/band/sunlinsol_band.c): LU in place, zero pivot -> SUNLS_LUFACT_FAIL
int ls_setup(SUNLinearSolver L, SUNMatrix A)
{
  if (!A || SUNMatGetID(A) != SUNMATRIX_CUSTOM || M(A)->npts != S(L)->npts) return SUN_ERR_ARG_INCOMPATIBLE;
  long long info = 0;
  if (b200_blk2_factor(S(L)->ctx, M(A)->d, S(L)->piv, S(L)->npts, &info))
  {
    fprintf(stderr, "SUNLinSol_B200Block2: %s\n", b200_last_error());
    return SUN_ERR_EXT_FAIL;
  }
  S(L)->last_flag = info;
  return info > 0 ? SUNLS_LUFACT_FAIL : SUN_SUCCESS;
}

// SUNLinSolSolve_Band: x = b, then the triangular solves on x
int ls_solve(SUNLinearSolver L, SUNMatrix A, N_Vector x, N_Vector b, sunrealtype)
{
  const double* bd = N_VGetDeviceArrayPointer_B200(b);
  double* xd       = (x == b) ? const_cast<double*>(bd) : N_VGetDeviceArrayPointerForWrite_B200(x);
  if (b200_blk2_solve(S(L)->ctx, M(A)->d, S(L)->piv, bd, xd, S(L)->npts))
  {
    fprintf(stderr, "SUNLinSol_B200Block2: %s\n", b200_last_error());
    return SUN_ERR_EXT_FAIL;
  }
  S(L)->last_flag = 0;
  return SUN_SUCCESS;
}

sunindextype ls_lastflag(SUNLinearSolver L) { return (sunindextype)S(L)->last_flag; }

SUNErrCode ls_space(SUNLinearSolver L, long int* lenrw, long int* leniw)
{
  *lenrw = 0;
  *leniw = (long int)S(L)->npts;
  return SUN_SUCCESS;
}

SUNErrCode ls_free(SUNLinearSolver L)
{
  if (!L) return SUN_SUCCESS;
  if (L->content)
  {
    if (S(L)->piv) b200_free(S(L)->ctx, S(L)->piv);
    delete S(L);
    L->content = nullptr;
  }
  SUNLinSolFreeEmpty(L);
  return SUN_SUCCESS;
}

} // namespace

extern "C" SUNMatrix SUNMatrix_B200Block2(b200_ctx* ctx, int64_t npts, SUNContext sunctx)
{
  if (!ctx || npts < 1) return nullptr;
  SUNMatrix A = SUNMatNewEmpty(sunctx);
  if (!A) return nullptr;
  A->ops->getid     = mat_getid;
  A->ops->clone     = mat_clone;
  A->ops->destroy   = mat_destroy;
  A->ops->zero      = mat_zero;
  A->ops->copy      = mat_copy;
  A->ops->scaleaddi = mat_scaleaddi;
  A->ops->space     = mat_space;
  Blk2Mat* m = new Blk2Mat();
  m->ctx = ctx; m->npts = npts; m->d = nullptr;
  A->content = m;
  if (b200_malloc(ctx, 4 * npts, &m->d) || b200_const(ctx, 0.0, m->d, 4 * npts))
  {
    mat_destroy(A);
    return nullptr;
  }
  return A;
}

extern "C" double* SUNMatrix_B200Block2_Data(SUNMatrix A) { return (A && A->content) ? M(A)->d : nullptr; }
extern "C" int64_t SUNMatrix_B200Block2_Points(SUNMatrix A) { return (A && A->content) ? M(A)->npts : 0; }

extern "C" SUNLinearSolver SUNLinSol_B200Block2(N_Vector y, SUNMatrix A, SUNContext sunctx)
{
  if (!y || !A || SUNMatGetID(A) != SUNMATRIX_CUSTOM || N_VGetLocalLength_B200(y) != 2 * M(A)->npts) return nullptr;
  SUNLinearSolver L = SUNLinSolNewEmpty(sunctx);
  if (!L) return nullptr;
  L->ops->gettype    = ls_gettype;
  L->ops->getid      = ls_getid;
  L->ops->initialize = ls_initialize;
  L->ops->setup      = ls_setup;
  L->ops->solve      = ls_solve;
  L->ops->lastflag   = ls_lastflag;
  L->ops->space      = ls_space;
  L->ops->free       = ls_free;
  Blk2Sol* s = new Blk2Sol();
  s->ctx = M(A)->ctx; s->npts = M(A)->npts; s->piv = nullptr; s->last_flag = 0;
  L->content = s;
  if (b200_malloc(s->ctx, s->npts, &s->piv))
  {
    ls_free(L);
    return nullptr;
  }
  return L;
}
