/* b200_blockdiag.h -- SUNMatrix / SUNLinearSolver pair on device data for the implicit reaction partition of the
 * adr 2-D driver: replaces SUNBandMatrix(neq, 2, 2) + SUNLinSol_Band of
 * adr/advection_diffusion_reaction_2d.cpp:820-835 (SetupExtSTS) and :1213-1226 (SetupStrang).
 *
 * The matrix is block diagonal, one 2 x 2 block per grid point (species interleaved, ...2d.hpp:44-45); storage is
 * 4 doubles per grid point { dU/du, dV/du, dU/dv, dV/dv } in device memory.  ARKODE uses the pair through the
 * generic SUNMatrix / SUNLinearSolver operations only (clone, zero, copy, scaleaddI; setup = LU, solve), which the
 * kernels of csrc/react_kernels.cuh realise with the operation order and roundings of the band routines
 * (SUN/src/sundials/sundials_band.c), so Newton iterates match the reference.
 */
#ifndef B200_BLOCKDIAG_H
#define B200_BLOCKDIAG_H

#include <stdint.h>
#include <sundials/sundials_linearsolver.h>
#include <sundials/sundials_matrix.h>
#include <sundials/sundials_nvector.h>

#ifdef __cplusplus
extern "C" {
#endif

struct b200_ctx;
/* npts = nx*ny grid points (a (2 npts) x (2 npts) matrix) */
SUNMatrix SUNMatrix_B200Block2(struct b200_ctx* ctx, int64_t npts, SUNContext sunctx);
double* SUNMatrix_B200Block2_Data(SUNMatrix A); /* device pointer, 4*npts doubles */
int64_t SUNMatrix_B200Block2_Points(SUNMatrix A);
/* direct solver: setup = in-place LU with partial pivoting inside each block, solve = forward / back substitution */
SUNLinearSolver SUNLinSol_B200Block2(N_Vector y, SUNMatrix A, SUNContext sunctx);

#ifdef __cplusplus
}
#endif
#endif
