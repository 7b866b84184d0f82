/* b200_callbacks.h -- the user-supplied ARKODE callbacks of the two reference drivers,
 * re-implemented for N_Vector_B200.  Same signatures (SUNDIALS types), same return
 * convention (0 ok, <0 fatal), registered with the same ARKODE calls -- see
 * INTEGRATION.md.  `user_data` must be the B200 problem object the matching session
 * constructor built (b200_d2d_create / b200_adr_create), which is what those
 * constructors pass to ARKodeSetUserData.
 *
 * None of the RHS callbacks launches a kernel: each marks `f` as the deferred value
 * "operator applied to y" (N_VSetDeferredRhs_B200); the vector fuses the operator into
 * the N_VLinearCombination / N_VLinearSum that consumes it (nvector_b200.h).
 */
#ifndef B200_CALLBACKS_H
#define B200_CALLBACKS_H

#include <sundials/sundials_matrix.h>
#include <sundials/sundials_nvector.h>
#include <sundials/sundials_types.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- diffusion_2D ---------------------------------------------------------------- */
/* The `user_data` object of the b200_diffusion_* callbacks, built from the fields of the reference's own UserData
 * (diffusion_2D/diffusion_2D.hpp:66-216) AFTER its setup(): grid, domain, coefficients and the position of this rank
 * in the process grid.  This is what lets the reference's main.cpp keep its UserData / UserOptions / UserOutput and
 * its whole ARKODE call sequence and swap only the vector constructor, the callbacks and the user_data pointer
 * (tests/native/patch_reference_main.py builds exactly that program).
 * ctx: the device context the vectors were created on (with b200_comm_init done when nranks > 1).
 * diag: the vector PSetup fills and PSolve multiplies by (udata.diag = N_VClone(u), main.cpp:227), or NULL. */
struct b200_ctx;
typedef struct b200_d2d_problem b200_d2d_problem;
int b200_d2d_problem_create(struct b200_ctx* ctx, long long nx, long long ny, double xl, double xu, double yl, double yu,
                            double kx, double ky, int inhomogeneous, int npx, int npy, int rank, int nranks,
                            N_Vector diag, b200_d2d_problem** out);
void* b200_d2d_problem_user_data(b200_d2d_problem* p); /* pass to ARKodeSetUserData */
int b200_d2d_problem_destroy(b200_d2d_problem* p);

/* ARKRhsFn: diffusion(), diffusion_2D/diffusion_2D.cpp:23-35 -> laplacian(), diffusion.cpp:9-209 */
int b200_diffusion_rhs(sunrealtype t, N_Vector u, N_Vector f, void* user_data);
/* ARKDomEigFn: dom_eig(), diffusion_2D/main.cpp:536-550 */
int b200_diffusion_domeig(sunrealtype t, N_Vector y, N_Vector fn, sunrealtype* lambdaR,
                          sunrealtype* lambdaI, void* user_data, N_Vector temp1,
                          N_Vector temp2, N_Vector temp3);
/* ARKLsPrecSetupFn / ARKLsPrecSolveFn: diffusion_2D/preconditioner_jacobi.cpp:9-46 / 49-62 */
int b200_diffusion_psetup(sunrealtype t, N_Vector u, N_Vector f, sunbooleantype jok,
                          sunbooleantype* jcurPtr, sunrealtype gamma, void* user_data);
int b200_diffusion_psolve(sunrealtype t, N_Vector u, N_Vector f, N_Vector r, N_Vector z,
                          sunrealtype gamma, sunrealtype delta, int lr, void* user_data);

/* ---- adr 2-D: adr/advection_diffusion_reaction_2d.cpp ------------------------------ */
int b200_adr_f_advection(sunrealtype t, N_Vector y, N_Vector f, void* user_data);      /* :1406-1445 */
int b200_adr_f_diffusion(sunrealtype t, N_Vector y, N_Vector f, void* user_data);      /* :1448-1491 */
int b200_adr_f_reaction(sunrealtype t, N_Vector y, N_Vector f, void* user_data);       /* :1494-1520 */
int b200_adr_f_adv_diff(sunrealtype t, N_Vector y, N_Vector f, void* user_data);       /* :1562-1579 */
int b200_adr_f_adv_react(sunrealtype t, N_Vector y, N_Vector f, void* user_data);      /* :1602-1619 */
int b200_adr_f_diff_react(sunrealtype t, N_Vector y, N_Vector f, void* user_data);     /* :1582-1599 */
int b200_adr_f_adv_diff_react(sunrealtype t, N_Vector y, N_Vector f, void* user_data); /* :1622-1646 */
int b200_adr_f_diffusion_forcing(sunrealtype t, N_Vector y, N_Vector f, void* user_data); /* :1649-1663 */
/* ARKLsJacFn: J_reaction(), :1523-1551 -- J must be a SUNMatrix_B200Block2 (b200_blockdiag.h), which stands in for the
   reference's SUNBandMatrix(neq, 2, 2) when the reaction is implicit */
int b200_adr_J_reaction(sunrealtype t, N_Vector y, N_Vector fy, SUNMatrix J, void* user_data, N_Vector tmp1,
                        N_Vector tmp2, N_Vector tmp3);
/* ARKDomEigFn: diffusion_domeig(), :1666-1679 */
int b200_adr_domeig(sunrealtype t, N_Vector y, N_Vector fn, sunrealtype* lambdaR,
                    sunrealtype* lambdaI, void* user_data, N_Vector temp1, N_Vector temp2,
                    N_Vector temp3);

#ifdef __cplusplus
}
#endif
#endif
