/* b200_adr2d.h -- C-ABI of the adr 2-D (Brusselator advection-diffusion-reaction)
 * problem layer in libb200sts_sundials.so, on top of nvector_b200.h and an UNMODIFIED
 * SUNDIALS ARKODE (SplittingStep / MRIStep / LSRKStep / ARKStep / ERKStep).
 *
 * Mirrors /root/reference/adr/advection_diffusion_reaction_2d.{cpp,hpp}: the same
 * command-line options (ReadInputs, ...2d.hpp:516-604), the same integrator set-ups
 *   --integrator 0  SetupERK     ...2d.cpp:308-355    ERKStep on adv+diff+react
 *   --integrator 1  SetupARK     ...2d.cpp:358-713    IMEX ARKStep, GMRES on diffusion
 *   --integrator 2  SetupExtSTS  ...2d.cpp:715-1120   MRIStep + LSRKStep inner stepper
 *   --integrator 3  SetupStrang  ...2d.cpp:1122-1333  SplittingStep: LSRKStep . ARS(2,2,2)
 * and the same callbacks (...2d.cpp:1406-1699), which enqueue sm_100a kernels through
 * b200_sts.h instead of looping over the grid on the host.  `--implicit-reaction` is
 * available with the two STS integrators (2, 3): the reference's SUNBandMatrix +
 * SUNLinSol_Band become the block-diagonal device pair of b200_blockdiag.h; with
 * --integrator 1 (BBD preconditioner on host arrays) it is rejected, as is `--calc_error`.
 *
 * The callbacks themselves -- ARKRhsFn / ARKDomEigFn signatures, SUNDIALS types -- are
 * declared in b200_callbacks.h.
 */
#ifndef B200_ADR2D_H
#define B200_ADR2D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_adr b200_adr; /* one configured problem + integrator on one GPU */

typedef struct b200_adr_stats
{
  double t;                  /* current time */
  double evolve_seconds;     /* host wall time inside ARKodeEvolve ("Total solve time") */
  long steps, step_attempts; /* outer integrator */
  long rhs_evals_explicit;   /* outer fe evaluations (ERK/ARK/MRI slow) */
  long rhs_evals_implicit;   /* outer fi evaluations + linear-solver RHS evaluations (ARK) */
  long lsrk_steps;           /* LSRKStep partition / inner stepper */
  long lsrk_rhs_evals;
  long lsrk_max_stages;
  long ark_steps;            /* ARKStep partition of the Strang splitting */
  long ark_rhs_evals;
  long fused_launches, plain_rhs_launches, aliased_copies, buffers_allocated;
  uint64_t kernel_launches;
  int64_t nx, ny, neq;
  /* the implicitly treated partition (IMEX ARK: diffusion; --implicit-reaction: the reaction) */
  long ark_rhs_evals_implicit; /* Strang: fi evaluations of the ARKStep partition */
  long nls_iters, ls_setups, jac_evals; /* Newton iterations, linear-solver set-ups, Jacobian evaluations of the
                                           integrator that owns the implicit partition */
} b200_adr_stats;

/* Build the problem from reference-style arguments, e.g.
     {"--nx","2048","--ny","2048","--integrator","3","--sts_method","0","--fixed_h","1e-3"}.
   device: CUDA device ordinal; stream: an existing cudaStream_t or NULL. */
int b200_adr_create(int argc, const char* const* argv, int device, void* stream, b200_adr** out);
int b200_adr_destroy(b200_adr* p);
/* ARKodeEvolve(mem, tout, y, &t, ARK_NORMAL) */
int b200_adr_evolve(b200_adr* p, double tout);
/* nsteps calls of ARKodeEvolve(..., ARK_ONE_STEP) */
int b200_adr_step(b200_adr* p, int nsteps);
/* state <-> host: 2*nx*ny doubles, interleaved [u,v] (...2d.hpp:44-45) */
int b200_adr_get_state(b200_adr* p, double* host);
int b200_adr_set_state(b200_adr* p, const double* host, double t);
int b200_adr_get_stats(b200_adr* p, b200_adr_stats* s);
/* the statistics block of the reference main() (...2d.cpp:252-265) */
int b200_adr_print_stats(b200_adr* p);
/* The whole reference main(): parse, set up, evolve over nout outputs, write
   solution.dat (WriteOutput, ...2d.hpp:794-829), print stats. */
int b200_adr_main(int argc, char** argv);

#ifdef __cplusplus
}
#endif
#endif
