/* b200_diffusion2d.h -- C-ABI of libb200sts_sundials.so: the diffusion_2D problem
 * layer (callbacks + driver session) on top of nvector_b200.h and an UNMODIFIED
 * SUNDIALS ARKODE.
 *
 * It mirrors /root/reference/diffusion_2D: the same command-line options
 * (diffusion_2D.cpp:42-163, main.cpp:556-687, diffusion_2D.cpp:651-677), the same
 * ARKODE call sequence (main.cpp:176-470) and the same callbacks
 *   diffusion()  ARKRhsFn          diffusion_2D.cpp:23-35   -> b200_diffusion_rhs
 *   dom_eig()    ARKDomEigFn       main.cpp:536-550         -> b200_diffusion_domeig
 *   PSetup/PSolve ARKLsPrec*Fn     preconditioner_jacobi.cpp -> b200_diffusion_psetup/psolve
 * so a maintainer can register them in the reference main.cpp unchanged (see
 * INTEGRATION.md).  The session API below is what bench.py / the tests bind with
 * ctypes; `b200_d2d_main` is the executable's main().
 */
#ifndef B200_DIFFUSION2D_H
#define B200_DIFFUSION2D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_d2d b200_d2d; /* one configured problem + integrator on one rank/GPU */

typedef struct b200_d2d_stats
{
  double t;                 /* current time */
  double h_last;            /* last step size */
  double urms;              /* sqrt(u.u / nx / ny), UserOutput::write diffusion_2D.cpp:801 */
  double evolve_seconds;    /* host wall time inside ARKodeEvolve (the reference's simtime) */
  double spectral_radius;   /* current rho */
  long steps, step_attempts, err_test_fails;
  long rhs_evals;           /* ARKodeGetNumRhsEvals (explicit partition; implicit for dirk) */
  long dee_rhs_evals;       /* LSRKStepGetNumDomEigEstRhsEvals */
  long dom_eig_updates;
  long max_stages;
  long lin_iters, lin_rhs_evals, prec_solves, nonlin_iters; /* dirk path */
  long fused_launches, plain_rhs_launches, aliased_copies, wrms_fused, buffers_allocated;
  uint64_t kernel_launches; /* b200_launch_count() */
  int64_t nx, ny, nx_loc, ny_loc, is, js;
  int npx, npy, rank, nranks;
  long chain_launches, chain_stages; /* temporally blocked launches and the stages they covered */
  long dq_fused, ew_fused;           /* difference-quotient matvecs in one stencil pass; elementwise results produced
                                        inside the reduction kernel that consumes them (implicit path) */
} b200_d2d_stats;

/* Build the problem from reference-style arguments, e.g.
     {"--nx","16384","--ny","16384","--integrator","rkc","--fixedstep","1e-4","--tf","1e-3"}.
   rank/nranks: this process' position in the 2-D block decomposition (one process
   per GPU); nccl_id: 128 bytes from b200_comm_unique_id (NULL when nranks == 1);
   device: CUDA device ordinal; stream: an existing cudaStream_t or NULL. */
int b200_d2d_create(int argc, const char* const* argv, int rank, int nranks,
                    const unsigned char* nccl_id, int device, void* stream, b200_d2d** out);
int b200_d2d_destroy(b200_d2d* p);

/* ARKodeEvolve(mem, tout, u, &t, ARK_NORMAL) with the stop time moved to tout. */
int b200_d2d_evolve(b200_d2d* p, double tout);
/* nsteps calls of ARKodeEvolve(..., ARK_ONE_STEP). */
int b200_d2d_step(b200_d2d* p, int nsteps);
/* local sub-domain state <-> host (row-major nx_loc*ny_loc doubles). */
int b200_d2d_get_state(b200_d2d* p, double* host);
/* overwrite the state from host memory and ARKodeReset to time t */
int b200_d2d_set_state(b200_d2d* p, const double* host, double t);
/* A stream of `nbatch` independent states through the same integrator: for each i, the state is
   overwritten from host_in[i] (pinned), ARKodeReset to time t, `nsteps` steps are taken and the
   result goes to host_out[i] (pinned).  Uploads and downloads are double-buffered on their own copy
   streams (b200_pipe_*), so the copies of neighbouring batches overlap the integration; results are
   identical to set_state / step / get_state per batch.  Returns after all results have landed. */
int b200_d2d_run_batches(b200_d2d* p, int nbatch, const double* const* host_in, double* const* host_out,
                         double t, int nsteps);
/* COLLECTIVE when nranks > 1 (urms is an all-reduced dot product): call on every rank. */
int b200_d2d_get_stats(b200_d2d* p, b200_d2d_stats* s);
/* print ARKodePrintAllStats exactly as the reference main.cpp:486 does */
int b200_d2d_print_stats(b200_d2d* p);
/* local extents for a given decomposition without creating a problem */
int b200_d2d_local_extent(int64_t nx, int64_t ny, int rank, int nranks, int npx, int npy,
                          int64_t* is, int64_t* nx_loc, int64_t* js, int64_t* ny_loc,
                          int* npx_out, int* npy_out);

/* The whole reference main(): parse, set up, evolve over nout outputs, print stats.
   Rank / world size come from RANK / WORLD_SIZE (torchrun) or B200_RANK / B200_NP;
   the NCCL id is exchanged through the file named by B200_NCCL_ID_FILE. */
int b200_d2d_main(int argc, char** argv);

#ifdef __cplusplus
}
#endif
#endif
