/* sts_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement (plain C, sequential loops)
 * of the reference algorithm on the explicit super-time-stepping hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library; nothing in ceda-demonstrations_b200/ links, imports or executes it.
 *
 * Parity pin: PINNED.  tests/test_oracle_pin.py checks this restatement against
 *   (a) SUNDIALS' own golden stage logs test_logging_arkode_lsrkstep_lvl5_{0,1,2}.out
 *       (fixtures under tests/golden/, extracted by tests/golden/make_golden.py), and
 *   (b) outputs of the unmodified reference binary oracle/_ref/diffusion_2D_ref
 *       (fixed-step runs: bit-identical final states; fixtures under tests/golden/).
 *
 * Every function cites the reference lines it follows; paths are relative to
 * /root/reference and SUN = deps/sundials.
 */
#ifndef STS_ORACLE_H
#define STS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- N_Vector operations: SUN/src/nvector/parallel/nvector_parallel.c ---- */
void orc_linear_sum(double a, const double* x, double b, const double* y, double* z, int64_t n); /* :424-517 (+helpers :1771-1950); aliasing-aware */
void orc_const(double c, double* z, int64_t n);                               /* :519 */
void orc_prod(const double* x, const double* y, double* z, int64_t n);         /* :534 */
void orc_div(const double* x, const double* y, double* z, int64_t n);          /* :551 */
void orc_scale(double c, const double* x, double* z, int64_t n);               /* :568 */
void orc_abs(const double* x, double* z, int64_t n);                           /* :594 */
void orc_inv(const double* x, double* z, int64_t n);                           /* :610 */
void orc_addconst(const double* x, double b, double* z, int64_t n);            /* :626 */
double orc_dot(const double* x, const double* y, int64_t n);                   /* :642 */
double orc_maxnorm(const double* x, int64_t n);                                /* :672 */
double orc_wsqrsum(const double* x, const double* w, int64_t n);               /* :700 */
double orc_wrmsnorm(const double* x, const double* w, int64_t n, int64_t nglobal); /* :721 */
double orc_min(const double* x, int64_t n);                                    /* :769 */
double orc_l1norm(const double* x, int64_t n);                                 /* :815 */
/* generic N_VLinearCombination fallback, SUN/src/sundials/sundials_nvector.c:546-569 */
void orc_linear_combination(int nvec, const double* c, const double* const* X, double* z, int64_t n);
/* arkEwtSetSS, SUN/src/arkode/arkode.c:2932-2944 (tmp is ARKODE's tempv1) */
void orc_ewt_ss(const double* y, double rtol, double atol, double* tmp, double* ewt, int64_t n);

/* ---- diffusion_2D problem: diffusion_2D/ ---- */
typedef struct orc_grid
{
  double kx, ky;      /* diffusion_2D.hpp:71-72 */
  int inhomogeneous;  /* :73 */
  double xl, yl;      /* :79-80 */
  double dx, dy;      /* :92-93, recomputed diffusion_2D.cpp:159-160 */
  int64_t nx_loc, ny_loc, is, js; /* local extents, diffusion_2D.cpp:286-317 */
} orc_grid;

double orc_coeff_x(double x, const orc_grid* g); /* Diffusion_Coeff_X diffusion_2D.cpp:887-891 */
double orc_coeff_y(double y, const orc_grid* g); /* Diffusion_Coeff_Y :893-897 */
/* the four per-index face-coefficient tables diffusion.cpp:36-46 evaluates per cell */
void orc_coeff_tables(const orc_grid* g, double* cxw, double* cxe, double* cys, double* cyn);
/* laplacian(): diffusion.cpp:9-209.  W/E/S/N are the received halo buffers; a NULL
   buffer means a single periodic rank in that direction (the exchange of
   diffusion_2D.cpp:400-584 then delivers the rank's own opposite edge). */
void orc_laplacian(const orc_grid* g, const double* u, double* f, const double* W,
                   const double* E, const double* S, const double* N);
void orc_pack(const orc_grid* g, const double* u, double* Ws, double* Es, double* Ss, double* Ns); /* buffers.cpp:20-43 */
void orc_initial(const orc_grid* g, double* u);                                /* initial.cpp:20-48 */
void orc_jacobi_setup(const orc_grid* g, double gamma, double* diag);          /* preconditioner_jacobi.cpp:9-46 */
double orc_dom_eig(const orc_grid* g);                                         /* main.cpp:536-550 */
/* 2-D block decomposition, diffusion_2D.cpp:243-317 (dims as MPI_Dims_create) */
void orc_dims_create(int np, int dims[2]);
void orc_decompose(int64_t n, int nproc, int coord, int64_t* start, int64_t* count);

/* ---- LSRKStep single-step recurrences: SUN/src/arkode/arkode_lsrkstep.c ---- */
typedef int (*orc_rhs_fn)(double t, const double* y, double* f, void* user);
typedef struct orc_step_ws
{
  int64_t n, nglobal;
  double *yn, *fn, *ycur, *tempv1, *tempv2, *tempv3, *ewt; /* ARKODE work vectors */
  int fixedstep;
  long nfe;
} orc_step_ws;
/* Each takes yn/fn/ewt as inputs (fn = f(tn,yn) must be current), writes ycur (the
   new solution), leaves F(ycur) in tempv2 where the reference does, and returns the
   stage count used (<0 on error).  *dsm receives the WRMS error estimate (adaptive). */
int orc_stages_rkc(double h, double spectral_radius);                          /* :563-565 */
int orc_stages_rkl(double h, double spectral_radius);                          /* :872-878 */
int orc_step_rkc(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double spectral_radius, double* dsm); /* :534-818 */
int orc_step_rkl(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double spectral_radius, double* dsm); /* :846-1106 */
int orc_step_ssps2(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, int stages, double* dsm);          /* :1128-1300 */
int orc_step_ssps3(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, int stages, double* dsm);          /* :1326-1584 */
int orc_step_ssp43(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double* dsm);                       /* :1610-1796 */
int orc_step_ssp104(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double* dsm);                      /* :1816-2023 */

/* Fixed-step diffusion_2D run on one periodic rank: nsteps RKC (method 0) or RKL
   (method 1) steps of size h from the initial condition, analytic dom_eig with the
   1.01 safety factor (arkode_lsrkstep.c:2340), exactly the sequence ARKodeEvolve
   drives in fixed-step mode.  Returns the number of RHS evaluations. */
long orc_diffusion_fixed_run(const orc_grid* g, int method, double h, int nsteps, double* u);

/* ---- power iteration: SUN/src/sundomeigest/power/sundomeigest_power.c:261-330 ---- */
typedef int (*orc_atimes_fn)(void* user, const double* v, double* Av);
int orc_power_iteration(orc_atimes_fn A, void* user, double* V, double* q, int64_t n,
                        int num_warmups, int max_iters, double rel_tol,
                        double* lambdaR, int* iters);

/* ---- adr 2-D Brusselator: adr/advection_diffusion_reaction_2d.cpp ---- */
typedef struct orc_adr
{
  int64_t nx, ny;
  double dx, dy, cux, cuy, cvx, cvy, d, A, B;
} orc_adr;
void orc_adr_advection(const orc_adr* p, const double* y, double* f); /* :1406-1445 */
void orc_adr_diffusion(const orc_adr* p, const double* y, double* f); /* :1448-1491 */
void orc_adr_reaction(const orc_adr* p, const double* y, double* f);  /* :1494-1520 */
void orc_adr_adv_react(const orc_adr* p, const double* y, double* tmp, double* f); /* :1602-1619 */
double orc_adr_domeig(const orc_adr* p);                               /* :1666-1679 */
void orc_adr_ic(const orc_adr* p, double xl, double yl, double* y);    /* :1682-1699 */

#ifdef __cplusplus
}
#endif
#endif
