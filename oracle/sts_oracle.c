/* sts_oracle.c -- TEST INFRASTRUCTURE: see sts_oracle.h (parity pin: PINNED).
 *
 * Sequential C restatement of the reference's algorithm for the STS hot path.
 * Compiled with -O2 -ffp-contract=off so every operation rounds once, in source
 * order -- the arithmetic of the reference's default CPU build.
 * Paths in comments are relative to /root/reference; SUN = deps/sundials.
 */
#include "sts_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX(x, y, n) ((n) * (y) + (x)) /* diffusion_2D/diffusion_2D.hpp:55 */

/* ===================== N_Vector ops (nvector_parallel.c) ===================== */

/* Vaxpy_Parallel :1909-1926 */
static void v_axpy(double a, const double* x, double* y, int64_t n)
{
  int64_t i;
  if (a == 1.0) { for (i = 0; i < n; i++) y[i] += x[i]; return; }
  if (a == -1.0) { for (i = 0; i < n; i++) y[i] -= x[i]; return; }
  for (i = 0; i < n; i++) y[i] += a * x[i];
}

/* N_VLinearSum_Parallel :424-517: the special cases are tested in this order */
void orc_linear_sum(double a, const double* x, double b, const double* y, double* z, int64_t n)
{
  int64_t i;
  if (b == 1.0 && z == y) { v_axpy(a, x, z, n); return; }
  if (a == 1.0 && z == x) { v_axpy(b, y, z, n); return; }
  if (a == 1.0 && b == 1.0) { for (i = 0; i < n; i++) z[i] = x[i] + y[i]; return; }          /* VSum :1791 */
  if ((a == 1.0 && b == -1.0) || (a == -1.0 && b == 1.0))
  {                                                                                          /* VDiff :1807 */
    const double* v1 = (a == 1.0 && b == -1.0) ? y : x;
    const double* v2 = (a == 1.0 && b == -1.0) ? x : y;
    for (i = 0; i < n; i++) z[i] = v2[i] - v1[i];
    return;
  }
  if (a == 1.0 || b == 1.0)
  {                                                                                          /* VLin1 :1875 */
    double c = (a == 1.0) ? b : a;
    const double* v1 = (a == 1.0) ? y : x;
    const double* v2 = (a == 1.0) ? x : y;
    for (i = 0; i < n; i++) z[i] = (c * v1[i]) + v2[i];
    return;
  }
  if (a == -1.0 || b == -1.0)
  {                                                                                          /* VLin2 :1892 */
    double c = (a == -1.0) ? b : a;
    const double* v1 = (a == -1.0) ? y : x;
    const double* v2 = (a == -1.0) ? x : y;
    for (i = 0; i < n; i++) z[i] = (c * v1[i]) - v2[i];
    return;
  }
  if (a == b) { for (i = 0; i < n; i++) z[i] = a * (x[i] + y[i]); return; }                  /* VScaleSum :1839 */
  if (a == -b) { for (i = 0; i < n; i++) z[i] = a * (x[i] - y[i]); return; }                 /* VScaleDiff :1855 */
  for (i = 0; i < n; i++) z[i] = (a * x[i]) + (b * y[i]);                                    /* :514 */
}

void orc_const(double c, double* z, int64_t n) { for (int64_t i = 0; i < n; i++) z[i] = c; }
void orc_prod(const double* x, const double* y, double* z, int64_t n) { for (int64_t i = 0; i < n; i++) z[i] = x[i] * y[i]; }
void orc_div(const double* x, const double* y, double* z, int64_t n) { for (int64_t i = 0; i < n; i++) z[i] = x[i] / y[i]; }

/* N_VScale_Parallel :568-592 */
void orc_scale(double c, const double* x, double* z, int64_t n)
{
  int64_t i;
  if (z == x) { for (i = 0; i < n; i++) z[i] *= c; return; } /* VScaleBy :1937 */
  if (c == 1.0) { for (i = 0; i < n; i++) z[i] = x[i]; }
  else if (c == -1.0) { for (i = 0; i < n; i++) z[i] = -x[i]; }
  else { for (i = 0; i < n; i++) z[i] = c * x[i]; }
}

void orc_abs(const double* x, double* z, int64_t n) { for (int64_t i = 0; i < n; i++) z[i] = fabs(x[i]); }
void orc_inv(const double* x, double* z, int64_t n) { for (int64_t i = 0; i < n; i++) z[i] = 1.0 / x[i]; }
void orc_addconst(const double* x, double b, double* z, int64_t n) { for (int64_t i = 0; i < n; i++) z[i] = x[i] + b; }

double orc_dot(const double* x, const double* y, int64_t n)
{
  double sum = 0.0;
  for (int64_t i = 0; i < n; i++) sum += x[i] * y[i];
  return sum;
}

double orc_maxnorm(const double* x, int64_t n)
{
  double max = 0.0;
  for (int64_t i = 0; i < n; i++)
    if (fabs(x[i]) > max) max = fabs(x[i]);
  return max;
}

double orc_wsqrsum(const double* x, const double* w, int64_t n)
{
  double sum = 0.0;
  for (int64_t i = 0; i < n; i++)
  {
    double prodi = x[i] * w[i];
    sum += prodi * prodi;
  }
  return sum;
}

double orc_wrmsnorm(const double* x, const double* w, int64_t n, int64_t nglobal)
{
  return sqrt(orc_wsqrsum(x, w, n) / (double)nglobal);
}

double orc_min(const double* x, int64_t n)
{
  double m = 1.7976931348623157e308; /* SUN_BIG_REAL */
  if (n > 0)
  {
    m = x[0];
    for (int64_t i = 1; i < n; i++)
      if (x[i] < m) m = x[i];
  }
  return m;
}

double orc_l1norm(const double* x, int64_t n)
{
  double sum = 0.0;
  for (int64_t i = 0; i < n; i++) sum += fabs(x[i]);
  return sum;
}

/* sundials_nvector.c:557-565: nvscale(c[0],X[0],z) then nvlinearsum(c[i],X[i],1,z,z) */
void orc_linear_combination(int nvec, const double* c, const double* const* X, double* z, int64_t n)
{
  orc_scale(c[0], X[0], z, n);
  for (int i = 1; i < nvec; i++) orc_linear_sum(c[i], X[i], 1.0, z, z, n);
}

/* arkode.c:2932-2944 */
void orc_ewt_ss(const double* y, double rtol, double atol, double* tmp, double* ewt, int64_t n)
{
  orc_abs(y, tmp, n);
  orc_scale(rtol, tmp, tmp, n);
  orc_addconst(tmp, atol, tmp, n);
  orc_inv(tmp, ewt, n);
}

/* ============================ diffusion_2D problem ============================ */

double orc_coeff_x(double x, const orc_grid* g)
{
  if (g->inhomogeneous) return (g->kx * (1.0 + 0.99 * sin(x)));
  else return (g->kx);
}

double orc_coeff_y(double y, const orc_grid* g)
{
  if (g->inhomogeneous) return (g->ky * (1.0 + 0.99 * sin(y)));
  else return (g->ky);
}

/* diffusion.cpp:36-46 */
void orc_coeff_tables(const orc_grid* g, double* cxw, double* cxe, double* cys, double* cyn)
{
  const double dx = g->dx, dy = g->dy;
  for (int64_t j = 0; j < g->ny_loc; j++)
  {
    const double ylo = g->yl + (g->js + j - 0.5) * dy;
    const double yhi = g->yl + (g->js + j + 0.5) * dy;
    cys[j]           = orc_coeff_y(ylo, g) / (dy * dy);
    cyn[j]           = orc_coeff_y(yhi, g) / (dy * dy);
  }
  for (int64_t i = 0; i < g->nx_loc; i++)
  {
    const double xlo = g->xl + (g->is + i - 0.5) * dx;
    const double xhi = g->xl + (g->is + i + 0.5) * dx;
    cxw[i]           = orc_coeff_x(xlo, g) / (dx * dx);
    cxe[i]           = orc_coeff_x(xhi, g) / (dx * dx);
  }
}

/* One cell of diffusion.cpp:48-53 (and its face variants :79-203, which differ only
   in where the neighbour values are read from). */
static double cell(const orc_grid* g, int64_t i, int64_t j, double uc, double uw, double ue,
                   double us, double un)
{
  const double dx = g->dx, dy = g->dy;
  const double ylo  = g->yl + (g->js + j - 0.5) * dy;
  const double yhi  = g->yl + (g->js + j + 0.5) * dy;
  const double Dy_s = orc_coeff_y(ylo, g) / (dy * dy);
  const double Dy_n = orc_coeff_y(yhi, g) / (dy * dy);
  const double xlo  = g->xl + (g->is + i - 0.5) * dx;
  const double xhi  = g->xl + (g->is + i + 0.5) * dx;
  const double Dx_w = orc_coeff_x(xlo, g) / (dx * dx);
  const double Dx_e = orc_coeff_x(xhi, g) / (dx * dx);
  double f          = 0.0; /* N_VConst(ZERO, f) diffusion.cpp:31 */
  f += -((Dx_w + Dx_e) + (Dy_s + Dy_n)) * uc + Dx_w * uw + Dx_e * ue + Dy_s * us + Dy_n * un;
  return f;
}

void orc_laplacian(const orc_grid* g, const double* u, double* f, const double* W,
                   const double* E, const double* S, const double* N)
{
  const int64_t nx = g->nx_loc, ny = g->ny_loc;
  for (int64_t j = 0; j < ny; j++)
  {
    for (int64_t i = 0; i < nx; i++)
    {
      /* interior cells read the field (diffusion.cpp:34-55); face/corner cells read the
         receive buffers (:68-205); with one periodic rank in a direction the buffer
         holds this rank's opposite edge (diffusion_2D.cpp:421-503, buffers.cpp:28-42) */
      const double uw = (i > 0) ? u[IDX(i - 1, j, nx)] : (W ? W[j] : u[IDX(nx - 1, j, nx)]);
      const double ue = (i < nx - 1) ? u[IDX(i + 1, j, nx)] : (E ? E[j] : u[IDX(0, j, nx)]);
      const double us = (j > 0) ? u[IDX(i, j - 1, nx)] : (S ? S[i] : u[IDX(i, ny - 1, nx)]);
      const double un = (j < ny - 1) ? u[IDX(i, j + 1, nx)] : (N ? N[i] : u[IDX(i, 0, nx)]);
      f[IDX(i, j, nx)] = cell(g, i, j, u[IDX(i, j, nx)], uw, ue, us, un);
    }
  }
}

void orc_pack(const orc_grid* g, const double* u, double* Ws, double* Es, double* Ss, double* Ns)
{
  const int64_t nx = g->nx_loc, ny = g->ny_loc;
  for (int64_t i = 0; i < ny; i++) Ws[i] = u[IDX(0, i, nx)];
  for (int64_t i = 0; i < ny; i++) Es[i] = u[IDX(nx - 1, i, nx)];
  for (int64_t i = 0; i < nx; i++) Ss[i] = u[IDX(i, 0, nx)];
  for (int64_t i = 0; i < nx; i++) Ns[i] = u[IDX(i, ny - 1, nx)];
}

/* initial.cpp:20-48 (all HaveNbr* are true: periodic, diffusion_2D.cpp:324-327) */
void orc_initial(const orc_grid* g, double* u)
{
  for (int64_t j = 0; j < g->ny_loc; j++)
    for (int64_t i = 0; i < g->nx_loc; i++)
    {
      const double x = g->xl + (g->is + i) * g->dx;
      const double y = g->yl + (g->js + j) * g->dy;
      u[IDX(i, j, g->nx_loc)] = (1.0 + 0.3 * sin(2.0 * x)) / sqrt(5.5 * M_PI) * exp(-(y * y) / 5.5);
    }
}

/* preconditioner_jacobi.cpp:9-46 -- note the coordinates: (js+j)*dy, no yl, no half cell */
void orc_jacobi_setup(const orc_grid* g, double gamma, double* diag)
{
  for (int64_t j = 0; j < g->ny_loc; j++)
  {
    const double Dy_s = orc_coeff_y((g->js + j) * g->dy, g) / (g->dy * g->dy);
    const double Dy_n = orc_coeff_y((g->js + j + 1) * g->dy, g) / (g->dy * g->dy);
    for (int64_t i = 0; i < g->nx_loc; i++)
    {
      const double Dx_w = orc_coeff_x((g->is + i) * g->dx, g) / (g->dx * g->dx);
      const double Dx_e = orc_coeff_x((g->is + i + 1) * g->dx, g) / (g->dx * g->dx);
      const double d    = -((Dx_w + Dx_e) + (Dy_s + Dy_n));
      diag[IDX(i, j, g->nx_loc)] = 1.0 / (1.0 - gamma * d);
    }
  }
}

double orc_dom_eig(const orc_grid* g)
{
  const double a = g->kx / g->dx / g->dx, b = g->ky / g->dy / g->dy;
  return -8.0 * ((a < b) ? b : a); /* std::max(a,b) */
}

/* MPI_Dims_create for 2 dimensions: balanced factors, non-increasing order */
void orc_dims_create(int np, int dims[2])
{
  int b = 1;
  for (int f = 1; f * f <= np; f++)
    if (np % f == 0) b = f;
  dims[0] = np / b;
  dims[1] = b;
}

/* diffusion_2D.cpp:286-317 */
void orc_decompose(int64_t n, int nproc, int coord, int64_t* start, int64_t* count)
{
  int64_t q = n / nproc, r = n % nproc;
  int64_t s = q * coord + (coord < r ? coord : r);
  int64_t e = s + q - 1 + (coord < r ? 1 : 0);
  *start    = s;
  *count    = e - s + 1;
}

/* =========================== LSRKStep recurrences =========================== */

int orc_stages_rkc(double h, double sr)
{
  double ss = ceil(sqrt(1.54 * fabs(h) * sr));
  if (ss < 2.0) ss = 2.0;
  return (int)ss;
}

int orc_stages_rkl(double h, double sr)
{
  double ss = ceil((sqrt(9.0 + 8.0 * fabs(h) * sr) - 1.0) / 2.0);
  if (ss < 2.0) ss = 2.0;
  return (int)ss;
}

/* embedding + WRMS common to RKC (:766-798) and RKL (:1057-1085) */
static int sts_finish(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double* dsm)
{
  if (f(tn + h, w->ycur, w->tempv2, user)) return -1;
  w->nfe++;
  *dsm = 0.0;
  if (!w->fixedstep)
  {
    double c[4]          = {0.8, -0.8, 0.4 * h, 0.4 * h};
    const double* X[4]   = {w->yn, w->ycur, w->fn, w->tempv2};
    orc_linear_combination(4, c, X, w->tempv1, w->n);
    *dsm = orc_wrmsnorm(w->tempv1, w->ewt, w->n, w->nglobal);
  }
  return 0;
}

int orc_step_rkc(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double sr, double* dsm)
{
  const int s = orc_stages_rkc(h, sr);
  double w0, w1, temp1, temp2, arg, bjm1, bjm2, mus, thjm1, thjm2, zjm1, zjm2, dzjm1, dzjm2,
    d2zjm1, d2zjm2, zj, dzj, d2zj, bj, ajm1, mu, nu, thj;
  double *tempv1 = w->tempv1, *tempv2 = w->tempv2;

  w0    = (1.0 + 2.0 / (13.0 * ((double)s * (double)s)));                 /* :629 */
  temp1 = w0 * w0 - 1.0;
  temp2 = sqrt(temp1);
  arg   = s * log(w0 + temp2);
  w1    = sinh(arg) * temp1 / (cosh(arg) * s * temp2 - w0 * sinh(arg));  /* :635-636 */
  bjm1  = 1.0 / ((2.0 * w0) * (2.0 * w0));                                /* :638 */
  bjm2  = bjm1;

  orc_scale(1.0, w->yn, tempv1, w->n);                                    /* :642 */
  mus = w1 * bjm1;
  orc_linear_sum(1.0, w->yn, h * mus, w->fn, tempv2, w->n);               /* :649 */

  thjm2 = 0.0; thjm1 = mus; zjm1 = w0; zjm2 = 1.0; dzjm1 = 1.0; dzjm2 = 0.0; d2zjm1 = 0.0; d2zjm2 = 0.0;

  for (int j = 2; j <= s; j++)                                             /* :674 */
  {
    zj   = 2.0 * w0 * zjm1 - zjm2;
    dzj  = 2.0 * w0 * dzjm1 - dzjm2 + 2.0 * zjm1;
    d2zj = 2.0 * w0 * d2zjm1 - d2zjm2 + 4.0 * dzjm1;
    bj   = d2zj / (dzj * dzj);
    ajm1 = 1.0 - zjm1 * bjm1;
    mu   = 2.0 * w0 * bj / bjm1;
    nu   = -bj / bjm2;
    mus  = mu * w1 / w0;

    if (f(tn + h * thjm1, tempv2, w->ycur, user)) return -1;             /* :686 */
    w->nfe++;
    thj = mu * thjm1 + nu * thjm2 + mus * (1.0 - ajm1);

    double c[5]        = {mus * h, nu, 1.0 - mu - nu, mu, -mus * ajm1 * h}; /* :706-715 */
    const double* X[5] = {w->ycur, tempv1, w->yn, tempv2, w->fn};
    orc_linear_combination(5, c, X, w->ycur, w->n);                        /* :717 */

    if (j < s)
    {
      double* t = tempv1; tempv1 = tempv2; tempv2 = t;                     /* :742-744 */
      orc_scale(1.0, w->ycur, tempv2, w->n);                               /* :746 */
      thjm2 = thjm1; thjm1 = thj; bjm2 = bjm1; bjm1 = bj; zjm2 = zjm1; zjm1 = zj;
      dzjm2 = dzjm1; dzjm1 = dzj; d2zjm2 = d2zjm1; d2zjm1 = d2zj;
    }
  }
  w->tempv1 = tempv1; w->tempv2 = tempv2; /* ARKODE's pointers stay swapped */
  if (sts_finish(w, f, user, tn, h, dsm)) return -1;
  return s;
}

int orc_step_rkl(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double sr, double* dsm)
{
  const int s = orc_stages_rkl(h, sr);
  double w1, bjm1, bjm2, mus, bj, ajm1, cjm1, temj, cj, mu, nu;
  double *tempv1 = w->tempv1, *tempv2 = w->tempv2;

  w1   = 4.0 / ((s + 2.0) * (s - 1.0));                                   /* :944 */
  bjm2 = 1.0 / 3.0;
  bjm1 = bjm2;
  orc_scale(1.0, w->yn, tempv1, w->n);                                    /* :950 */
  mus  = w1 * bjm1;
  cjm1 = mus;
  orc_linear_sum(1.0, w->yn, h * mus, w->fn, tempv2, w->n);               /* :958 */

  for (int j = 2; j <= s; j++)                                             /* :974 */
  {
    temj = (j + 2.0) * (j - 1.0);
    bj   = temj / (2.0 * j * (j + 1.0));
    ajm1 = 1.0 - bjm1;
    mu   = (2.0 * j - 1.0) / j * (bj / bjm1);
    nu   = -(j - 1.0) / j * (bj / bjm2);
    mus  = w1 * mu;
    cj   = temj * w1 / 4.0;

    if (f(tn + h * cjm1, tempv2, w->ycur, user)) return -1;              /* :985 */
    w->nfe++;
    double c[5]        = {mus * h, nu, 1.0 - mu - nu, mu, -mus * ajm1 * h};
    const double* X[5] = {w->ycur, tempv1, w->yn, tempv2, w->fn};
    orc_linear_combination(5, c, X, w->ycur, w->n);                        /* :1020 */
    if (j < s)
    {
      double* t = tempv1; tempv1 = tempv2; tempv2 = t;
      orc_scale(1.0, w->ycur, tempv2, w->n);
      cjm1 = cj; bjm2 = bjm1; bjm1 = bj;
    }
  }
  w->tempv1 = tempv1; w->tempv2 = tempv2;
  if (sts_finish(w, f, user, tn, h, dsm)) return -1;
  return s;
}

int orc_step_ssps2(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, int stages, double* dsm)
{
  const double rs = (double)stages, sm1inv = 1.0 / (rs - 1.0);
  double bt1, bt2, bt3;
  if (stages == 2) { bt1 = 0.694021459207626; bt2 = 0.0; bt3 = 1.0 - bt1; }
  else { bt1 = (rs + 1.0) / (rs * rs); bt2 = 1.0 / rs; bt3 = (rs - 1.0) / (rs * rs); }
  *dsm = 0.0;
  orc_linear_sum(1.0, w->yn, sm1inv * h, w->fn, w->ycur, w->n);                       /* :1180 */
  if (!w->fixedstep) orc_linear_sum(1.0, w->yn, bt1 * h, w->fn, w->tempv1, w->n);
  for (int j = 2; j < stages; j++)
  {
    if (f(tn + ((double)j - 1.0) * sm1inv * h, w->ycur, w->tempv2, user)) return -1;
    w->nfe++;
    orc_linear_sum(1.0, w->ycur, sm1inv * h, w->tempv2, w->ycur, w->n);               /* :1213-1231 */
    if (!w->fixedstep) orc_linear_sum(1.0, w->tempv1, bt2 * h, w->tempv2, w->tempv1, w->n);
  }
  if (f(tn + h, w->ycur, w->tempv2, user)) return -1;
  w->nfe++;
  double c[3]        = {1.0 / (sm1inv * rs), 1.0 / rs, h / rs};
  const double* X[3] = {w->ycur, w->yn, w->tempv2};
  orc_linear_combination(3, c, X, w->ycur, w->n);
  if (!w->fixedstep)
  {
    orc_linear_sum(1.0, w->tempv1, bt3 * h, w->tempv2, w->tempv1, w->n);
    orc_linear_sum(1.0, w->ycur, -1.0, w->tempv1, w->tempv1, w->n);
    *dsm = orc_wrmsnorm(w->tempv1, w->ewt, w->n, w->nglobal);
  }
  return stages;
}

int orc_step_ssps3(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, int stages, double* dsm)
{
  const double rs = (double)stages, rn = sqrt(rs), rat = 1.0 / (rs - rn);
  const int in = (int)round(rn);
  *dsm = 0.0;
  orc_linear_sum(1.0, w->yn, h * rat, w->fn, w->ycur, w->n);                          /* :1372 */
  if (!w->fixedstep) orc_linear_sum(1.0, w->yn, h / rs, w->fn, w->tempv1, w->n);
  for (int j = 2; j <= ((in - 1) * (in - 2) / 2); j++)
  {
    if (f(tn + ((double)j - 1.0) * rat * h, w->ycur, w->tempv3, user)) return -1;
    w->nfe++;
    orc_linear_sum(1.0, w->ycur, h * rat, w->tempv3, w->ycur, w->n);
    if (!w->fixedstep) orc_linear_sum(1.0, w->tempv1, h / rs, w->tempv3, w->tempv1, w->n);
  }
  orc_scale(1.0, w->ycur, w->tempv2, w->n);
  for (int j = ((in - 1) * (in - 2) / 2 + 1); j <= (in * (in + 1) / 2 - 1); j++)
  {
    if (f(tn + ((double)j - 1.0) * rat * h, w->ycur, w->tempv3, user)) return -1;
    w->nfe++;
    orc_linear_sum(1.0, w->ycur, h * rat, w->tempv3, w->ycur, w->n);
    if (!w->fixedstep) orc_linear_sum(1.0, w->tempv1, h / rs, w->tempv3, w->tempv1, w->n);
  }
  if (f(tn + rat * (rn * (rn + 1.0) / 2.0 - 1.0) * h, w->ycur, w->tempv3, user)) return -1;
  w->nfe++;
  double c[3]        = {(rn - 1.0) / (2.0 * rn - 1.0), rn / (2.0 * rn - 1.0),
                        (rn - 1.0) * rat * h / (2.0 * rn - 1.0)};
  const double* X[3] = {w->ycur, w->tempv2, w->tempv3};
  orc_linear_combination(3, c, X, w->ycur, w->n);
  if (!w->fixedstep) orc_linear_sum(1.0, w->tempv1, h / rs, w->tempv3, w->tempv1, w->n);
  for (int j = (in * (in + 1) / 2 + 1); j <= stages; j++)
  {
    if (f(tn + ((double)j - rn - 1.0) * rat * h, w->ycur, w->tempv3, user)) return -1;
    w->nfe++;
    orc_linear_sum(1.0, w->ycur, h * rat, w->tempv3, w->ycur, w->n);
    if (!w->fixedstep) orc_linear_sum(1.0, w->tempv1, h / rs, w->tempv3, w->tempv1, w->n);
  }
  if (!w->fixedstep)
  {
    orc_linear_sum(1.0, w->ycur, -1.0, w->tempv1, w->tempv1, w->n);
    *dsm = orc_wrmsnorm(w->tempv1, w->ewt, w->n, w->nglobal);
  }
  return stages;
}

/* lsrkStep_TakeStepSSP43, arkode_lsrkstep.c:1610-1796: the 4-stage, 3rd-order SSP method with its 2nd-order
 * embedding accumulated in tempv1 (h/4 times every stage value). */
int orc_step_ssp43(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double* dsm)
{
  const double rs = 4.0, p5 = 0.5;
  *dsm = 0.0;
  orc_linear_sum(1.0, w->yn, h * p5, w->fn, w->ycur, w->n);                          /* :1655 */
  if (!w->fixedstep) orc_linear_sum(1.0, w->yn, h / rs, w->fn, w->tempv1, w->n);
  if (f(tn + h * p5, w->ycur, w->tempv3, user)) return -1;                           /* :1676 */
  w->nfe++;
  orc_linear_sum(1.0, w->ycur, h * p5, w->tempv3, w->ycur, w->n);
  if (!w->fixedstep) orc_linear_sum(1.0, w->tempv1, h / rs, w->tempv3, w->tempv1, w->n);
  if (f(tn + h, w->ycur, w->tempv3, user)) return -1;                                /* :1711 */
  w->nfe++;
  double c[3]        = {1.0 / 3.0, 2.0 / 3.0, 1.0 / 6.0 * h};
  const double* X[3] = {w->ycur, w->yn, w->tempv3};
  orc_linear_combination(3, c, X, w->ycur, w->n);                                    /* :1726-1732 */
  if (!w->fixedstep) orc_linear_sum(1.0, w->tempv1, h / rs, w->tempv3, w->tempv1, w->n);
  if (f(tn + h * p5, w->ycur, w->tempv3, user)) return -1;                           /* :1758 */
  w->nfe++;
  orc_linear_sum(1.0, w->ycur, h * p5, w->tempv3, w->ycur, w->n);
  if (!w->fixedstep)
  {
    orc_linear_sum(1.0, w->tempv1, h / rs, w->tempv3, w->tempv1, w->n);
    orc_linear_sum(1.0, w->ycur, -1.0, w->tempv1, w->tempv1, w->n);
    *dsm = orc_wrmsnorm(w->tempv1, w->ewt, w->n, w->nglobal);
  }
  return 4;
}

/* lsrkStep_TakeStepSSP104, arkode_lsrkstep.c:1816-2023: the 10-stage, 4th-order SSP method (two sweeps of
 * five sixth-steps joined by the 1/25, 9/25 | 15, -5 recombination) and its embedding in tempv1. */
int orc_step_ssp104(orc_step_ws* w, orc_rhs_fn f, void* user, double tn, double h, double* dsm)
{
  const double onesixth = 1.0 / 6.0, onefifth = 1.0 / 5.0;
  *dsm = 0.0;
  orc_scale(1.0, w->yn, w->tempv2, w->n);                                            /* :1858 */
  orc_linear_sum(1.0, w->yn, onesixth * h, w->fn, w->ycur, w->n);
  if (!w->fixedstep) orc_linear_sum(1.0, w->yn, onefifth * h, w->fn, w->tempv1, w->n);
  for (int j = 2; j <= 5; j++)
  {                                                                                  /* :1871-1911 */
    if (f(tn + ((double)j - 1.0) * onesixth * h, w->ycur, w->tempv3, user)) return -1;
    w->nfe++;
    orc_linear_sum(1.0, w->ycur, onesixth * h, w->tempv3, w->ycur, w->n);
    if (j == 4 && !w->fixedstep) orc_linear_sum(1.0, w->tempv1, 0.3 * h, w->tempv3, w->tempv1, w->n);
  }
  orc_linear_sum(1.0 / 25.0, w->tempv2, 9.0 / 25.0, w->ycur, w->tempv2, w->n);       /* :1916-1920 */
  orc_linear_sum(15.0, w->tempv2, -5.0, w->ycur, w->ycur, w->n);
  for (int j = 6; j <= 9; j++)
  {                                                                                  /* :1922-1968 */
    if (f(tn + ((double)j - 4.0) * onesixth * h, w->ycur, w->tempv3, user)) return -1;
    w->nfe++;
    orc_linear_sum(1.0, w->ycur, onesixth * h, w->tempv3, w->ycur, w->n);
    if (j == 7 && !w->fixedstep) orc_linear_sum(1.0, w->tempv1, onefifth * h, w->tempv3, w->tempv1, w->n);
    if (j == 9 && !w->fixedstep) orc_linear_sum(1.0, w->tempv1, 0.3 * h, w->tempv3, w->tempv1, w->n);
  }
  if (f(tn + h, w->ycur, w->tempv3, user)) return -1;                                /* :1984 */
  w->nfe++;
  double c[3]        = {0.6, 1.0, 0.1 * h};
  const double* X[3] = {w->ycur, w->tempv2, w->tempv3};
  orc_linear_combination(3, c, X, w->ycur, w->n);                                    /* :1996-2002 */
  if (!w->fixedstep)
  {
    orc_linear_sum(1.0, w->ycur, -1.0, w->tempv1, w->tempv1, w->n);
    *dsm = orc_wrmsnorm(w->tempv1, w->ewt, w->n, w->nglobal);
  }
  return 10;
}

/* ------------------------------ fixed-step driver ------------------------------ */
static int lap_rhs(double t, const double* y, double* f, void* user)
{
  (void)t;
  orc_laplacian((const orc_grid*)user, y, f, NULL, NULL, NULL, NULL);
  return 0;
}

long orc_diffusion_fixed_run(const orc_grid* g, int method, double h, int nsteps, double* u)
{
  const int64_t n = g->nx_loc * g->ny_loc;
  orc_step_ws w;
  memset(&w, 0, sizeof(w));
  w.n = n; w.nglobal = n; w.fixedstep = 1;
  w.yn     = (double*)malloc(sizeof(double) * n);
  w.fn     = (double*)malloc(sizeof(double) * n);
  w.tempv1 = (double*)malloc(sizeof(double) * n);
  w.tempv2 = (double*)malloc(sizeof(double) * n);
  w.ycur   = u; /* ARKODE uses the caller's vector as ycur (arkode.c:690) */
  orc_initial(g, w.yn);
  /* lambda *= dom_eig_safety (1.01); rho = sqrt(lR^2 + lI^2)  arkode_lsrkstep.c:2340-2343 */
  double lam = orc_dom_eig(g) * 1.01;
  double sr  = sqrt(lam * lam + 0.0 * 0.0);
  lap_rhs(0.0, w.yn, w.fn, (void*)g);
  w.nfe     = 1;
  double tn = 0.0, dsm;
  for (int k = 0; k < nsteps; k++)
  {
    if (method == 0) orc_step_rkc(&w, lap_rhs, (void*)g, tn, h, sr, &dsm);
    else orc_step_rkl(&w, lap_rhs, (void*)g, tn, h, sr, &dsm);
    orc_scale(1.0, w.tempv2, w.fn, n); /* lsrkStep_DomEigUpdateLogic :2242 */
    orc_scale(1.0, w.ycur, w.yn, n);   /* arkCompleteStep, arkode.c:2737 */
    tn += h;
  }
  if (nsteps == 0) orc_scale(1.0, w.yn, u, n);
  long nfe = w.nfe;
  free(w.yn); free(w.fn); free(w.tempv1); free(w.tempv2);
  return nfe;
}

/* ================================ power iteration ================================ */
int orc_power_iteration(orc_atimes_fn A, void* user, double* V, double* q, int64_t n,
                        int num_warmups, int max_iters, double rel_tol, double* lambdaR, int* iters)
{
  double newl = 0.0, oldl = 0.0, normq;
  int it = 0;
  for (int i = 0; i < num_warmups; i++)
  {
    if (A(user, V, q)) return -1;
    it++;
    normq = sqrt(orc_dot(q, q, n));
    orc_scale(1.0 / normq, q, V, n);
  }
  for (int k = 0; k < max_iters; k++)
  {
    if (A(user, V, q)) return -1;
    it++;
    newl       = orc_dot(V, q, n);
    double res = fabs(newl - oldl) / fabs(newl);
    if (res < rel_tol) break;
    normq = sqrt(orc_dot(q, q, n));
    orc_scale(1.0 / normq, q, V, n);
    oldl = newl;
  }
  *lambdaR = newl;
  *iters   = it;
  return 0;
}

/* ==================================== adr 2-D ==================================== */
#define UIDX(i, j, n) (2 * ((i) + (j) * (n)))
#define VIDX(i, j, n) (2 * ((i) + (j) * (n)) + 1)

void orc_adr_advection(const orc_adr* p, const double* y, double* f)
{
  const double cux = 1.0 * p->cux / (2.0 * p->dx), cuy = 1.0 * p->cuy / (2.0 * p->dy);
  const double cvx = 1.0 * p->cvx / (2.0 * p->dx), cvy = 1.0 * p->cvy / (2.0 * p->dy);
  const int64_t nx = p->nx, ny = p->ny;
  for (int64_t j = 0; j < ny; j++)
    for (int64_t i = 0; i < nx; i++)
    {
      const double ulx = (i > 0) ? y[UIDX(i - 1, j, nx)] : y[UIDX(nx - 1, j, nx)];
      const double urx = (i < nx - 1) ? y[UIDX(i + 1, j, nx)] : y[UIDX(0, j, nx)];
      const double uby = (j > 0) ? y[UIDX(i, j - 1, nx)] : y[UIDX(i, ny - 1, nx)];
      const double uty = (j < ny - 1) ? y[UIDX(i, j + 1, nx)] : y[UIDX(i, 0, nx)];
      const double vlx = (i > 0) ? y[VIDX(i - 1, j, nx)] : y[VIDX(nx - 1, j, nx)];
      const double vrx = (i < nx - 1) ? y[VIDX(i + 1, j, nx)] : y[VIDX(0, j, nx)];
      const double vby = (j > 0) ? y[VIDX(i, j - 1, nx)] : y[VIDX(i, ny - 1, nx)];
      const double vty = (j < ny - 1) ? y[VIDX(i, j + 1, nx)] : y[VIDX(i, 0, nx)];
      f[UIDX(i, j, nx)] = cux * (urx - ulx) + cuy * (uty - uby);
      f[VIDX(i, j, nx)] = cvx * (vrx - vlx) + cvy * (vty - vby);
    }
}

void orc_adr_diffusion(const orc_adr* p, const double* y, double* f)
{
  const double d = p->d, dxinv2 = 1.0 / (p->dx * p->dx), dyinv2 = 1.0 / (p->dy * p->dy);
  const int64_t nx = p->nx, ny = p->ny;
  for (int64_t j = 0; j < ny; j++)
    for (int64_t i = 0; i < nx; i++)
    {
      const double uc  = y[UIDX(i, j, nx)];
      const double ulx = (i > 0) ? y[UIDX(i - 1, j, nx)] : y[UIDX(nx - 1, j, nx)];
      const double urx = (i < nx - 1) ? y[UIDX(i + 1, j, nx)] : y[UIDX(0, j, nx)];
      const double uby = (j > 0) ? y[UIDX(i, j - 1, nx)] : y[UIDX(i, ny - 1, nx)];
      const double uty = (j < ny - 1) ? y[UIDX(i, j + 1, nx)] : y[UIDX(i, 0, nx)];
      const double vc  = y[VIDX(i, j, nx)];
      const double vlx = (i > 0) ? y[VIDX(i - 1, j, nx)] : y[VIDX(nx - 1, j, nx)];
      const double vrx = (i < nx - 1) ? y[VIDX(i + 1, j, nx)] : y[VIDX(0, j, nx)];
      const double vby = (j > 0) ? y[VIDX(i, j - 1, nx)] : y[VIDX(i, ny - 1, nx)];
      const double vty = (j < ny - 1) ? y[VIDX(i, j + 1, nx)] : y[VIDX(i, 0, nx)];
      f[UIDX(i, j, nx)] = d * dxinv2 * (ulx + urx - 2.0 * uc) + d * dyinv2 * (uby + uty - 2.0 * uc);
      f[VIDX(i, j, nx)] = d * dxinv2 * (vlx + vrx - 2.0 * vc) + d * dyinv2 * (vby + vty - 2.0 * vc);
    }
}

void orc_adr_reaction(const orc_adr* p, const double* y, double* f)
{
  for (int64_t j = 0; j < p->ny; j++)
    for (int64_t i = 0; i < p->nx; i++)
    {
      const double u = y[UIDX(i, j, p->nx)], v = y[VIDX(i, j, p->nx)];
      f[UIDX(i, j, p->nx)] = p->A + u * u * v - (p->B + 1.0) * u;
      f[VIDX(i, j, p->nx)] = p->B * u - u * u * v;
    }
}

void orc_adr_adv_react(const orc_adr* p, const double* y, double* tmp, double* f)
{
  orc_adr_advection(p, y, f);
  orc_adr_reaction(p, y, tmp);
  orc_linear_sum(1.0, f, 1.0, tmp, f, 2 * p->nx * p->ny);
}

double orc_adr_domeig(const orc_adr* p)
{
  return -4.0 * p->d / p->dx / p->dx - 4.0 * p->d / p->dy / p->dy;
}

void orc_adr_ic(const orc_adr* p, double xl, double yl, double* yv)
{
  for (int64_t j = 0; j < p->ny; j++)
  {
    const double y = yl + j * p->dy;
    for (int64_t i = 0; i < p->nx; i++)
    {
      const double x = xl + i * p->dx;
      yv[UIDX(i, j, p->nx)] = 22.0 * y * pow((1.0 - y), 1.5);
      yv[VIDX(i, j, p->nx)] = 27.0 * x * pow((1.0 - x), 1.5);
    }
  }
}
