// adr_b200.cpp -- the adr 2-D (Brusselator advection-diffusion-reaction) problem layer on
// the B200 vector.
//
// Host C++ mirror of /root/reference/adr/advection_diffusion_reaction_2d.{cpp,hpp}: UserData /
// UserOptions, ReadInputs, the RHS + eigenvalue callbacks, SetupERK / SetupARK / SetupExtSTS /
// SetupStrang and the main() evolve loop -- with the data path on the GPU.  The callbacks never
// loop over the grid: each marks its output as a deferred operator value, and the vector fuses
// the operator (b200_adr_lincomb, one sm_100a kernel) into the linear combination that consumes
// it.  ARKODE (SplittingStep, MRIStep, LSRKStep, ARKStep, ERKStep, SPGMR, Newton) is linked
// unchanged.
//
// --implicit-reaction is available for the two STS integrators (--integrator 2 ExtSTS, 3 Strang): the reference's
// SUNBandMatrix + SUNLinSol_Band pair is replaced by the block-diagonal pair of b200_blockdiag.h on device data
// (J_reaction -> b200_adr_J_reaction).  With --integrator 1 it needs the reference's BBD preconditioner on host arrays
// and stays rejected, as does --calc_error (4th-order ARK reference run); SURVEY.md section 2.1.

#include <arkode/arkode_arkstep.h>
#include <arkode/arkode_erkstep.h>
#include <arkode/arkode_lsrkstep.h>
#include <arkode/arkode_mristep.h>
#include <arkode/arkode_splittingstep.h>
#include <sundials/sundials_core.h>
#include <sunlinsol/sunlinsol_spgmr.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200_adr2d.h"
#include "b200_blockdiag.h"
#include "b200_callbacks.h"
#include "b200_sts.h"
#include "nvector_b200.h"

namespace {

double wall_seconds()
{
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct AdrData;

struct ModeOp
{ // one deferred operator per composite callback: mode bits 1 advection, 2 diffusion, 4 reaction
  AdrData* ud;
  int mode;
  B200RhsOp op;
};

// Field names follow UserData, adr/advection_diffusion_reaction_2d.hpp:53-98
struct AdrData
{
  bool impl_reaction = false;
  bool advection     = true;
  double cux = -0.5, cuy = 1.0, cvx = 0.4, cvy = 0.7, d = 1.0e-2;
  double A = 1.3, B = 1.0;
  double tf = 1.0;
  double xl = 0.0, xu = 1.0, yl = 0.0, yu = 1.0;
  int64_t nx = 400, ny = 400;
  double dx = 0.0, dy = 0.0;
  int64_t neq = 0;
  MRIStepInnerStepper sts_mem = nullptr;

  b200_ctx* ctx = nullptr;
  ModeOp ops[8];
  long rhs_calls = 0;

  b200_adr_params params() const
  {
    b200_adr_params p;
    p.nx = nx; p.ny = ny; p.dx = dx; p.dy = dy;
    p.cux = cux; p.cuy = cuy; p.cvx = cvx; p.cvy = cvy;
    p.d = d; p.A = A; p.B = B;
    return p;
  }
};

// Field names follow UserOptions, ...2d.hpp:104-172
struct AdrOptions
{
  int integrator = 1, table_id = 0, order = 2, sts_method = 0, extsts_method = 0;
  double rtol = 1.0e-3, atol = 1.0e-11, error_bias = 1.0, fixed_h = 0.0;
  int maxsteps = 100000, predictor = 0, ls_setup_freq = 0, maxl = 0, maxnewt = 10;
  double nlscoef = 0.01, epslin = 0.01;
  bool linear = false, calc_error = false, write_solution = false;
  int output = 1, nout = 1;
  bool no_fusion = false; // B200 extra
  int sts_chain  = 4;     // B200 extra: temporal-blocking depth of the STS diffusion stages (1 = off; 4 measured best
                          // at 2048^2 on the B200: 1.52 ms per Strang step against 2.86 ms with one launch per stage)
};

int adr_fused(void* self, b200_ctx* ctx, const double* y, int nterms, const double* c, const int* src,
              const double* const* v, double* z, double* f_out, const double* wrms_w,
              double* wrms_result, int* wrms_done)
{
  (void)wrms_w; (void)wrms_result;
  ModeOp* m  = static_cast<ModeOp*>(self);
  *wrms_done = 0;
  m->ud->rhs_calls++;
  const b200_adr_params p = m->ud->params();
  return b200_adr_lincomb(ctx, &p, m->mode, y, nterms, c, src, v, z, f_out);
}

// K consecutive STS stages of the diffusion partition in one pass (B200RhsOp::chain; one periodic rank)
int adr_chain_cb(void* self, b200_ctx* ctx, int nstages, const double* x, const double* prev2, const double* yn,
                 const double* fn, const double* coeffs, double* const* z_out, double* const* halos,
                 const int* halo_valid)
{
  (void)halos; (void)halo_valid;
  ModeOp* m = static_cast<ModeOp*>(self);
  m->ud->rhs_calls += nstages;
  const b200_adr_params p = m->ud->params();
  return b200_adr_chain(ctx, &p, nstages, x, prev2, yn, fn, coeffs, z_out);
}

int defer(int mode, N_Vector y, N_Vector f, void* user_data)
{
  AdrData* ud = static_cast<AdrData*>(user_data);
  return N_VSetDeferredRhs_B200(f, &ud->ops[mode].op, y) ? -1 : 0;
}

} // namespace

extern "C" {

int b200_adr_f_advection(sunrealtype, N_Vector y, N_Vector f, void* ud) { return defer(1, y, f, ud); }
int b200_adr_f_diffusion(sunrealtype, N_Vector y, N_Vector f, void* ud) { return defer(2, y, f, ud); }
int b200_adr_f_reaction(sunrealtype, N_Vector y, N_Vector f, void* ud) { return defer(4, y, f, ud); }
int b200_adr_f_adv_diff(sunrealtype, N_Vector y, N_Vector f, void* ud) { return defer(3, y, f, ud); }
int b200_adr_f_adv_react(sunrealtype, N_Vector y, N_Vector f, void* ud) { return defer(5, y, f, ud); }
int b200_adr_f_diff_react(sunrealtype, N_Vector y, N_Vector f, void* ud) { return defer(6, y, f, ud); }
int b200_adr_f_adv_diff_react(sunrealtype, N_Vector y, N_Vector f, void* ud) { return defer(7, y, f, ud); }

// f_diffusion_forcing, ...2d.cpp:1649-1663: the forcing polynomial is added by
// MRIStepInnerStepper_AddForcing through N_VLinearCombination with f as its first operand,
// i.e. it fuses with the deferred diffusion operator into one launch.
int b200_adr_f_diffusion_forcing(sunrealtype t, N_Vector y, N_Vector f, void* user_data)
{
  AdrData* ud = static_cast<AdrData*>(user_data);
  if (defer(2, y, f, user_data)) return -1;
  return MRIStepInnerStepper_AddForcing(ud->sts_mem, t, f) < 0 ? -1 : 0;
}

// J_reaction, ...2d.cpp:1523-1551 (ARKLsJacFn), into the block-diagonal device matrix of b200_blockdiag.h
int b200_adr_J_reaction(sunrealtype, N_Vector y, N_Vector, SUNMatrix J, void* user_data, N_Vector, N_Vector, N_Vector)
{
  AdrData* ud = static_cast<AdrData*>(user_data);
  double* Jd  = SUNMatrix_B200Block2_Data(J);
  if (!Jd || SUNMatrix_B200Block2_Points(J) != ud->nx * ud->ny) return -1;
  const b200_adr_params p = ud->params();
  return b200_adr_jac_reaction(ud->ctx, &p, N_VGetDeviceArrayPointer_B200(y), Jd) ? -1 : 0;
}

// diffusion_domeig, ...2d.cpp:1666-1679
int b200_adr_domeig(sunrealtype, N_Vector, N_Vector, sunrealtype* lambdaR, sunrealtype* lambdaI,
                    void* user_data, N_Vector, N_Vector, N_Vector)
{
  AdrData* ud = static_cast<AdrData*>(user_data);
  *lambdaR    = -4.0 * ud->d / ud->dx / ud->dx - 4.0 * ud->d / ud->dy / ud->dy;
  *lambdaI    = 0.0;
  return 0;
}

} // extern "C"

// ------------------------------------------------------------------- session
struct STSInnerContent
{ // STSInnerStepperContent, ...2d.hpp:178-186
  void* sts_arkode_mem = nullptr;
  void* user_data      = nullptr;
};

struct b200_adr
{
  AdrData ud;
  AdrOptions uo;
  SUNContext sunctx         = nullptr;
  b200_ctx* ctx             = nullptr;
  N_Vector y                = nullptr;
  void* arkode_mem          = nullptr;
  void* lsrkstep_mem        = nullptr; // Strang partition 0 / ExtSTS inner integrator
  void* arkstep_mem         = nullptr; // Strang partition 1
  SUNStepper steppers[2]    = {nullptr, nullptr};
  MRIStepInnerStepper inner = nullptr;
  STSInnerContent* inner_content = nullptr;
  SUNLinearSolver LS        = nullptr;
  SUNMatrix A               = nullptr; // implicit reaction: block-diagonal Newton matrix
  double t = 0.0, evolve_seconds = 0.0;
  B200VecStats vs0{};
  bool settings_held = false; // N_VAcquireSettings_B200 succeeded for this session
  uint64_t launches0 = 0;
};

namespace {

#define CHK(call, name)                                                       \
  do {                                                                        \
    int flag_ = (call);                                                       \
    if (flag_ < 0)                                                            \
    {                                                                         \
      fprintf(stderr, "ERROR: %s returned %d\n", name, flag_);                \
      return -1;                                                              \
    }                                                                         \
  }                                                                           \
  while (0)
#define CHKP(ptr, name)                                                       \
  do {                                                                        \
    if ((ptr) == nullptr)                                                     \
    {                                                                         \
      fprintf(stderr, "ERROR: %s returned NULL\n", name);                     \
      return -1;                                                              \
    }                                                                         \
  }                                                                           \
  while (0)

// ReadInputs, ...2d.hpp:516-604 (one pass; unknown flags are ignored by the reference's
// find_arg scheme, here they are an error so typos do not silently change the problem)
int read_inputs(const std::vector<std::string>& args, AdrData& ud, AdrOptions& uo)
{
  for (size_t k = 0; k < args.size(); k++)
  {
    const std::string& a = args[k];
    auto need = [&](const char* what) -> const std::string* {
      if (k + 1 >= args.size()) { fprintf(stderr, "ERROR: %s needs a value\n", what); return nullptr; }
      return &args[++k];
    };
#define ARG_D(flag, dst) if (a == flag) { auto s = need(flag); if (!s) return -1; dst = std::stod(*s); continue; }
#define ARG_I(flag, dst) if (a == flag) { auto s = need(flag); if (!s) return -1; dst = std::stoi(*s); continue; }
#define ARG_L(flag, dst) if (a == flag) { auto s = need(flag); if (!s) return -1; dst = std::stoll(*s); continue; }
#define ARG_B(flag, dst, val) if (a == flag) { dst = val; continue; }
    ARG_B("--no-advection", ud.advection, false) ARG_B("--implicit-reaction", ud.impl_reaction, true)
    ARG_D("--cux", ud.cux) ARG_D("--cuy", ud.cuy) ARG_D("--cvx", ud.cvx) ARG_D("--cvy", ud.cvy)
    ARG_D("--d", ud.d) ARG_D("--A", ud.A) ARG_D("--B", ud.B) ARG_D("--tf", ud.tf)
    ARG_D("--xl", ud.xl) ARG_D("--xu", ud.xu) ARG_D("--yl", ud.yl) ARG_D("--yu", ud.yu)
    ARG_L("--nx", ud.nx) ARG_L("--ny", ud.ny)
    ARG_I("--integrator", uo.integrator) ARG_I("--table_id", uo.table_id) ARG_I("--order", uo.order)
    ARG_I("--sts_method", uo.sts_method) ARG_I("--extsts_method", uo.extsts_method)
    ARG_D("--rtol", uo.rtol) ARG_D("--atol", uo.atol) ARG_D("--error_bias", uo.error_bias)
    ARG_D("--fixed_h", uo.fixed_h) ARG_I("--predictor", uo.predictor)
    ARG_I("--lssetupfreq", uo.ls_setup_freq) ARG_I("--maxl", uo.maxl) ARG_I("--maxnewt", uo.maxnewt)
    ARG_D("--nlscoef", uo.nlscoef) ARG_D("--epslin", uo.epslin) ARG_I("--maxsteps", uo.maxsteps)
    ARG_B("--linear", uo.linear, true) ARG_B("--calc_error", uo.calc_error, true)
    ARG_B("--write_solution", uo.write_solution, true)
    ARG_I("--output", uo.output) ARG_I("--nout", uo.nout)
    ARG_B("--no-fusion", uo.no_fusion, true) ARG_I("--sts_chain", uo.sts_chain)
    fprintf(stderr, "ERROR: Unknown input: %s\n", a.c_str());
    return -1;
  }
  // ...2d.hpp:566-569
  ud.dx  = (ud.xu - ud.xl) / (ud.nx);
  ud.dy  = (ud.yu - ud.yl) / (ud.ny);
  ud.neq = 2 * ud.nx * ud.ny;
  if (uo.integrator < 0 || uo.integrator > 3) { fprintf(stderr, "ERROR: Invalid integrator option\n"); return -1; }
  if (uo.table_id < 0 || uo.table_id > 5) { fprintf(stderr, "ERROR: Invalid ARK table ID\n"); return -1; }
  if (ud.impl_reaction && uo.integrator != 2 && uo.integrator != 3)
  {
    fprintf(stderr, "ERROR: --implicit-reaction is available with --integrator 2 (ExtSTS) and 3 (Strang) on the B200 path\n");
    return -1;
  }
  if (uo.calc_error || uo.write_solution)
  {
    fprintf(stderr, "ERROR: --calc_error / --write_solution are not available on the B200 path\n");
    return -1;
  }
  return 0;
}

// SetIC, ...2d.cpp:1682-1699: host libm pow, one upload
int set_ic(N_Vector y, const AdrData& ud)
{
  std::vector<double> h((size_t)ud.neq);
  for (int64_t j = 0; j < ud.ny; j++)
  {
    const double yy = ud.yl + j * ud.dy;
    for (int64_t i = 0; i < ud.nx; i++)
    {
      const double xx = ud.xl + i * ud.dx;
      h[(size_t)(2 * (i + j * ud.nx))]     = 22.0 * yy * std::pow(1.0 - yy, 1.5);
      h[(size_t)(2 * (i + j * ud.nx) + 1)] = 27.0 * xx * std::pow(1.0 - xx, 1.5);
    }
  }
  return N_VCopyFromHost_B200(y, h.data());
}

// the ARS(2,2,2) explicit table both SetupStrang (...2d.cpp:1262-1277) and SetupARK use
ARKodeButcherTable ars222_explicit(bool embedding)
{
  ARKodeButcherTable Be = ARKodeButcherTable_Alloc(3, embedding ? SUNTRUE : SUNFALSE);
  const double gamma = (2.0 - std::sqrt(2.0)) / 2.0;
  const double delta = 1.0 - 1.0 / (2.0 * gamma);
  Be->c[1] = gamma; Be->c[2] = 1.0;
  Be->A[1][0] = gamma; Be->A[2][0] = delta; Be->A[2][1] = 1.0 - delta;
  Be->b[0] = delta; Be->b[1] = 1.0 - delta;
  Be->q = 2;
  if (embedding) { Be->d[1] = 3.0 / 5.0; Be->d[2] = 2.0 / 5.0; Be->p = 1; }
  return Be;
}
ARKodeButcherTable ars222_implicit(bool embedding)
{
  ARKodeButcherTable Bi = ARKodeButcherTable_Alloc(3, embedding ? SUNTRUE : SUNFALSE);
  const double gamma = (2.0 - std::sqrt(2.0)) / 2.0;
  Bi->c[1] = gamma; Bi->c[2] = 1.0;
  Bi->A[1][1] = gamma; Bi->A[2][1] = 1.0 - gamma; Bi->A[2][2] = gamma;
  Bi->b[1] = 1.0 - gamma; Bi->b[2] = gamma;
  Bi->q = 2;
  if (embedding) { Bi->d[1] = 3.0 / 5.0; Bi->d[2] = 2.0 / 5.0; Bi->p = 1; }
  return Bi;
}

// SetupERK, ...2d.cpp:308-355
int setup_erk(b200_adr* p)
{
  AdrData& ud = p->ud; AdrOptions& uo = p->uo;
  ARKRhsFn f = ud.advection ? b200_adr_f_adv_diff_react : b200_adr_f_diff_react;
  p->arkode_mem = ERKStepCreate(f, 0.0, p->y, p->sunctx);
  CHKP(p->arkode_mem, "ERKStepCreate");
  void* mem = p->arkode_mem;
  CHK(ARKodeSStolerances(mem, uo.rtol, uo.atol), "ARKodeSStolerances");
  CHK(ARKodeSetUserData(mem, &ud), "ARKodeSetUserData");
  CHK(ARKodeSetOrder(mem, uo.order), "ARKodeSetOrder");
  if (uo.fixed_h > 0.0) CHK(ARKodeSetFixedStep(mem, uo.fixed_h), "ARKodeSetFixedStep");
  CHK(ARKodeSetMaxNumSteps(mem, uo.maxsteps), "ARKodeSetMaxNumSteps");
  CHK(ARKodeSetStopTime(mem, ud.tf), "ARKodeSetStopTime");
  return 0;
}

// SetupARK, ...2d.cpp:358-713 (explicit-reaction branches; tables 0 = default, 1 = ARS(2,2,2))
int setup_ark(b200_adr* p)
{
  AdrData& ud = p->ud; AdrOptions& uo = p->uo;
  ARKRhsFn fe = ud.advection ? b200_adr_f_adv_react : b200_adr_f_reaction;
  ARKRhsFn fi = b200_adr_f_diffusion;
  if (uo.table_id > 1)
  {
    fprintf(stderr, "ERROR: --table_id %d is not available on the B200 path (0 or 1)\n", uo.table_id);
    return -1;
  }
  p->arkode_mem = ARKStepCreate(fe, fi, 0.0, p->y, p->sunctx);
  CHKP(p->arkode_mem, "ARKStepCreate");
  void* mem = p->arkode_mem;
  CHK(ARKodeSStolerances(mem, uo.rtol, uo.atol), "ARKodeSStolerances");
  CHK(ARKodeSetUserData(mem, &ud), "ARKodeSetUserData");
  p->LS = SUNLinSol_SPGMR(p->y, SUN_PREC_NONE, uo.maxl, p->sunctx);
  CHKP(p->LS, "SUNLinSol_SPGMR");
  CHK(ARKodeSetLinearSolver(mem, p->LS, nullptr), "ARKodeSetLinearSolver");
  CHK(ARKodeSetMaxNonlinIters(mem, uo.maxnewt), "ARKodeSetMaxNonlinIters");
  CHK(ARKodeSetNonlinConvCoef(mem, uo.nlscoef), "ARKodeSetNonlinConvCoef");
  CHK(ARKodeSetEpsLin(mem, uo.epslin), "ARKodeSetEpsLin");
  CHK(ARKodeSetDeduceImplicitRhs(mem, SUNTRUE), "ARKodeSetDeduceImplicitRhs");
  CHK(ARKodeSetPredictorMethod(mem, uo.predictor), "ARKodeSetPredictorMethod");
  if (uo.linear) CHK(ARKodeSetLinear(mem, SUNFALSE), "ARKodeSetLinear");
  if (uo.table_id == 1)
  {
    ARKodeButcherTable Be = ars222_explicit(true), Bi = ars222_implicit(true);
    CHK(ARKStepSetTables(mem, 2, 1, Bi, Be), "ARKStepSetTables");
    ARKodeButcherTable_Free(Be);
    ARKodeButcherTable_Free(Bi);
  }
  else CHK(ARKodeSetOrder(mem, uo.order), "ARKodeSetOrder");
  if (uo.fixed_h > 0.0) CHK(ARKodeSetFixedStep(mem, uo.fixed_h), "ARKodeSetFixedStep");
  else CHK(ARKodeSetErrorBias(mem, uo.error_bias), "ARKodeSetErrorBias");
  CHK(ARKodeSetMaxNumSteps(mem, uo.maxsteps), "ARKodeSetMaxNumSteps");
  CHK(ARKodeSetStopTime(mem, ud.tf), "ARKodeSetStopTime");
  return 0;
}

// STSInnerStepper_Evolve / _FullRhs / _Reset, ...2d.cpp:1340-1399
int inner_evolve(MRIStepInnerStepper stepper, sunrealtype t0, sunrealtype tout, N_Vector y)
{
  void* c = nullptr;
  if (MRIStepInnerStepper_GetContent(stepper, &c) < 0) return -1;
  STSInnerContent* content = static_cast<STSInnerContent*>(c);
  if (ARKodeReset(content->sts_arkode_mem, t0, y) < 0) return 1;
  if (ARKodeSetFixedStep(content->sts_arkode_mem, tout - t0) < 0) return 1;
  if (ARKodeSetStopTime(content->sts_arkode_mem, tout) < 0) return 1;
  sunrealtype tret;
  int flag = ARKodeEvolve(content->sts_arkode_mem, tout, y, &tret, ARK_ONE_STEP);
  return flag < 0 ? flag : 0;
}
int inner_fullrhs(MRIStepInnerStepper stepper, sunrealtype t, N_Vector y, N_Vector f, int)
{
  void* c = nullptr;
  if (MRIStepInnerStepper_GetContent(stepper, &c) < 0) return -1;
  STSInnerContent* content = static_cast<STSInnerContent*>(c);
  return b200_adr_f_diffusion(t, y, f, content->user_data) ? -1 : 0;
}
int inner_reset(MRIStepInnerStepper stepper, sunrealtype tR, N_Vector yR)
{
  void* c = nullptr;
  if (MRIStepInnerStepper_GetContent(stepper, &c) < 0) return -1;
  STSInnerContent* content = static_cast<STSInnerContent*>(c);
  return ARKodeReset(content->sts_arkode_mem, tR, yR) < 0 ? 1 : 0;
}

// the solver block both STS set-ups share when the reaction is implicit (...2d.cpp:820-854, :1207-1246): matrix,
// direct solver, Jacobian, set-up frequency, Newton limits, predictor
int attach_implicit_reaction(b200_adr* p, void* mem, bool deduce_fi)
{
  AdrData& ud = p->ud; AdrOptions& uo = p->uo;
  p->A = SUNMatrix_B200Block2(p->ctx, ud.nx * ud.ny, p->sunctx); // SUNBandMatrix(neq, 2, 2)
  CHKP(p->A, "SUNMatrix_B200Block2");
  p->LS = SUNLinSol_B200Block2(p->y, p->A, p->sunctx);           // SUNLinSol_Band
  CHKP(p->LS, "SUNLinSol_B200Block2");
  CHK(ARKodeSetLinearSolver(mem, p->LS, p->A), "ARKodeSetLinearSolver");
  CHK(ARKodeSetJacFn(mem, b200_adr_J_reaction), "ARKodeSetJacFn");
  CHK(ARKodeSetLSetupFrequency(mem, uo.ls_setup_freq), "ARKodeSetLSetupFrequency");
  CHK(ARKodeSetMaxNonlinIters(mem, uo.maxnewt), "ARKodeSetMaxNonlinIters");
  CHK(ARKodeSetNonlinConvCoef(mem, uo.nlscoef), "ARKodeSetNonlinConvCoef");
  if (deduce_fi) CHK(ARKodeSetDeduceImplicitRhs(mem, SUNTRUE), "ARKodeSetDeduceImplicitRhs");
  CHK(ARKodeSetPredictorMethod(mem, uo.predictor), "ARKodeSetPredictorMethod");
  return 0;
}

// the implicit (G) halves of the ExtSTS couplings when the reaction is implicit, ...2d.cpp:861-1000, :1063-1086:
// rows of { stage, column, value }
void set_implicit_coupling(MRIStepCoupling C, int method)
{
  const double one = 1.0, two = 2.0, four = 4.0, eight = 8.0;
  const double sqrt2 = std::sqrt(two);
  if (method == 0)
  { // ARS(2,2,2)
    const double gamma = one - one / std::sqrt(2.0);
    const double G[8][3] = {{1, 0, gamma}, {2, 0, -gamma}, {2, 2, gamma}, {3, 2, one - gamma},
                            {4, 2, -gamma}, {4, 4, gamma}, {5, 2, -0.4}, {5, 4, 0.4}};
    for (auto& g : G) C->G[0][(int)g[0]][(int)g[1]] = g[2];
  }
  else if (method == 1)
  { // Giraldo
    const double G[11][3] = {{1, 0, two - sqrt2},
                             {2, 0, one - one / sqrt2 - (two - sqrt2)},
                             {2, 2, one - one / sqrt2},
                             {3, 0, one / sqrt2 - one},
                             {3, 2, one / sqrt2},
                             {4, 0, one / (two * sqrt2)},
                             {4, 2, one / (two * sqrt2) - one},
                             {4, 4, one - one / sqrt2},
                             {6, 0, (four - sqrt2) / eight - one / (two * sqrt2)},
                             {6, 2, (four - sqrt2) / eight - one / (two * sqrt2)},
                             {6, 4, one / (two * sqrt2) - (one - one / sqrt2)}};
    for (auto& g : G) C->G[0][(int)g[0]][(int)g[1]] = g[2];
  }
}

// SSP SDIRK 2, ...2d.cpp:1063-1086 (implicit only)
MRIStepCoupling extsts_ssp_sdirk2()
{
  MRIStepCoupling C = MRIStepCoupling_Alloc(1, 6, MRISTEP_IMPLICIT);
  const double one = 1.0, two = 2.0, seven = 7.0, twelve = 12.0;
  const double gamma = one - one / std::sqrt(two);
  C->q = 2; C->p = 1;
  C->c[1] = gamma; C->c[2] = gamma; C->c[3] = one - gamma; C->c[4] = one - gamma; C->c[5] = one;
  const double G[11][3] = {{1, 0, gamma}, {2, 0, -gamma}, {2, 2, gamma}, {3, 2, one - two * gamma},
                           {4, 2, -gamma}, {4, 4, gamma}, {5, 2, two * gamma - one / two}, {5, 4, one / two - gamma},
                           {6, 2, two * gamma - seven / twelve}, {6, 4, seven / twelve - gamma}, {0, 0, 0.0}};
  for (auto& g : G) C->G[0][(int)g[0]][(int)g[1]] = g[2];
  return C;
}

// the MRI couplings of SetupExtSTS, ...2d.cpp:856-1098.  kind: MRISTEP_EXPLICIT (explicit reaction), MRISTEP_IMEX
// (explicit advection + implicit reaction) or MRISTEP_IMPLICIT (implicit reaction, no advection): the slow-explicit
// half W is filled unless kind is IMPLICIT, the slow-implicit half G (set_implicit_coupling) unless it is EXPLICIT
MRIStepCoupling extsts_coupling_explicit(int method, MRISTEP_METHOD_TYPE kind);
MRIStepCoupling extsts_coupling(int method, MRISTEP_METHOD_TYPE kind = MRISTEP_EXPLICIT)
{
  MRIStepCoupling C = extsts_coupling_explicit(method, kind);
  if (C && kind != MRISTEP_EXPLICIT && method >= 0) set_implicit_coupling(C, method);
  return C;
}

MRIStepCoupling extsts_coupling_explicit(int method, MRISTEP_METHOD_TYPE kind)
{
  MRIStepCoupling C = nullptr;
  const double one = 1.0, two = 2.0, three = 3.0, four = 4.0, six = 6.0, eight = 8.0;
  if (method == 4) return kind == MRISTEP_IMPLICIT ? extsts_ssp_sdirk2() : nullptr;
  if (method >= 2 && kind != MRISTEP_EXPLICIT) return nullptr; // Ralston, Heun-Euler: explicit only
  if (method == 0)
  { // ARS(2,2,2)
    C = MRIStepCoupling_Alloc(1, 5, kind);
    const double gamma = one - one / std::sqrt(2.0);
    const double delta = one - one / (2.0 * gamma);
    C->q = 2; C->p = 1;
    C->c[1] = gamma; C->c[2] = gamma; C->c[3] = one; C->c[4] = one;
    if (kind == MRISTEP_IMPLICIT) return C;
    C->W[0][1][0] = gamma;
    C->W[0][3][0] = delta - gamma;
    C->W[0][3][2] = one - delta;
    C->W[0][5][0] = -delta;
    C->W[0][5][2] = delta - 0.4;
    C->W[0][5][4] = 0.4;
  }
  else if (method == 1)
  { // Giraldo
    C = MRIStepCoupling_Alloc(1, 6, kind);
    const double sqrt2 = std::sqrt(two);
    C->q = 2; C->p = 1;
    C->c[1] = two - sqrt2; C->c[2] = two - sqrt2; C->c[3] = one; C->c[4] = one; C->c[5] = one;
    if (kind == MRISTEP_IMPLICIT) return C;
    C->W[0][1][0] = two - sqrt2;
    C->W[0][3][0] = (three - two * sqrt2) / six - (two - sqrt2);
    C->W[0][3][2] = (three + two * sqrt2) / six;
    C->W[0][5][0] = one / (two * sqrt2) - (three - two * sqrt2) / six;
    C->W[0][5][2] = one / (two * sqrt2) - (three + two * sqrt2) / six;
    C->W[0][5][4] = one - one / std::sqrt(2.0);
    C->W[0][6][0] = (four - sqrt2) / eight - (three - two * sqrt2) / six;
    C->W[0][6][2] = (four - sqrt2) / eight - (three + two * sqrt2) / six;
    C->W[0][6][4] = one / (two * sqrt2);
  }
  else if (method == 2)
  { // Ralston
    C = MRIStepCoupling_Alloc(1, 3, MRISTEP_EXPLICIT);
    C->q = 2; C->p = 1;
    C->c[1] = two / three; C->c[2] = one;
    C->W[0][1][0] = two / three;
    C->W[0][2][0] = one / four - two / three;
    C->W[0][2][1] = three / four;
    C->W[0][3][0] = 5.0 / 37.0 - two / three;
    C->W[0][3][1] = two / three - three / four;
    C->W[0][3][2] = 22.0 / 111.0;
  }
  else if (method == 3)
  { // Heun-Euler
    C = MRIStepCoupling_Alloc(1, 3, MRISTEP_EXPLICIT);
    C->q = 2; C->p = 1;
    C->c[1] = one; C->c[2] = one;
    C->W[0][1][0] = one;
    C->W[0][2][0] = -one / two;
    C->W[0][2][1] = one / two;
  }
  else if (method < 0) { C = MRIStepCoupling_LoadTable(static_cast<ARKODE_MRITableID>(-method)); }
  return C;
}

// SetupExtSTS, ...2d.cpp:715-1120
int setup_extsts(b200_adr* p)
{
  AdrData& ud = p->ud; AdrOptions& uo = p->uo;
  // ...2d.cpp:724-733: who treats advection / reaction
  ARKRhsFn fi = ud.impl_reaction ? b200_adr_f_reaction : nullptr;
  ARKRhsFn fe = ud.advection ? (ud.impl_reaction ? b200_adr_f_advection : b200_adr_f_adv_react)
                             : (ud.impl_reaction ? nullptr : b200_adr_f_reaction);
  const MRISTEP_METHOD_TYPE kind = !ud.impl_reaction ? MRISTEP_EXPLICIT : (ud.advection ? MRISTEP_IMEX : MRISTEP_IMPLICIT);
  if (uo.extsts_method == 4 && kind != MRISTEP_IMPLICIT)
  {
    fprintf(stderr, "ERROR: --extsts_method 4 (SSP SDIRK 2) is implicit-only: it needs --implicit-reaction --no-advection\n");
    return -1;
  }
  if ((uo.extsts_method == 2 || uo.extsts_method == 3) && ud.impl_reaction)
  {
    fprintf(stderr, "ERROR: --extsts_method %d is explicit-only (no --implicit-reaction)\n", uo.extsts_method);
    return -1;
  }
  void* sts = LSRKStepCreateSTS(b200_adr_f_diffusion_forcing, 0.0, p->y, p->sunctx);
  CHKP(sts, "LSRKStepCreateSTS");
  p->lsrkstep_mem = sts;
  CHK(ARKodeSetUserData(sts, &ud), "ARKodeSetUserData");
  CHK(LSRKStepSetSTSMethod(sts, uo.sts_method == 0 ? ARKODE_LSRK_RKC_2 : ARKODE_LSRK_RKL_2), "LSRKStepSetSTSMethod");
  CHK(LSRKStepSetDomEigFn(sts, b200_adr_domeig), "LSRKStepSetDomEigFn");
  CHK(LSRKStepSetDomEigFrequency(sts, uo.ls_setup_freq), "LSRKStepSetDomEigFrequency");
  CHK(LSRKStepSetMaxNumStages(sts, 10000), "LSRKStepSetMaxNumStages");
  CHK(ARKodeSetInterpolantType(sts, ARK_INTERP_NONE), "ARKodeSetInterpolantType");
  CHK(MRIStepInnerStepper_Create(p->sunctx, &p->inner), "MRIStepInnerStepper_Create");
  p->inner_content                 = new STSInnerContent();
  p->inner_content->sts_arkode_mem = sts;
  p->inner_content->user_data      = &ud;
  CHK(MRIStepInnerStepper_SetContent(p->inner, p->inner_content), "MRIStepInnerStepper_SetContent");
  CHK(MRIStepInnerStepper_SetEvolveFn(p->inner, inner_evolve), "MRIStepInnerStepper_SetEvolveFn");
  CHK(MRIStepInnerStepper_SetFullRhsFn(p->inner, inner_fullrhs), "MRIStepInnerStepper_SetFullRhsFn");
  CHK(MRIStepInnerStepper_SetResetFn(p->inner, inner_reset), "MRIStepInnerStepper_SetResetFn");
  ud.sts_mem = p->inner;

  p->arkode_mem = MRIStepCreate(fe, fi, 0.0, p->y, p->inner, p->sunctx);
  CHKP(p->arkode_mem, "MRIStepCreate");
  void* mem = p->arkode_mem;
  if (uo.fixed_h > 0.0) CHK(ARKodeSetFixedStep(mem, uo.fixed_h), "ARKodeSetFixedStep");
  else CHK(ARKodeSetErrorBias(mem, uo.error_bias), "ARKodeSetErrorBias");
  CHK(ARKodeSStolerances(mem, uo.rtol, uo.atol), "ARKodeSStolerances");
  CHK(ARKodeSetUserData(mem, &ud), "ARKodeSetUserData");
  if (ud.impl_reaction && attach_implicit_reaction(p, mem, false)) return -1; // ...2d.cpp:820-854
  MRIStepCoupling C = extsts_coupling(uo.extsts_method, kind);
  if (!C) { fprintf(stderr, "ERROR: Invalid extsts method %d\n", uo.extsts_method); return -1; }
  CHK(MRIStepSetCoupling(mem, C), "MRIStepSetCoupling");
  MRIStepCoupling_Free(C);
  CHK(ARKodeSetMaxNumSteps(mem, uo.maxsteps), "ARKodeSetMaxNumSteps");
  CHK(ARKodeSetSafetyFactor(mem, 0.8), "ARKodeSetSafetyFactor");
  CHK(ARKodeSetStopTime(mem, ud.tf), "ARKodeSetStopTime");
  return 0;
}

// SetupStrang, ...2d.cpp:1122-1333
int setup_strang(b200_adr* p)
{
  AdrData& ud = p->ud; AdrOptions& uo = p->uo;
  // ...2d.cpp:1131-1140
  ARKRhsFn fi = ud.impl_reaction ? b200_adr_f_reaction : nullptr;
  ARKRhsFn fe = ud.advection ? (ud.impl_reaction ? b200_adr_f_advection : b200_adr_f_adv_react)
                             : (ud.impl_reaction ? nullptr : b200_adr_f_reaction);
  if (!(uo.fixed_h > 0.0))
  {
    fprintf(stderr, "ERROR: Fixed step size must be specified for Strang splitting.\n");
    return -1;
  }
  // LSRKStep partition
  p->lsrkstep_mem = LSRKStepCreateSTS(b200_adr_f_diffusion, 0.0, p->y, p->sunctx);
  CHKP(p->lsrkstep_mem, "LSRKStepCreateSTS");
  void* ls = p->lsrkstep_mem;
  CHK(ARKodeSetUserData(ls, &ud), "ARKodeSetUserData");
  CHK(LSRKStepSetSTSMethod(ls, uo.sts_method == 0 ? ARKODE_LSRK_RKC_2 : ARKODE_LSRK_RKL_2), "LSRKStepSetSTSMethod");
  CHK(LSRKStepSetDomEigFn(ls, b200_adr_domeig), "LSRKStepSetDomEigFn");
  CHK(LSRKStepSetDomEigFrequency(ls, uo.ls_setup_freq), "LSRKStepSetDomEigFrequency");
  CHK(LSRKStepSetMaxNumStages(ls, 10000), "LSRKStepSetMaxNumStages");
  CHK(ARKodeSetFixedStep(ls, uo.fixed_h), "ARKodeSetFixedStep");
  CHK(ARKodeSetMaxNumSteps(ls, uo.maxsteps), "ARKodeSetMaxNumSteps");
  CHK(ARKodeSetInterpolantType(ls, ARK_INTERP_NONE), "ARKodeSetInterpolantType");
  CHK(ARKodeCreateSUNStepper(ls, &p->steppers[0]), "ARKodeCreateSUNStepper");
  // ARKStep partition: ARS(2,2,2) without embedding -- the explicit table for fe, the implicit one for fi
  p->arkstep_mem = ARKStepCreate(fe, fi, 0.0, p->y, p->sunctx);
  CHKP(p->arkstep_mem, "ARKStepCreate");
  void* as = p->arkstep_mem;
  CHK(ARKodeSetUserData(as, &ud), "ARKodeSetUserData");
  CHK(ARKodeSetFixedStep(as, uo.fixed_h), "ARKodeSetFixedStep");
  CHK(ARKodeSetMaxNumSteps(as, uo.maxsteps), "ARKodeSetMaxNumSteps");
  if (ud.impl_reaction)
  { // ...2d.cpp:1207-1246
    CHK(ARKodeSStolerances(as, uo.rtol, uo.atol), "ARKodeSStolerances");
    if (attach_implicit_reaction(p, as, true)) return -1;
  }
  ARKodeButcherTable Be = fe ? ars222_explicit(false) : nullptr;
  ARKodeButcherTable Bi = fi ? ars222_implicit(false) : nullptr;
  CHK(ARKStepSetTables(as, 2, 0, Bi, Be), "ARKStepSetTables");
  if (Be) ARKodeButcherTable_Free(Be);
  if (Bi) ARKodeButcherTable_Free(Bi);
  CHK(ARKodeCreateSUNStepper(as, &p->steppers[1]), "ARKodeCreateSUNStepper");
  // SplittingStep with Strang coefficients
  p->arkode_mem = SplittingStepCreate(p->steppers, 2, 0.0, p->y, p->sunctx);
  CHKP(p->arkode_mem, "SplittingStepCreate");
  void* mem = p->arkode_mem;
  CHK(ARKodeSetFixedStep(mem, uo.fixed_h), "ARKodeSetFixedStep");
  CHK(ARKodeSetUserData(mem, &ud), "ARKodeSetUserData");
  SplittingStepCoefficients co = SplittingStepCoefficients_LoadCoefficientsByName("ARKODE_SPLITTING_STRANG_2_2_2");
  CHKP(co, "SplittingStepCoefficients_LoadCoefficientsByName");
  CHK(SplittingStepSetCoefficients(mem, co), "SplittingStepSetCoefficients");
  SplittingStepCoefficients_Destroy(&co);
  CHK(ARKodeSetMaxNumSteps(mem, uo.maxsteps), "ARKodeSetMaxNumSteps");
  CHK(ARKodeSetStopTime(mem, ud.tf), "ARKodeSetStopTime");
  return 0;
}

} // namespace

extern "C" int b200_adr_destroy(b200_adr* p);

extern "C" int b200_adr_create(int argc, const char* const* argv, int device, void* stream, b200_adr** out)
{
  b200_adr* p = new b200_adr();
  N_VGetStats_B200(&p->vs0);
  p->launches0 = b200_launch_count();
  std::vector<std::string> args(argv, argv + argc);
  if (read_inputs(args, p->ud, p->uo)) { delete p; return -1; }
  if (b200_ctx_create(device, stream, &p->ctx))
  {
    fprintf(stderr, "b200_adr_create: %s\n", b200_last_error());
    delete p;
    return -1;
  }
  p->ud.ctx = p->ctx;
  for (int m = 0; m < 8; m++)
  {
    p->ud.ops[m].ud       = &p->ud;
    p->ud.ops[m].mode     = m;
    p->ud.ops[m].op.self  = &p->ud.ops[m];
    p->ud.ops[m].op.fused = adr_fused;
    p->ud.ops[m].op.fused_ewt = nullptr;
    p->ud.ops[m].op.chain        = nullptr;
    p->ud.ops[m].op.chain_max    = 0;
    p->ud.ops[m].op.halo_doubles = 0;
    p->ud.ops[m].op.halo_alloc   = nullptr;
    p->ud.ops[m].op.halo_free    = nullptr;
    p->ud.ops[m].op.dq           = nullptr;
    p->ud.ops[m].op.chain_head   = nullptr;
  }
  int depth = -1; // -1: this session does not care about the process-wide chain depth
  if (p->uo.sts_chain >= 2 && p->ud.nx >= 64 && p->ud.ny >= 16)
  { // the pure diffusion operator (Strang's STS partition) may be chained
    depth                     = p->uo.sts_chain > B200_MAX_CHAIN ? B200_MAX_CHAIN : p->uo.sts_chain;
    p->ud.ops[2].op.chain     = adr_chain_cb;
    p->ud.ops[2].op.chain_max = depth;
  }
  // process-wide switches: refuse to change them under another live session (nvector_b200.h)
  if (N_VAcquireSettings_B200(p->uo.no_fusion ? 0 : 1, depth, -1, -1)) { b200_adr_destroy(p); return -1; }
  p->settings_held = true;
  if (SUNContext_Create(SUN_COMM_NULL, &p->sunctx)) { b200_adr_destroy(p); return -1; }
  p->y = N_VNew_B200(p->ctx, p->ud.neq, p->ud.neq, p->sunctx); // ...2d.cpp:83
  if (!p->y) { b200_adr_destroy(p); return -1; }
  if (set_ic(p->y, p->ud)) { b200_adr_destroy(p); return -1; }                           // ...2d.cpp:86
  int rc = -1;
  switch (p->uo.integrator)
  { // ...2d.cpp:125-134
  case 0: rc = setup_erk(p); break;
  case 1: rc = setup_ark(p); break;
  case 2: rc = setup_extsts(p); break;
  case 3: rc = setup_strang(p); break;
  }
  if (rc) { b200_adr_destroy(p); return -1; }
  *out = p;
  return 0;
}

extern "C" int b200_adr_destroy(b200_adr* p)
{
  if (!p) return 0;
  // ...2d.cpp:268-300
  if (p->uo.integrator == 2)
  {
    if (p->lsrkstep_mem) ARKodeFree(&p->lsrkstep_mem);
    delete p->inner_content;
    if (p->inner) MRIStepInnerStepper_Free(&p->inner);
    if (p->arkode_mem) ARKodeFree(&p->arkode_mem);
  }
  else if (p->uo.integrator == 3)
  {
    if (p->lsrkstep_mem) ARKodeFree(&p->lsrkstep_mem);
    if (p->arkstep_mem) ARKodeFree(&p->arkstep_mem);
    if (p->steppers[0]) SUNStepper_Destroy(&p->steppers[0]);
    if (p->steppers[1]) SUNStepper_Destroy(&p->steppers[1]);
    if (p->arkode_mem) ARKodeFree(&p->arkode_mem);
  }
  else if (p->arkode_mem) ARKodeFree(&p->arkode_mem);
  if (p->LS) SUNLinSolFree(p->LS);
  if (p->A) SUNMatDestroy(p->A);
  if (p->y) N_VDestroy(p->y);
  if (p->sunctx) SUNContext_Free(&p->sunctx);
  if (p->ctx) b200_ctx_destroy(p->ctx);
  if (p->settings_held) N_VReleaseSettings_B200();
  delete p;
  return 0;
}

extern "C" int b200_adr_evolve(b200_adr* p, double tout)
{
  const double t0 = wall_seconds();
  int flag        = ARKodeEvolve(p->arkode_mem, tout, p->y, &p->t, ARK_NORMAL);
  b200_ctx_sync(p->ctx);
  p->evolve_seconds += wall_seconds() - t0;
  CHK(flag, "ARKodeEvolve");
  return 0;
}

extern "C" int b200_adr_step(b200_adr* p, int nsteps)
{
  const double t0 = wall_seconds();
  for (int k = 0; k < nsteps; k++)
  {
    int flag = ARKodeEvolve(p->arkode_mem, p->ud.tf, p->y, &p->t, ARK_ONE_STEP);
    if (flag < 0) { fprintf(stderr, "ERROR: ARKodeEvolve returned %d\n", flag); return -1; }
  }
  p->evolve_seconds += wall_seconds() - t0;
  return 0;
}

extern "C" int b200_adr_get_state(b200_adr* p, double* host) { return N_VCopyToHost_B200(p->y, host); }

extern "C" int b200_adr_set_state(b200_adr* p, const double* host, double t)
{
  if (N_VCopyFromHost_B200(p->y, host)) return -1;
  CHK(ARKodeReset(p->arkode_mem, t, p->y), "ARKodeReset");
  p->t = t;
  return 0;
}

extern "C" int b200_adr_get_stats(b200_adr* p, b200_adr_stats* s)
{
  memset(s, 0, sizeof(*s));
  s->t              = p->t;
  s->evolve_seconds = p->evolve_seconds;
  void* mem         = p->arkode_mem;
  ARKodeGetNumSteps(mem, &s->steps);
  ARKodeGetNumStepAttempts(mem, &s->step_attempts);
  if (p->uo.integrator != 3) ARKodeGetNumRhsEvals(mem, 0, &s->rhs_evals_explicit);
  if (p->uo.integrator == 1)
  {
    long nfi = 0, nfils = 0;
    ARKodeGetNumRhsEvals(mem, 1, &nfi);
    ARKodeGetNumLinRhsEvals(mem, &nfils);
    s->rhs_evals_implicit = nfi + nfils;
  }
  if (p->lsrkstep_mem)
  {
    int ms = 0;
    ARKodeGetNumSteps(p->lsrkstep_mem, &s->lsrk_steps);
    ARKodeGetNumRhsEvals(p->lsrkstep_mem, 0, &s->lsrk_rhs_evals);
    LSRKStepGetMaxNumStages(p->lsrkstep_mem, &ms);
    s->lsrk_max_stages = ms;
  }
  if (p->arkstep_mem)
  {
    ARKodeGetNumSteps(p->arkstep_mem, &s->ark_steps);
    ARKodeGetNumRhsEvals(p->arkstep_mem, 0, &s->ark_rhs_evals);
    if (p->ud.impl_reaction) ARKodeGetNumRhsEvals(p->arkstep_mem, 1, &s->ark_rhs_evals_implicit);
  }
  if (p->uo.integrator == 2 && p->ud.impl_reaction) ARKodeGetNumRhsEvals(mem, 1, &s->rhs_evals_implicit);
  if (p->uo.integrator == 1 || p->ud.impl_reaction)
  {
    void* im = (p->uo.integrator == 3) ? p->arkstep_mem : mem; // who owns the implicit partition
    ARKodeGetNumNonlinSolvIters(im, &s->nls_iters);
    ARKodeGetNumLinSolvSetups(im, &s->ls_setups);
    ARKodeGetNumJacEvals(im, &s->jac_evals);
  }
  B200VecStats vs;
  N_VGetStats_B200(&vs);
  s->fused_launches     = vs.fused_launches - p->vs0.fused_launches;
  s->plain_rhs_launches = vs.plain_rhs_launches - p->vs0.plain_rhs_launches;
  s->aliased_copies     = vs.aliased_copies - p->vs0.aliased_copies;
  s->buffers_allocated  = vs.buffers_allocated - p->vs0.buffers_allocated;
  s->kernel_launches    = b200_launch_count() - p->launches0;
  s->nx = p->ud.nx; s->ny = p->ud.ny; s->neq = p->ud.neq;
  return 0;
}

extern "C" int b200_adr_print_stats(b200_adr* p)
{
  // OutputStats*, ...2d.hpp:262-405
  switch (p->uo.integrator)
  {
  case 0:
  case 1:
  {
    b200_adr_stats s;
    b200_adr_get_stats(p, &s);
    long netf = 0;
    ARKodeGetNumErrTestFails(p->arkode_mem, &netf);
    printf("  Steps              = %ld\n  Step attempts      = %ld\n  Error test fails   = %ld\n", s.steps,
           s.step_attempts, netf);
    if (p->uo.integrator == 0) printf("  RHS evals          = %ld\n", s.rhs_evals_explicit);
    else
    {
      long nni = 0, ncfn = 0, nsetups = 0, nje = 0;
      ARKodeGetNumNonlinSolvIters(p->arkode_mem, &nni);
      ARKodeGetNumNonlinSolvConvFails(p->arkode_mem, &ncfn);
      ARKodeGetNumLinSolvSetups(p->arkode_mem, &nsetups);
      ARKodeGetNumJacEvals(p->arkode_mem, &nje);
      printf("  Explicit RHS evals = %ld\n  Implicit RHS evals = %ld\n", s.rhs_evals_explicit, s.rhs_evals_implicit);
      printf("  NLS iters          = %ld\n  NLS fails          = %ld\n  LS setups          = %ld\n  J evals            = %ld\n\n",
             nni, ncfn, nsetups, nje);
    }
    break;
  }
  case 2:
    printf("\nExtSTS Integrator:\n");
    ARKodePrintAllStats(p->arkode_mem, stdout, SUN_OUTPUTFORMAT_TABLE);
    printf("\n\nInner STS Method:\n");
    ARKodePrintAllStats(p->lsrkstep_mem, stdout, SUN_OUTPUTFORMAT_TABLE);
    break;
  case 3:
    printf("\nStrang Integrator:\n");
    ARKodePrintAllStats(p->arkode_mem, stdout, SUN_OUTPUTFORMAT_TABLE);
    printf("\n\nARKStep Stepper:\n");
    ARKodePrintAllStats(p->arkstep_mem, stdout, SUN_OUTPUTFORMAT_TABLE);
    printf("\n\nLSRKStep Stepper:\n");
    ARKodePrintAllStats(p->lsrkstep_mem, stdout, SUN_OUTPUTFORMAT_TABLE);
    printf("\n");
    break;
  }
  return 0;
}

extern "C" int b200_adr_main(int argc, char** argv)
{
  for (int k = 1; k < argc; k++)
    if (std::string(argv[k]) == "--help")
    {
      printf("options: see /root/reference/adr/advection_diffusion_reaction_2d.hpp InputHelp (same flags), plus --no-fusion\n");
      return 0;
    }
  const char* dev = getenv("B200_DEVICE");
  b200_adr* p     = nullptr;
  if (b200_adr_create(argc - 1, argv + 1, dev ? atoi(dev) : 0, nullptr, &p)) return 1;
  AdrData& ud    = p->ud;
  AdrOptions& uo = p->uo;
  static const char* names[] = {"ERK", "ARK", "ExtSTS", "Strang"};
  printf("\nProblem parameters and options (B200, N_Vector_B200):\n");
  printf("  cux = %g  cuy = %g  cvx = %g  cvy = %g  d = %g  A = %g  B = %g\n", ud.cux, ud.cuy, ud.cvx, ud.cvy, ud.d,
         ud.A, ud.B);
  printf("  tf = %g  nx = %lld  ny = %lld  dx = %.17g  dy = %.17g\n", ud.tf, (long long)ud.nx, (long long)ud.ny, ud.dx,
         ud.dy);
  printf("  integrator = %s  sts_method = %d  fixed h = %g  rtol = %g  atol = %g\n\n", names[uo.integrator],
         uo.sts_method, uo.fixed_h, uo.rtol, uo.atol);
  // normal mode, ...2d.cpp:217-239
  const double dTout = ud.tf / uo.nout;
  double tout        = dTout;
  for (int iout = 0; iout < uo.nout; iout++)
  {
    if (uo.output == 3 && ARKodeSetStopTime(p->arkode_mem, tout) < 0) return 1;
    if (b200_adr_evolve(p, tout)) return 1;
    tout += dTout;
    tout = (tout > ud.tf) ? ud.tf : tout;
  }
  if (uo.output)
  { // WriteOutput, ...2d.hpp:794-829: t, all u (row-major), all v
    std::vector<double> h((size_t)ud.neq);
    if (b200_adr_get_state(p, h.data())) return 1;
    FILE* f = fopen("solution.dat", "w");
    if (!f) return 1;
    fprintf(f, "%.15g", ud.tf);
    for (int s = 0; s < 2; s++)
      for (int64_t j = 0; j < ud.ny; j++)
        for (int64_t i = 0; i < ud.nx; i++) fprintf(f, " %.15g", h[(size_t)(2 * (i + j * ud.nx) + s)]);
    fprintf(f, "\n");
    fclose(f);
    printf("Solution is written to solution.dat\n");
    printf("Final integrator statistics:\n");
    printf("  Total solve time   = %.6f\n", p->evolve_seconds);
    b200_adr_print_stats(p);
    b200_adr_stats s;
    b200_adr_get_stats(p, &s);
    printf("B200 fused launches           = %ld\n", s.fused_launches);
    printf("B200 plain RHS launches       = %ld\n", s.plain_rhs_launches);
    printf("B200 aliased copies           = %ld\n", s.aliased_copies);
    printf("B200 kernel launches          = %llu\n", (unsigned long long)s.kernel_launches);
  }
  else if (getenv("B200_STATS"))
  { // the reference prints nothing with --output 0; B200_STATS=1 still reports the timing
    b200_adr_stats s;
    b200_adr_get_stats(p, &s);
    printf("Total solve time = %.6f  steps = %ld  lsrk rhs evals = %ld  lsrk max stages = %ld  ark rhs evals = %ld  "
           "outer explicit rhs evals = %ld  fused launches = %ld  kernel launches = %llu\n",
           p->evolve_seconds, s.steps, s.lsrk_rhs_evals, s.lsrk_max_stages, s.ark_rhs_evals, s.rhs_evals_explicit,
           s.fused_launches, (unsigned long long)s.kernel_launches);
  }
  b200_adr_destroy(p);
  return 0;
}
