// nvector_b200.cpp -- N_Vector ops table over B200 HBM with lazy stage fusion.
//
// Replaces SUN/src/nvector/parallel/nvector_parallel.c (the reference's backend for
// diffusion_2D) and nvector_serial.c (adr) behind the unchanged ops table
// (SUN/include/sundials/sundials_nvector.h:98-192).  Host C++ only; all device work
// goes through the C-ABI in include/b200_sts.h.  There is no CPU fallback: a failed
// device call aborts loudly.
//
// Value model: every vector points at an immutable, reference-counted Value that is
// either materialised (device buffer from a per-length pool) or deferred (op applied
// to another Value).  All N_V* ops that write z build a NEW Value and re-point z, so
//   * N_VScale(1,x,z) is a handle share (arkode_lsrkstep.c:642,746,2242; arkode.c:2737),
//   * in-place forms such as N_VLinearSum(1,y,c,F,y) with F = L(y) (SSP stages,
//     arkode_lsrkstep.c:1213-1231) are automatically ping-ponged,
//   * ARKODE swapping its tempv1/tempv2 handles (arkode_lsrkstep.c:742-744) is harmless.
//
// Temporal blocking.  An STS stage  z = c0 L(x) + c1 p + c2 yn + c3 x + c4 fn
// (arkode_lsrkstep.c:706-717 RKC, :1001-1012 RKL) is not launched when it is requested either: z
// becomes a PENDING STAGE value.  If the next stage is built on it with the matching operands
// (its p is this stage's x, same yn / fn) the chain grows; at the configured depth, or as soon
// as anything else needs one of the values, the whole chain goes out as ONE kernel
// (B200RhsOp::chain -> b200_stencil_chain) that stores only the levels somebody still points at.

#include "nvector_b200.h"

#include <sundials/sundials_core.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

namespace {

// A device failure inside a vector operation: reported once on stderr, remembered (g_failed: every later operation
// returns at once -- the state behind the vectors can no longer be trusted), and thrown up to the boundary of this
// library, where it becomes what the interface can express: SUN_ERR_EXT_FAIL from the fused operations (LSRKStep maps it to
// ARK_VECTOROP_ERR, arkode_lsrkstep.c:718-723), NaN from the reductions (the step fails its error test and ARKODE
// gives up with an error return), NULL / -1 from the accessors.  Nothing is recomputed on the host: there is no
// fallback, only an orderly way down instead of abort().
struct DeviceFailure
{
};
bool g_failed = false;
[[noreturn]] void die(const char* what, int rc)
{
  fprintf(stderr, "nvector_b200: FATAL: %s failed (code %d): %s\n", what, rc, b200_last_error());
  g_failed = true;
  throw DeviceFailure();
}
#define DEV(call)                       \
  do {                                  \
    int rc_ = (call);                   \
    if (rc_ != 0) die(#call, rc_);      \
  }                                     \
  while (0)

B200VecStats g_stats = {0, 0, 0, 0, 0, 0, 0, 0, 0};
bool g_lazy          = true;
int g_chain_max      = 4; // stages per temporally blocked launch (1 = off)

// buffers of one local length on one context, shared by all clones
struct Shared
{
  b200_ctx* ctx;
  sunindextype nloc, nglob;
  int refs;
  std::vector<double*> free_bufs;
  std::vector<double*> free_halos; // deep-halo buffers (one size per problem)
  int64_t halo_doubles;
  double* wrms_slots; // scalars the fused WRMS partial sums are stored to (device memory, or the device alias of wrms_host)
  double* wrms_host;  // one rank: the slots live in mapped pinned host memory and are read after a stream sync
  int next_slot;
  long slot_owner[8]; // creation number of the value whose partial sum the slot holds (the ring is reused)
  bool slot_unread[8]; // a launch was given this slot and nobody has read (= waited for) its result yet
  bool spec_sigs[96]; // fused-launch signatures whose result a matching WRMS norm followed
  // the next step's error weights, speculatively (see launch_fused): signatures of fused launches whose stencil input
  // ARKODE then asked the error weights of, and the tolerances of the most recent such request
  bool spec_ewt_sigs[96];
  bool ewt_seen;
  double ewt_rtol, ewt_atol;
};
const int kSlots = 8;

struct Value;

// a requested-but-not-launched STS stage:  c[0] L(x) + c[1] p2 + c[2] yn + c[3] x + c[4] fn
struct StageRec
{
  const B200RhsOp* op;
  Value *x, *p2, *yn, *fn; // one reference held on each
  double c[5];
  int depth;               // 1 for the first stage of a chain
  bool head;               // stage 1 of a step: z = x + c[0] L(x); fn is the deferred value L(x) itself (yn == p2 == x)
};

// a requested-but-not-evaluated elementwise result (the implicit path's vector work): evaluated when something
// reads it -- alone, or inside the reduction / stencil kernel that consumes it
enum EwKind { EWK_LIN2 = 1, EWK_SCALESUM, EWK_SCALEDIFF, EWK_PROD, EWK_ABS, EWK_SCALE, EWK_ADDCONST, EWK_INV, EWK_EWT };
struct EwRec
{
  int kind;       // LIN2: ca*a + cb*b ; SCALESUM / SCALEDIFF: ca*(a +- b) ; PROD: a.*b
                  // unary (b == nullptr): ABS |a| ; SCALE ca*a ; ADDCONST a + cb ; INV 1/a ; EWT 1/(ca*|a| + cb)
  Value *a, *b;   // one reference held on each
  double ca, cb;
};

struct Value
{
  int refs;
  double* d;           // device data, nullptr while deferred / pending
  const B200RhsOp* op; // deferred: value = op(src)
  Value* src;
  StageRec* st;        // pending stage (d == nullptr, op == nullptr)
  EwRec* ew;           // pending elementwise result (d == nullptr, op == nullptr, st == nullptr)
  bool is_const;       // every entry equals cval (N_VConst); the buffer is only filled if somebody needs one
  double cval;
  double* halo;        // deep halo of this value (multi-rank temporal blocking), filled on demand
  bool halo_valid;
  const B200RhsOp* halo_op; // who handed out `halo` (halo_alloc / halo_free); nullptr = this vector's pool
  // fused WRMS partial: sum (this_i * w_i)^2 already sits in slot
  Value* wrms_w;
  int wrms_slot;
  int sig;  // signature of the fused launch that produced it (0 = none)
  long seq; // creation order
  // provenance of stored data: d holds prov_op(prov_src), bit for bit as the operator's kernels compute it (one
  // reference held on prov_src).  Lets an adaptive STS step start its first chain from y_n alone (launch_chain, HEAD).
  const B200RhsOp* prov_op;
  Value* prov_src;
  int centre_sig;       // signature of the fused launch this value was the stencil input (and a term) of
  bool spec_ewt;        // this value is 1/(e_rtol*|y| + e_atol) of the value y whose wrms_w it is (computed ahead of the request)
  double e_rtol, e_atol;
};

struct Content
{
  Shared* sh;
  Value* val;
  double* host;    // pinned mirror for N_VGetArrayPointer
  bool host_dirty; // host mirror may have been written since it was handed out
};

inline Content* C(N_Vector v) { return static_cast<Content*>(v->content); }

// weight value of the most recent WRMS norm (one reference held) and when it was set
Value* g_last_weight     = nullptr;
Shared* g_last_weight_sh = nullptr;
long g_last_weight_seq   = 0;
long g_seq               = 0;

// Fused WRMS results on one rank live in mapped pinned memory.  The host waits for the VALUE, not for the stream: a
// slot is armed with a NaN bit pattern no sum produces when it is handed to a launch, and read_slot spins on the word
// until the kernel's last block has stored the result (a stream synchronisation costs several microseconds after the
// kernel has ended, every adaptive step); after ~2 ms of spinning, or with B200_NO_POLL, it synchronises instead.
const unsigned long long kArmed = 0x7ff8dead0000beefULL;
int g_poll = -1;
void arm_slot(Shared* sh, int slot)
{
  if (!sh->wrms_host) return;
  // The launch that had this slot before may still be queued if its result was never asked for (a speculation that
  // missed): it would overwrite the armed word with ITS sum and the poll would take that for the new result.  Every
  // successful wait drains the stream up to its kernel, so this is rare; when it can happen, drain first.
  if (sh->slot_unread[slot]) DEV(b200_ctx_sync(sh->ctx));
  *reinterpret_cast<volatile unsigned long long*>(sh->wrms_host + slot) = kArmed;
  sh->slot_unread[slot] = true;
}
double read_slot(Shared* sh, int slot)
{
  if (g_poll < 0) g_poll = getenv("B200_NO_POLL") ? 0 : 1;
  volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(sh->wrms_host + slot);
  const bool prof = b200_host_profile_on() != 0;
  timespec t0;
  if (prof) clock_gettime(CLOCK_MONOTONIC, &t0);
  if (g_poll == 1)
    for (int spin = 0; spin < 40000; spin++)
    {
      const unsigned long long b = *w;
      if (b != kArmed)
      {
        double r;
        memcpy(&r, &b, sizeof(r));
        for (int k = 0; k < kSlots; k++) sh->slot_unread[k] = false; // (stream order: everything launched before it is done)
        if (prof)
        {
          timespec t1;
          clock_gettime(CLOCK_MONOTONIC, &t1);
          b200_host_profile_wait(1e3 * (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0.tv_nsec));
        }
        return r;
      }
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
  DEV(b200_ctx_sync(sh->ctx));
  for (int k = 0; k < kSlots; k++) sh->slot_unread[k] = false;
  return sh->wrms_host[slot];
}

double* pool_get(Shared* sh)
{
  if (!sh->free_bufs.empty())
  {
    double* p = sh->free_bufs.back();
    sh->free_bufs.pop_back();
    return p;
  }
  double* p = nullptr;
  DEV(b200_malloc(sh->ctx, sh->nloc, &p));
  g_stats.buffers_allocated++;
  return p;
}

void value_release(Shared* sh, Value* v)
{
  while (v)
  {
    if (--v->refs > 0) return;
    if (v->d) sh->free_bufs.push_back(v->d);
    if (v->halo)
    {
      if (v->halo_op) v->halo_op->halo_free(v->halo_op->self, v->halo);
      else sh->free_halos.push_back(v->halo);
    }
    if (v->wrms_w) value_release(sh, v->wrms_w);
    if (v->prov_src) value_release(sh, v->prov_src);
    Value* next = v->src; // a deferred value owns a reference on its source
    if (v->ew)
    { // a pending elementwise result that was never needed
      EwRec* e = v->ew;
      if (e->b) value_release(sh, e->b);
      next = e->a;
      delete e;
    }
    if (v->st)
    { // a pending stage that was never needed: drop its operands
      StageRec* r = v->st;
      value_release(sh, r->p2);
      value_release(sh, r->yn);
      value_release(sh, r->fn);
      next = r->x; // (src is null for a pending stage)
      delete r;
    }
    delete v;
    v = next;
  }
}

Value* value_new(Shared* sh, bool with_buffer)
{
  Value* v     = new Value();
  v->refs      = 1;
  v->d         = with_buffer ? pool_get(sh) : nullptr;
  v->op        = nullptr;
  v->src       = nullptr;
  v->st        = nullptr;
  v->ew        = nullptr;
  v->is_const  = false;
  v->cval      = 0.0;
  v->halo      = nullptr;
  v->halo_valid = false;
  v->halo_op   = nullptr;
  v->wrms_w    = nullptr;
  v->wrms_slot = -1;
  v->sig       = 0;
  v->seq       = ++g_seq;
  v->prov_op   = nullptr;
  v->prov_src  = nullptr;
  v->centre_sig = 0;
  v->spec_ewt  = false;
  v->e_rtol = v->e_atol = 0.0;
  return v;
}

void assign(Content* c, Value* nv) // takes ownership of one reference on nv
{
  Value* old = c->val;
  c->val     = nv;
  if (old) value_release(c->sh, old);
  c->host_dirty = false;
}

void materialise(Shared* sh, Value* v);
void force_ew(Shared* sh, Value* v);

// make sure the host mirror (if it was handed out) is reflected on the device
void sync_from_host(Content* c)
{
  if (c->host_dirty)
  {
    Value* nv = value_new(c->sh, true);
    DEV(b200_h2d(c->sh->ctx, nv->d, c->host, c->sh->nloc));
    assign(c, nv);
  }
  if (!c->val)
  { // never written: SUNDIALS would read uninitialised memory; give zeros
    Value* nv = value_new(c->sh, true);
    DEV(b200_const(c->sh->ctx, 0.0, nv->d, c->sh->nloc));
    assign(c, nv);
  }
}

// launch the deferred operator (optionally fused with a linear combination)
//   out = sum_k cf[k] * T_k ; terms equal to L become the stencil term
void launch_fused(Shared* sh, Value* L, int nterms, const double* cf, Value* const* X,
                  Value* out, bool store_f)
{
  Value* src = L->src;
  materialise(sh, src);
  int srcs[B200_MAX_TERMS];
  const double* vp[B200_MAX_TERMS];
  int lpos = -1;
  for (int k = 0; k < nterms; k++)
  {
    if (X[k] == L) { srcs[k] = B200_SRC_STENCIL; vp[k] = nullptr; if (lpos < 0) lpos = k; }
    else if (X[k] == src) { srcs[k] = B200_SRC_CENTRE; vp[k] = nullptr; }
    else { srcs[k] = B200_SRC_VECTOR; vp[k] = X[k]->d; }
  }
  double* f_out = nullptr;
  if (store_f) { L->d = pool_get(sh); f_out = L->d; }
  // speculative WRMS: only for launch signatures a norm has followed before
  const int sig        = 1 + nterms * 8 + lpos;
  const double* w      = nullptr;
  double* wres         = nullptr;
  Value* wv            = nullptr;
  int slot             = -1;
  if (out && sh->spec_sigs[sig] && sh->wrms_slots)
  {
    // weight guess: the value the most recent N_VWrmsNorm used, kept alive by a ref
    if (g_last_weight && g_last_weight_sh == sh && g_last_weight->d)
    {
      wv   = g_last_weight;
      w    = wv->d;
      slot = sh->next_slot;
      sh->next_slot = (sh->next_slot + 1) % kSlots;
      wres = sh->wrms_slots + slot;
      arm_slot(sh, slot);
    }
  }
  // The closing stage of an adaptive STS step (arkode_lsrkstep.c:768-796) has the candidate y_{n+1} as its stencil
  // input; if the step is accepted ARKODE next asks for ewt = 1/(rtol |y_{n+1}| + atol) and ||y_{n+1}||_wrms
  // (arkode.c:2985, :835).  Once that has been seen to follow a launch of this signature, the launch produces both as
  // well: no launch and no host synchronisation between two steps.
  bool centre = false;
  for (int k = 0; k < nterms; k++) centre = centre || (X[k] == src);
  Value* E   = nullptr;
  int slot2  = -1;
  int edone  = 0;
  int wdone  = 0;
  if (w && centre && L->op->fused_ewt && sh->ewt_seen && sh->spec_ewt_sigs[sig] && !src->wrms_w && src->wrms_slot < 0)
  {
    E      = value_new(sh, true);
    slot2  = sh->next_slot;
    sh->next_slot = (sh->next_slot + 1) % kSlots;
    arm_slot(sh, slot2);
    DEV(L->op->fused_ewt(L->op->self, sh->ctx, src->d, nterms, cf, srcs, vp, out->d, f_out, w, wres, &wdone, sh->ewt_rtol,
                         sh->ewt_atol, E->d, sh->wrms_slots + slot2, &edone));
    if (edone && wdone)
    {
      E->spec_ewt = true;
      E->e_rtol = sh->ewt_rtol; E->e_atol = sh->ewt_atol;
      src->wrms_w    = E; // (takes over the creation reference)
      src->wrms_slot = slot2;
      sh->slot_owner[slot2] = src->seq;
      g_stats.ew_fused++;
    }
    else value_release(sh, E);
  }
  else
    DEV(L->op->fused(L->op->self, sh->ctx, src->d, nterms, cf, srcs, vp, out ? out->d : nullptr, f_out,
                     w, wres, &wdone));
  if (centre) src->centre_sig = sig;
  if (out)
  {
    out->sig = sig;
    if (wdone)
    {
      out->wrms_w = wv;
      wv->refs++;
      out->wrms_slot = slot;
      sh->slot_owner[slot] = out->seq;
    }
  }
  if (store_f)
  { // L is now a plain materialised value (that remembers what it is the image of: its reference on src stays)
    L->prov_op  = L->op;
    L->prov_src = src;
    L->op  = nullptr;
    L->src = nullptr;
  }
}

// Launch the chain of pending stages that ends in `top` as one kernel (or, for a single
// stage, as the ordinary fused stage launch).
void launch_chain(Shared* sh, Value* top)
{
  Value* lv[B200_MAX_CHAIN] = {}; // lv[0] = first stage of the chain ... lv[n-1] = top
  int n = 0;
  {
    Value* rev[B200_MAX_CHAIN];
    Value* u = top;
    while (u && !u->d && u->st && n < B200_MAX_CHAIN) { rev[n++] = u; u = u->st->x; }
    for (int k = 0; k < n; k++) lv[k] = rev[n - 1 - k];
  }
  if (n == 0) return; // nothing pending (callers only come here with a pending stage)
  StageRec* first = lv[0]->st;
  const bool head = first->head; // the chain begins the step: f_n = L(x) is produced by this launch
  materialise(sh, first->x); // (only reachable if a chain longer than B200_MAX_CHAIN was built)
  if (!head)
  {
    materialise(sh, first->p2);
    materialise(sh, first->yn);
    materialise(sh, first->fn);
  }
  // a level must be stored iff somebody outside this chain still points at it: chain-internal
  // references are the next stage's x and the stage after that's p2
  double* outs[B200_MAX_CHAIN];
  double cf[B200_MAX_CHAIN * 5];
  for (int k = 0; k < n; k++)
  {
    int internal = 0;
    if (k + 1 < n) internal++;
    if (k + 2 < n) internal++;
    const bool keep = (k == n - 1) || (lv[k]->refs - internal > 0);
    outs[k]         = keep ? pool_get(sh) : nullptr;
    for (int q = 0; q < 5; q++) cf[5 * k + q] = lv[k]->st->c[q];
  }
  Value* F            = head ? first->fn : nullptr; // the value L(x) that ARKODE keeps as fn: still deferred, or stored
  const bool have_f   = head && F->d != nullptr;    // (an adaptive step: it came out of the previous closing stage)
  double* f_out       = (head && !have_f) ? pool_get(sh) : nullptr;
  if (n == 1)
  {
    int wdone = 0;
    if (head && have_f)
    { // a lone stage 1 whose f_n is stored: the plain two-term combination
      const double c2[2]  = {first->c[3], first->c[0]};
      const double* vp[2] = {first->x->d, F->d};
      DEV(b200_lincomb(sh->ctx, 2, c2, vp, outs[0], sh->nloc));
    }
    else if (head)
    { // a lone stage 1: the ordinary fused launch z = 1*x + c*L(x) that also stores L(x)
      const double c2[2]  = {first->c[3], first->c[0]};
      int srcs[2]         = {B200_SRC_CENTRE, B200_SRC_STENCIL};
      const double* vp[2] = {nullptr, nullptr};
      DEV(first->op->fused(first->op->self, sh->ctx, first->x->d, 2, c2, srcs, vp, outs[0], f_out, nullptr, nullptr, &wdone));
    }
    else
    {
      int srcs[5]         = {B200_SRC_STENCIL, B200_SRC_VECTOR, B200_SRC_VECTOR, B200_SRC_CENTRE, B200_SRC_VECTOR};
      const double* vp[5] = {nullptr, first->p2->d, first->yn->d, nullptr, first->fn->d};
      DEV(first->op->fused(first->op->self, sh->ctx, first->x->d, 5, cf, srcs, vp, outs[0], nullptr, nullptr, nullptr,
                           &wdone));
    }
  }
  else
  {
    double* halos[4] = {nullptr, nullptr, nullptr, nullptr};
    int valid[4]     = {1, 1, 1, 1};
    Value* opv[4]    = {first->x, first->p2, first->yn, first->fn};
    const int nop    = head ? 1 : 4; // a chain that begins the step reads x only
    const bool deep  = first->op->halo_doubles > 0;
    if (deep)
    { // multi-rank: each operand carries a deep halo, exchanged by the operator when stale
      if (sh->halo_doubles != first->op->halo_doubles)
      {
        for (double* p : sh->free_halos) b200_free(sh->ctx, p);
        sh->free_halos.clear();
        sh->halo_doubles = first->op->halo_doubles;
      }
      for (int q = 0; q < nop; q++)
      {
        if (!opv[q]->halo)
        {
          if (first->op->halo_alloc)
          {
            opv[q]->halo    = first->op->halo_alloc(first->op->self);
            opv[q]->halo_op = first->op;
            if (!opv[q]->halo) die("halo_alloc", -1);
          }
          else if (!sh->free_halos.empty()) { opv[q]->halo = sh->free_halos.back(); sh->free_halos.pop_back(); }
          else DEV(b200_malloc(sh->ctx, sh->halo_doubles, &opv[q]->halo));
          opv[q]->halo_valid = false;
        }
        halos[q] = opv[q]->halo;
        valid[q] = opv[q]->halo_valid ? 1 : 0;
        for (int e = 0; e < q; e++)
          if (opv[e] == opv[q]) valid[q] = 1; // same value twice: exchange it once
      }
    }
    if (head)
      DEV(first->op->chain_head(first->op->self, sh->ctx, n, first->x->d, cf, outs, f_out, deep ? halos[0] : nullptr, valid[0]));
    else
      DEV(first->op->chain(first->op->self, sh->ctx, n, first->x->d, first->p2->d, first->yn->d, first->fn->d, cf, outs,
                           deep ? halos : nullptr, valid));
    if (deep)
      for (int q = 0; q < nop; q++) opv[q]->halo_valid = true;
    g_stats.chain_launches++;
    g_stats.chain_stages += n;
  }
  if (head && !have_f)
  { // f_n is now plain data (as after a fused launch with f_out)
    F->d        = f_out;
    F->prov_op  = F->op;
    F->prov_src = F->src; // (keeps the reference)
    F->op  = nullptr;
    F->src = nullptr;
  }
  // retire the records bottom-up; values nobody points at any more disappear with them
  for (int k = 0; k < n; k++) lv[k]->refs++; // pin while we rewire
  for (int k = 0; k < n; k++)
  {
    StageRec* r = lv[k]->st;
    lv[k]->st   = nullptr;
    lv[k]->d    = outs[k];
    value_release(sh, r->x);
    value_release(sh, r->p2);
    value_release(sh, r->yn);
    value_release(sh, r->fn);
    delete r;
  }
  for (int k = 0; k < n; k++)
  {
    if (lv[k]->refs == 1 && !lv[k]->d) { delete lv[k]; } // unpinned, never stored, unreachable
    else value_release(sh, lv[k]);
  }
}

void materialise(Shared* sh, Value* v)
{
  if (v->d) return;
  if (v->st) { launch_chain(sh, v); return; }
  if (v->ew) { force_ew(sh, v); return; }
  if (v->is_const)
  { // a constant that somebody wants to read element by element after all
    v->d = pool_get(sh);
    DEV(b200_const(sh->ctx, v->cval, v->d, sh->nloc));
    return;
  }
  if (!v->op) die("materialise: value has neither data nor operator", -1);
  // f = 1 * L(src), stored through the f_out path so v itself becomes plain
  const double one = 1.0;
  Value* X[1]      = {v};
  Value* src       = v->src;
  materialise(sh, src);
  int srcs[1]          = {B200_SRC_STENCIL};
  const double* vp[1]  = {nullptr};
  v->d                 = pool_get(sh);
  int wdone            = 0;
  DEV(v->op->fused(v->op->self, sh->ctx, src->d, 1, &one, srcs, vp, v->d, nullptr, nullptr, nullptr, &wdone));
  (void)X;
  v->prov_op  = v->op;
  v->prov_src = src; // (keeps the reference)
  v->op  = nullptr;
  v->src = nullptr;
  g_stats.plain_rhs_launches++;
}

// ---------------------------------------------------------------- pending elementwise results
bool is_ew(const Value* v) { return !v->d && v->ew; }
bool is_rhs(const Value* v) { return !v->d && v->op; }

Value* ew_new(Shared* sh, int kind, double ca, Value* a, double cb, Value* b)
{
  Value* out = value_new(sh, false);
  EwRec* e   = new EwRec();
  e->kind = kind; e->a = a; e->b = b; e->ca = ca; e->cb = cb;
  a->refs++;
  if (b) b->refs++;
  out->ew = e;
  return out;
}

void ew_retire(Shared* sh, Value* v) // v->d has been filled: drop the record
{
  EwRec* e = v->ew;
  v->ew    = nullptr;
  value_release(sh, e->a);
  if (e->b) value_release(sh, e->b);
  delete e;
}

// 1/(rtol*|y| + atol) requested as N_VAbs, N_VScale, N_VAddConst, N_VInv (arkEwtSetSS, arkode.c:2932-2944): the chain
// INV(ADDCONST(SCALE(ABS(y)))) collapses into one EWT node as it is built
Value* ewt_source(const Value* v, double* rtol, double* atol)
{
  if (!is_ew(v) || v->ew->kind != EWK_ADDCONST) return nullptr;
  const Value* s = v->ew->a;
  if (!is_ew(s) || s->ew->kind != EWK_SCALE) return nullptr;
  const Value* a = s->ew->a;
  if (!is_ew(a) || a->ew->kind != EWK_ABS) return nullptr;
  *rtol = s->ew->ca;
  *atol = v->ew->cb;
  return a->ew->a;
}

// The difference-quotient pattern of arkLsATimes / arkLsDQJtimes (arkode_ls.c:2316-2372, :2839-2877):
//   top = LIN2(ca, V, cb, J) [outer] or J itself,  J = SCALEDIFF(siginv, F, FY),  F = deferred op(W),
//   W = LIN2(sigma, V, 1, Y).  All leaves V, Y, FY are (made) materialised; nothing in between is stored.
struct DqMatch
{
  bool ok = false;
  int outer = 0;
  double ca = 0, cb = 0, sigma = 0, siginv = 0;
  Value *V = nullptr, *Y = nullptr, *FY = nullptr, *F = nullptr;
};
DqMatch match_dq(const Value* top)
{
  DqMatch m;
  if (!is_ew(top)) return m;
  const Value* J = top;
  if (top->ew->kind == EWK_LIN2 && is_ew(top->ew->b) && top->ew->b->ew->kind == EWK_SCALEDIFF)
  {
    m.outer = 1; m.ca = top->ew->ca; m.cb = top->ew->cb;
    J = top->ew->b;
  }
  if (J->ew->kind != EWK_SCALEDIFF) return m;
  Value* F = J->ew->a;
  if (!is_rhs(F) || !F->op->dq) return m;
  Value* W = F->src;
  if (!is_ew(W) || W->ew->kind != EWK_LIN2 || W->ew->cb != 1.0) return m;
  m.V = W->ew->a; m.Y = W->ew->b; m.FY = J->ew->b; m.F = F;
  m.sigma = W->ew->ca; m.siginv = J->ew->ca;
  if (m.outer && top->ew->a != m.V) return m;
  m.ok = true;
  return m;
}

// evaluate the difference-quotient pattern in one stencil pass; dot_result != nullptr: also <top, V>
bool run_dq(Shared* sh, Value* top, const DqMatch& m, double* dot_result)
{
  materialise(sh, m.V);
  materialise(sh, m.Y);
  materialise(sh, m.FY);
  double* out = pool_get(sh);
  int rc = m.F->op->dq(m.F->op->self, sh->ctx, m.V->d, m.Y->d, m.FY->d, m.sigma, m.siginv, m.outer, m.ca, m.cb, out, dot_result);
  if (rc != 0)
  {
    sh->free_bufs.push_back(out);
    if (rc < 0) die("B200RhsOp::dq", rc);
    return false;
  }
  top->d = out;
  ew_retire(sh, top);
  g_stats.dq_fused++;
  g_stats.fused_launches++;
  return true;
}

void force_ew(Shared* sh, Value* v)
{
  {
    DqMatch m = match_dq(v);
    if (m.ok && run_dq(sh, v, m, nullptr)) return;
  }
  EwRec* e = v->ew;
  materialise(sh, e->a);
  if (e->b) materialise(sh, e->b);
  double* out = pool_get(sh);
  switch (e->kind)
  {
  case EWK_ABS: DEV(b200_abs(sh->ctx, e->a->d, out, sh->nloc)); break;
  case EWK_SCALE:
  {
    const double* vp[1] = {e->a->d};
    DEV(b200_lincomb(sh->ctx, 1, &e->ca, vp, out, sh->nloc));
    break;
  }
  case EWK_ADDCONST: DEV(b200_addconst(sh->ctx, e->a->d, e->cb, out, sh->nloc)); break;
  case EWK_INV: DEV(b200_inv(sh->ctx, e->a->d, out, sh->nloc)); break;
  case EWK_EWT: DEV(b200_ewt_ss(sh->ctx, e->a->d, e->ca, e->cb, out, sh->nloc)); break;
  case EWK_LIN2:
  {
    const double cf[2]  = {e->ca, e->cb};
    const double* vp[2] = {e->a->d, e->b->d};
    DEV(b200_lincomb(sh->ctx, 2, cf, vp, out, sh->nloc));
    break;
  }
  case EWK_SCALESUM: DEV(b200_scale_sumdiff(sh->ctx, e->ca, e->a->d, e->b->d, +1, out, sh->nloc)); break;
  case EWK_SCALEDIFF: DEV(b200_scale_sumdiff(sh->ctx, e->ca, e->a->d, e->b->d, -1, out, sh->nloc)); break;
  default: DEV(b200_prod(sh->ctx, e->a->d, e->b->d, out, sh->nloc)); break;
  }
  v->d = out;
  ew_retire(sh, v);
}

// operands of a new pending elementwise result: anything that is not plain data is evaluated first, except the two
// shapes the fused kernels consume (keep_a / keep_b say which operand may stay pending)
void ew_operand(Shared* sh, Value* x, bool keep)
{
  if (x->d) return;
  if (keep && (is_ew(x) || is_rhs(x))) return;
  materialise(sh, x);
}

// z = 1*y + c*F with F stored and known to be op(y) for an operator that can begin a chain (see eval_lincomb)
bool head_from_provenance(int nterms, const double* cf, Value* const* X)
{
  return g_lazy && g_chain_max >= 2 && nterms == 2 && cf[0] == 1.0 && X[0]->d && X[1]->d && X[1]->prov_op &&
         X[1]->prov_op->chain_head && X[1]->prov_op->chain_max >= 2 && X[1]->prov_src == X[0] && X[1] != X[0];
}

// z = sum_k cf[k]*X[k], left to right; handles deferred operands by fusion
void eval_lincomb(int nterms, const double* cf, N_Vector* Xv, N_Vector zv)
{
  Content* zc = C(zv);
  Shared* sh  = zc->sh;
  Value* X[B200_MAX_TERMS];
  for (int k = 0; k < nterms; k++)
  {
    sync_from_host(C(Xv[k]));
    X[k] = C(Xv[k])->val;
  }
  // pick the deferred RHS operand to fuse (first one)
  Value* L = nullptr;
  for (int k = 0; k < nterms; k++)
    if (!X[k]->d && X[k]->op && !L) L = X[k];
  if (L)
  {
    int uses = 0;
    for (int k = 0; k < nterms; k++) uses += (X[k] == L);
    if (uses > 1 || !g_lazy) { materialise(sh, L); L = nullptr; }
  }
  // F itself must be kept if anything other than z will still point at it
  const bool store_f = L && (L->refs - (zc->val == L ? 1 : 0)) > 0;

  // Stage 1 of an STS step, z_1 = 1*y_n + c*L(y_n) with L(y_n) = f_n kept by ARKODE (arkode_lsrkstep.c:640 / :930):
  // pending as the HEAD of a chain -- if stage 2 extends it, one launch produces f_n and the stages together.
  // f_n may also be stored already and known to be L(y_n) (provenance: an adaptive step, whose f_n the previous step's
  // closing stage wrote): the chain then recomputes it from y_n instead of streaming it.
  const bool head_prov = head_from_provenance(nterms, cf, X);
  if (g_chain_max >= 2 && nterms == 2 && cf[0] == 1.0 &&
      ((L && store_f && L->op->chain_head && L->op->chain_max >= 2 && X[1] == L && X[0] == L->src) || head_prov))
  {
    Value* F            = X[1];
    const B200RhsOp* op = head_prov ? F->prov_op : F->op;
    Value* xin          = X[0];
    materialise(sh, xin);
    Value* out  = value_new(sh, false);
    StageRec* r = new StageRec();
    r->op = op; r->x = xin; r->p2 = xin; r->yn = xin; r->fn = F;
    r->x->refs++; r->p2->refs++; r->yn->refs++; r->fn->refs++;
    r->c[0] = cf[1]; r->c[1] = r->c[2] = r->c[4] = 0.0; r->c[3] = 1.0;
    r->depth = 1;
    r->head  = true;
    out->st  = r;
    assign(zc, out);
    g_stats.fused_launches++;
    return;
  }

  // STS stage pattern [L(x), p, yn, x, fn]: defer it as a pending stage (temporal blocking)
  if (L && g_chain_max >= 2 && nterms == 5 && !store_f && L->op->chain && L->op->chain_max >= 2 && X[0] == L &&
      X[3] == L->src && X[1] != L && X[2] != L && X[4] != L && X[1] != X[3])
  {
    const int cmax = g_chain_max < L->op->chain_max ? g_chain_max : L->op->chain_max;
    Value* xin     = L->src;
    int depth      = 1;
    if (!xin->d && xin->st)
    { // the input is itself a pending stage: extend its chain if the operands line up
      StageRec* pr = xin->st;
      if (pr->op == L->op && X[1] == pr->x && X[2] == pr->yn && X[4] == pr->fn && pr->depth < cmax) depth = pr->depth + 1;
      else materialise(sh, xin);
    }
    else materialise(sh, xin);
    if (depth == 1)
    { // a new chain: its operands must be plain data (an extended chain shares the operands of its first stage --
      // for a chain that begins the step f_n is the still-deferred L(y_n), which the launch itself will produce)
      materialise(sh, X[1]);
      materialise(sh, X[2]);
      materialise(sh, X[4]);
    }
    Value* out  = value_new(sh, false);
    StageRec* r = new StageRec();
    r->op = L->op; r->x = xin; r->p2 = X[1]; r->yn = X[2]; r->fn = X[4];
    r->x->refs++; r->p2->refs++; r->yn->refs++; r->fn->refs++;
    for (int q = 0; q < 5; q++) r->c[q] = cf[q];
    r->depth = depth;
    r->head  = false;
    out->st  = r;
    // Not launched even at full depth: ARKODE still holds z_{j-2} in tempv1 at this point and
    // drops it right after (pointer swap + N_VScale, arkode_lsrkstep.c:742-746); launching lazily,
    // when the next stage or anything else needs the data, stores two levels instead of three.
    assign(zc, out);  // (drops L, which releases its reference on xin)
    g_stats.fused_launches++;
    return;
  }

  for (int k = 0; k < nterms; k++)
    if (!X[k]->d && X[k] != L) materialise(sh, X[k]);
  if (L) materialise(sh, L->src);
  Value* out = value_new(sh, true);
  if (L)
  {
    launch_fused(sh, L, nterms, cf, X, out, store_f);
    g_stats.fused_launches++;
  }
  else
  {
    const double* vp[B200_MAX_TERMS];
    for (int k = 0; k < nterms; k++) vp[k] = X[k]->d;
    DEV(b200_lincomb(sh->ctx, nterms, cf, vp, out->d, sh->nloc));
  }
  assign(zc, out);
}

const double* mat(N_Vector v) // materialised device pointer of v
{
  Content* c = C(v);
  sync_from_host(c);
  materialise(c->sh, c->val);
  return c->val->d;
}

// ------------------------------------------------------------------ ops table
N_Vector_ID op_getvectorid(N_Vector) { return SUNDIALS_NVEC_CUSTOM; }

N_Vector op_clone(N_Vector w);

void op_destroy(N_Vector v)
{
  if (!v) return;
  Content* c = C(v);
  if (c)
  {
    if (c->val) value_release(c->sh, c->val);
    if (c->host) b200_host_free(c->host);
    Shared* sh = c->sh;
    if (--sh->refs == 0)
    {
      if (g_last_weight && g_last_weight_sh == sh)
      {
        value_release(sh, g_last_weight);
        g_last_weight = nullptr;
      }
      b200_ctx_sync(sh->ctx);
      for (double* p : sh->free_bufs) b200_free(sh->ctx, p);
      for (double* p : sh->free_halos) b200_free(sh->ctx, p);
      if (sh->wrms_host) b200_host_free(sh->wrms_host);
      else if (sh->wrms_slots) b200_free(sh->ctx, sh->wrms_slots);
      delete sh;
    }
    delete c;
  }
  N_VFreeEmpty(v);
}

void op_space(N_Vector v, sunindextype* lrw, sunindextype* liw)
{
  *lrw = C(v)->sh->nglob;
  *liw = 2;
}

sunindextype op_getlength(N_Vector v) { return C(v)->sh->nglob; }

sunrealtype* op_getarraypointer(N_Vector v)
{
  Content* c = C(v);
  if (!c->host) DEV(b200_host_alloc(c->sh->nloc, &c->host));
  if (!c->host_dirty)
  {
    if (c->val)
    {
      materialise(c->sh, c->val);
      DEV(b200_d2h(c->sh->ctx, c->host, c->val->d, c->sh->nloc));
    }
    else { memset(c->host, 0, sizeof(double) * (size_t)c->sh->nloc); }
  }
  c->host_dirty = true; // caller may write through the pointer
  return c->host;
}

void op_linearsum(sunrealtype a, N_Vector x, sunrealtype b, N_Vector y, N_Vector z)
{
  // nvector_parallel.c:424-517: every branch except a==+-b (|a| != 1) evaluates
  // (a*x) + (b*y) with exact +-1 products; those two evaluate a*(x +- y).
  const bool unit = (a == 1.0 || a == -1.0 || b == 1.0 || b == -1.0);
  Content* zc = C(z);
  Shared* sh  = zc->sh;
  if (g_lazy)
  { // nothing is launched: z becomes a pending elementwise result (see EwRec)
    sync_from_host(C(x));
    sync_from_host(C(y));
    Value* xv = C(x)->val;
    Value* yv = C(y)->val;
    if (!unit && (a == b || a == -b))
    {
      // Jv = siginv*(F(y + sig*v) - fy), arkode_ls.c:2874: F stays deferred, its argument stays pending
      const bool keepx = is_rhs(xv) && xv->op->dq && is_ew(xv->src) && xv->src->ew->kind == EWK_LIN2;
      ew_operand(sh, xv, keepx);
      ew_operand(sh, yv, false);
      assign(zc, ew_new(sh, (a == b) ? EWK_SCALESUM : EWK_SCALEDIFF, a, xv, 0.0, yv));
      return;
    }
    const bool sts_x = is_rhs(xv) || (!xv->d && xv->st), sts_y = is_rhs(yv) || (!yv->d && yv->st);
    Value* XY[2]     = {xv, yv};
    const double ab[2] = {a, b};
    if (!sts_x && !sts_y && !head_from_provenance(2, ab, XY))
    { // (deferred right-hand sides and pending stages take the stage-fusion path below)
      const bool keepy = is_ew(yv) && yv->ew->kind == EWK_SCALEDIFF && is_rhs(yv->ew->a); // z = v - gamma*Jv, :2366
      ew_operand(sh, xv, false);
      ew_operand(sh, yv, keepy);
      assign(zc, ew_new(sh, EWK_LIN2, a, xv, b, yv));
      return;
    }
  }
  if (!unit && (a == b || a == -b))
  {
    const double* xd = mat(x);
    const double* yd = mat(y);
    Value* out       = value_new(sh, true);
    DEV(b200_scale_sumdiff(sh->ctx, a, xd, yd, (a == b) ? +1 : -1, out->d, sh->nloc));
    assign(zc, out);
    return;
  }
  double cf[2]  = {a, b};
  N_Vector X[2] = {x, y};
  eval_lincomb(2, cf, X, z);
}

void op_const(sunrealtype c, N_Vector z)
{ // nothing is written: the value is the number itself until an operation needs an array (materialise).  Fixed-step
  // explicit runs set ewt = N_VConst(SUN_SMALL_REAL) every step (arkode.c:2985-2990) and only ever take a norm with it.
  Content* zc   = C(z);
  Value* out    = value_new(zc->sh, !g_lazy);
  out->is_const = true;
  out->cval     = c;
  if (out->d) DEV(b200_const(zc->sh->ctx, c, out->d, zc->sh->nloc));
  assign(zc, out);
}

#define BINARY_OP(NAME, KERNEL)                                      \
  void NAME(N_Vector x, N_Vector y, N_Vector z)                      \
  {                                                                  \
    const double* xd = mat(x);                                       \
    const double* yd = mat(y);                                       \
    Content* zc      = C(z);                                         \
    Value* out       = value_new(zc->sh, true);                      \
    DEV(KERNEL(zc->sh->ctx, xd, yd, out->d, zc->sh->nloc));          \
    assign(zc, out);                                                 \
  }
BINARY_OP(op_prod_now, b200_prod)
BINARY_OP(op_div, b200_div)

void op_prod(N_Vector x, N_Vector y, N_Vector z)
{
  if (!g_lazy) { op_prod_now(x, y, z); return; }
  sync_from_host(C(x));
  sync_from_host(C(y));
  Value* xv  = C(x)->val;
  Value* yv  = C(y)->val;
  Shared* sh = C(z)->sh;
  ew_operand(sh, xv, is_ew(xv) && xv->ew->kind == EWK_LIN2); // Ap = r.*s with r = r - alpha*Ap pending, sunlinsol_pcg.c:543-551
  ew_operand(sh, yv, false);
  assign(C(z), ew_new(sh, EWK_PROD, 0.0, xv, 0.0, yv));
}

bool unary_pending(int kind, double ca, N_Vector x, double cb, N_Vector z, bool keep_operand);

void op_scale(sunrealtype c, N_Vector x, N_Vector z)
{
  if (c == 1.0)
  { // VCopy (nvector_parallel.c:1771): share the value, deferred or not
    if (x == z) return;
    Content* xc = C(x);
    sync_from_host(xc);
    xc->val->refs++;
    assign(C(z), xc->val);
    g_stats.aliased_copies++;
    return;
  }
  if (g_lazy && C(x)->val && is_ew(C(x)->val) && C(x)->val->ew->kind == EWK_ABS &&
      unary_pending(EWK_SCALE, c, x, 0.0, z, true))
    return; // rtol*|y| of arkEwtSetSS
  double cf[1]  = {c};
  N_Vector X[1] = {x};
  eval_lincomb(1, cf, X, z);
}

#define UNARY_OP(NAME, KERNEL)                                  \
  void NAME(N_Vector x, N_Vector z)                             \
  {                                                             \
    const double* xd = mat(x);                                  \
    Content* zc      = C(z);                                    \
    Value* out       = value_new(zc->sh, true);                 \
    DEV(KERNEL(zc->sh->ctx, xd, out->d, zc->sh->nloc));         \
    assign(zc, out);                                            \
  }
UNARY_OP(op_abs_now, b200_abs)
UNARY_OP(op_inv_now, b200_inv)

// pending unary result over a plain operand (a pending operand of one of the shapes below is kept: arkEwtSetSS)
bool unary_pending(int kind, double ca, N_Vector x, double cb, N_Vector z, bool keep_operand)
{
  if (!g_lazy) return false;
  Content* xc = C(x);
  sync_from_host(xc);
  Value* xv  = xc->val;
  Shared* sh = C(z)->sh;
  if (is_rhs(xv) || (!xv->d && xv->st)) materialise(sh, xv); // a deferred right-hand side / pending stage chain goes out now
  ew_operand(sh, xv, keep_operand);
  assign(C(z), ew_new(sh, kind, ca, xv, cb, nullptr));
  return true;
}

void op_abs(N_Vector x, N_Vector z)
{
  if (!unary_pending(EWK_ABS, 0.0, x, 0.0, z, false)) op_abs_now(x, z);
}

void op_inv(N_Vector x, N_Vector z)
{
  if (g_lazy)
  {
    sync_from_host(C(x));
    double rtol = 0.0, atol = 0.0;
    Value* y = ewt_source(C(x)->val, &rtol, &atol);
    if (y)
    { // 1/(rtol*|y| + atol): one node over y; the three intermediate results are never evaluated
      Shared* sh = C(z)->sh;
      if (y->centre_sig > 0 && y->centre_sig < 96) sh->spec_ewt_sigs[y->centre_sig] = true;
      sh->ewt_seen = true;
      sh->ewt_rtol = rtol; sh->ewt_atol = atol;
      if (y->wrms_w && y->wrms_w->spec_ewt && y->wrms_w->d && y->wrms_w->e_rtol == rtol && y->wrms_w->e_atol == atol)
      { // the launch that consumed y already produced exactly this vector (launch_fused)
        y->wrms_w->refs++;
        assign(C(z), y->wrms_w);
        return;
      }
      materialise(sh, y);
      assign(C(z), ew_new(sh, EWK_EWT, rtol, y, atol, nullptr));
      return;
    }
  }
  if (!unary_pending(EWK_INV, 0.0, x, 0.0, z, false)) op_inv_now(x, z);
}

void op_addconst_now(N_Vector x, sunrealtype b, N_Vector z);
void op_addconst(N_Vector x, sunrealtype b, N_Vector z)
{
  const bool keep = g_lazy && C(x)->val && is_ew(C(x)->val) && C(x)->val->ew->kind == EWK_SCALE;
  if (!unary_pending(EWK_ADDCONST, 0.0, x, b, z, keep)) op_addconst_now(x, b, z);
}

void op_addconst_now(N_Vector x, sunrealtype b, N_Vector z)
{
  const double* xd = mat(x);
  Content* zc      = C(z);
  Value* out       = value_new(zc->sh, true);
  DEV(b200_addconst(zc->sh->ctx, xd, b, out->d, zc->sh->nloc));
  assign(zc, out);
}

// z = ca*A + cb*B pending: evaluate it inside the weighted-square-sum kernel (w: vector or constant)
double lin2_wsqr(Shared* sh, Value* z, Value* w)
{
  EwRec* e = z->ew;
  materialise(sh, e->a);
  materialise(sh, e->b);
  const bool wc = w->is_const && !w->d;
  if (!wc) materialise(sh, w);
  double* out = pool_get(sh);
  double r    = 0.0;
  DEV(b200_lin2_wsqrsum(sh->ctx, e->ca, e->a->d, e->cb, e->b->d, wc ? nullptr : w->d, wc ? w->cval : 0.0, out, sh->nloc, &r));
  z->d = out;
  ew_retire(sh, z);
  g_stats.ew_fused++;
  return r;
}

sunrealtype op_dotprod(N_Vector x, N_Vector y)
{
  Shared* sh = C(x)->sh;
  if (g_lazy)
  {
    sync_from_host(C(x));
    sync_from_host(C(y));
    Value* xv = C(x)->val;
    Value* yv = C(y)->val;
    for (int swap = 0; swap < 2; swap++)
    { // <Ap, p> with Ap = A*p pending as the difference-quotient pattern around p (sunlinsol_pcg.c:519)
      Value* a  = swap ? yv : xv;
      Value* b  = swap ? xv : yv;
      DqMatch m = match_dq(a);
      double r  = 0.0;
      if (m.ok && m.V == b && run_dq(sh, a, m, &r)) return r;
    }
    if (xv == yv && is_ew(xv) && xv->ew->kind == EWK_PROD)
    { // <r.*s, r.*s> (:543-552): a weighted square sum of r; the product itself is never needed
      Value* A = xv->ew->a;
      Value* W = xv->ew->b;
      if (is_ew(A) && A->ew->kind == EWK_LIN2) return lin2_wsqr(sh, A, W);
      materialise(sh, A);
      double r = 0.0;
      if (W->is_const && !W->d) { DEV(b200_wsqrsum_scalar(sh->ctx, A->d, W->cval, sh->nloc, &r)); }
      else
      {
        materialise(sh, W);
        DEV(b200_wsqrsum(sh->ctx, A->d, W->d, sh->nloc, &r));
      }
      return r;
    }
    for (int swap = 0; swap < 2; swap++)
    { // <r, z> with z = P^-1 r = diag.*r pending (:571-589)
      Value* zv = swap ? yv : xv;
      Value* u  = swap ? xv : yv;
      if (zv != u && is_ew(zv) && zv->ew->kind == EWK_PROD)
      {
        EwRec* e = zv->ew;
        materialise(sh, u);
        materialise(sh, e->a);
        materialise(sh, e->b);
        double* out = pool_get(sh);
        double r    = 0.0;
        DEV(b200_prod_dot(sh->ctx, e->a->d, e->b->d, u->d, out, sh->nloc, &r));
        zv->d = out;
        ew_retire(sh, zv);
        g_stats.ew_fused++;
        return r;
      }
    }
  }
  const double* xd = mat(x);
  const double* yd = mat(y);
  double r         = 0.0;
  DEV(b200_dot(sh->ctx, xd, yd, sh->nloc, &r));
  return r;
}

sunrealtype op_maxnorm(N_Vector x)
{
  const double* xd = mat(x);
  double r         = 0.0;
  DEV(b200_maxnorm(C(x)->sh->ctx, xd, C(x)->sh->nloc, &r));
  return r;
}

// bookkeeping of the speculative fused WRMS norm: remember the weight of the most recent norm, and learn which
// launch signatures a norm with that weight follows
void note_norm(Shared* sh, Value* xv, Value* wv)
{
  // learn: x came out of a fused launch issued while w was already the most recent
  // norm weight, i.e. fusing the norm into that launch would have hit -> do so next time
  if (xv->sig && g_last_weight == wv && xv->seq > g_last_weight_seq) sh->spec_sigs[xv->sig] = true;
  if (g_last_weight != wv)
  {
    if (g_last_weight) value_release(g_last_weight_sh, g_last_weight);
    g_last_weight     = wv;
    g_last_weight_sh  = sh;
    g_last_weight_seq = g_seq;
    wv->refs++;
  }
}

sunrealtype wsqrsum(N_Vector x, N_Vector w)
{
  Content* xc      = C(x);
  Shared* sh       = xc->sh;
  sync_from_host(xc);
  sync_from_host(C(w));
  if (g_lazy && is_ew(C(w)->val) && C(w)->val->ew->kind == EWK_EWT && C(w)->val->ew->a == xc->val && xc->val->d)
  { // ||y_n||_wrms with the error weights just requested for y_n (arkode.c:835): weights and norm in one pass over y_n
    Value* wv0  = C(w)->val;
    EwRec* e    = wv0->ew;
    double* out = pool_get(sh);
    double r    = 0.0;
    DEV(b200_ewt_ss_wsqrsum(sh->ctx, xc->val->d, e->ca, e->cb, out, sh->nloc, &r));
    wv0->d = out;
    ew_retire(sh, wv0);
    g_stats.ew_fused++;
    note_norm(sh, xc->val, wv0);
    return r;
  }
  if (g_lazy && is_ew(xc->val) && xc->val->ew->kind == EWK_LIN2)
  { // p = z + beta*p, then sig = 1/||p||_wrms in arkLsDQJtimes (sunlinsol_pcg.c:596, arkode_ls.c:2852)
    Value* wv0     = C(w)->val;
    const double r = lin2_wsqr(sh, xc->val, wv0);
    note_norm(sh, xc->val, wv0);
    return r;
  }
  const double* xd = mat(x);
  Value* wv        = C(w)->val;
  const bool wc    = wv->is_const && !wv->d; // constant weight that was never stored: pass the number
  const double* wd = wc ? nullptr : mat(w);
  Value* xv        = xc->val;
  double r         = 0.0;
  if (xv->wrms_w == wv && xv->wrms_slot >= 0 && sh->slot_owner[xv->wrms_slot] == xv->seq)
  { // the fused kernel that produced x already reduced sum (x*w)^2 (and no later launch has taken the slot over)
    double* slot = sh->wrms_slots + xv->wrms_slot;
    if (sh->wrms_host)
    { // the kernel stores the sum into mapped host memory: wait for it
      r = read_slot(sh, xv->wrms_slot);
    }
    else
    {
      DEV(b200_allreduce(sh->ctx, slot, 1, 0));
      DEV(b200_d2h(sh->ctx, &r, slot, 1));
    }
    xv->wrms_slot = -1; // the all-reduce is in place: do not reuse
    g_stats.wrms_fused++;
  }
  else if (wc) { DEV(b200_wsqrsum_scalar(sh->ctx, xd, wv->cval, sh->nloc, &r)); }
  else { DEV(b200_wsqrsum(sh->ctx, xd, wd, sh->nloc, &r)); }
  note_norm(sh, xv, wv);
  return r;
}

sunrealtype op_wrmsnorm(N_Vector x, N_Vector w)
{
  // nvector_parallel.c:721-730: sqrt(global sum / global length)
  return std::sqrt(wsqrsum(x, w) / (double)C(x)->sh->nglob);
}

sunrealtype op_wl2norm(N_Vector x, N_Vector w) { return std::sqrt(wsqrsum(x, w)); }

sunrealtype op_min(N_Vector x)
{
  const double* xd = mat(x);
  double r         = 0.0;
  DEV(b200_min(C(x)->sh->ctx, xd, C(x)->sh->nloc, &r));
  return r;
}

sunrealtype op_l1norm(N_Vector x)
{
  const double* xd = mat(x);
  double r         = 0.0;
  DEV(b200_l1norm(C(x)->sh->ctx, xd, C(x)->sh->nloc, &r));
  return r;
}

SUNErrCode op_linearcombination(int nvec, sunrealtype* c, N_Vector* X, N_Vector z)
{
  if (nvec < 1) return SUN_ERR_ARG_OUTOFRANGE;
  if (nvec <= B200_MAX_TERMS)
  {
    if (nvec == 1 && c[0] == 1.0) { op_scale(1.0, X[0], z); return SUN_SUCCESS; }
    eval_lincomb(nvec, c, X, z);
    return SUN_SUCCESS;
  }
  // more than 8 terms: z = first 8, then z = 1*z + next 7, ... (same left-to-right sums)
  eval_lincomb(B200_MAX_TERMS, c, X, z);
  int done = B200_MAX_TERMS;
  while (done < nvec)
  {
    int m = nvec - done;
    if (m > B200_MAX_TERMS - 1) m = B200_MAX_TERMS - 1;
    double cf[B200_MAX_TERMS];
    N_Vector Xs[B200_MAX_TERMS];
    cf[0] = 1.0;
    Xs[0] = z;
    for (int k = 0; k < m; k++) { cf[k + 1] = c[done + k]; Xs[k + 1] = X[done + k]; }
    eval_lincomb(m + 1, cf, Xs, z);
    done += m;
  }
  return SUN_SUCCESS;
}

void op_print(N_Vector v)
{
  sunrealtype* h = op_getarraypointer(v);
  C(v)->host_dirty = false;
  for (sunindextype i = 0; i < C(v)->sh->nloc; i++) printf("%.16e\n", h[i]);
}

// what an operation returns after a device failure, by return type
template <class R> R failed_result();
template <> void failed_result<void>() {}
template <> double failed_result<double>() { return std::nan(""); }
template <> int failed_result<int>() { return (int)SUN_ERR_EXT_FAIL; } // SUNErrCode
template <> double* failed_result<double*>() { return nullptr; }

// the ops-table entry for f: f behind the failure boundary described at die()
template <auto f> struct Guarded;
template <class R, class... A, R (*f)(A...)> struct Guarded<f>
{
  static R call(A... a)
  {
    if (g_failed) return failed_result<R>();
    try { return f(a...); }
    catch (const DeviceFailure&) { return failed_result<R>(); }
  }
};
#define GUARDED(f) Guarded<f>::call

void fill_ops(N_Vector v)
{
  v->ops->nvgetvectorid       = op_getvectorid;
  v->ops->nvclone             = op_clone;
  v->ops->nvcloneempty        = op_clone;
  v->ops->nvdestroy           = op_destroy;
  v->ops->nvspace             = op_space;
  v->ops->nvgetarraypointer   = GUARDED(op_getarraypointer);
  v->ops->nvgetlength         = op_getlength;
  v->ops->nvlinearsum         = GUARDED(op_linearsum);
  v->ops->nvconst             = GUARDED(op_const);
  v->ops->nvprod              = GUARDED(op_prod);
  v->ops->nvdiv               = GUARDED(op_div);
  v->ops->nvscale             = GUARDED(op_scale);
  v->ops->nvabs               = GUARDED(op_abs);
  v->ops->nvinv               = GUARDED(op_inv);
  v->ops->nvaddconst          = GUARDED(op_addconst);
  v->ops->nvdotprod           = GUARDED(op_dotprod);
  v->ops->nvmaxnorm           = GUARDED(op_maxnorm);
  v->ops->nvwrmsnorm          = GUARDED(op_wrmsnorm);
  v->ops->nvwl2norm           = GUARDED(op_wl2norm);
  v->ops->nvmin               = GUARDED(op_min);
  v->ops->nvl1norm            = GUARDED(op_l1norm);
  v->ops->nvlinearcombination = GUARDED(op_linearcombination);
  v->ops->nvprint             = op_print;
}

N_Vector op_clone(N_Vector w)
{
  N_Vector v = N_VNewEmpty(w->sunctx);
  if (!v) return nullptr;
  fill_ops(v);
  Content* c    = new Content();
  c->sh         = C(w)->sh;
  c->sh->refs++;
  c->val        = nullptr;
  c->host       = nullptr;
  c->host_dirty = false;
  v->content    = c;
  return v;
}

} // namespace

// ------------------------------------------------------------------ public API
extern "C" {

N_Vector N_VNew_B200(b200_ctx* ctx, sunindextype local_length, sunindextype global_length,
                     SUNContext sunctx)
{
  if (!ctx)
  {
    fprintf(stderr, "N_VNew_B200: a device context is required (no CPU fallback)\n");
    return nullptr;
  }
  N_Vector v = N_VNewEmpty(sunctx);
  if (!v) return nullptr;
  fill_ops(v);
  Shared* sh     = new Shared();
  sh->ctx        = ctx;
  sh->nloc       = local_length;
  sh->nglob      = global_length;
  sh->refs       = 1;
  sh->wrms_slots = nullptr;
  sh->halo_doubles = 0;
  sh->next_slot  = 0;
  for (int k = 0; k < kSlots; k++) sh->slot_owner[k] = -1;
  for (int k = 0; k < kSlots; k++) sh->slot_unread[k] = false;
  memset(sh->spec_sigs, 0, sizeof(sh->spec_sigs));
  memset(sh->spec_ewt_sigs, 0, sizeof(sh->spec_ewt_sigs));
  sh->ewt_seen = false;
  sh->ewt_rtol = sh->ewt_atol = 0.0;
  sh->wrms_host = nullptr;
  {
    int rank = 0, nranks = 1;
    b200_comm_rank(ctx, &rank, &nranks);
    const int rc = (nranks <= 1) ? b200_mapped_alloc(ctx, kSlots, &sh->wrms_host, &sh->wrms_slots)
                                 : b200_malloc(ctx, kSlots, &sh->wrms_slots);
    if (rc)
    {
      fprintf(stderr, "N_VNew_B200: %s\n", b200_last_error());
      delete sh;
      N_VFreeEmpty(v);
      return nullptr;
    }
  }
  Content* c    = new Content();
  c->sh         = sh;
  c->val        = nullptr;
  c->host       = nullptr;
  c->host_dirty = false;
  v->content    = c;
  return v;
}

b200_ctx* N_VGetContext_B200(N_Vector v) { return C(v)->sh->ctx; }
sunindextype N_VGetLocalLength_B200(N_Vector v) { return C(v)->sh->nloc; }

// (the accessors return NULL / non-zero after a device failure -- see die())
const double* N_VGetDeviceArrayPointer_B200(N_Vector v)
{
  if (g_failed) return nullptr;
  try { return mat(v); }
  catch (const DeviceFailure&) { return nullptr; }
}

double* N_VGetDeviceArrayPointerForWrite_B200(N_Vector v)
{
  if (g_failed) return nullptr;
  try
  {
    Content* c = C(v);
    Value* out = value_new(c->sh, true);
    assign(c, out);
    return out->d;
  }
  catch (const DeviceFailure&) { return nullptr; }
}

int N_VCopyFromHost_B200(N_Vector v, const double* host)
{
  double* d = N_VGetDeviceArrayPointerForWrite_B200(v);
  if (!d) return -1;
  return b200_h2d(C(v)->sh->ctx, d, host, C(v)->sh->nloc);
}

int N_VCopyToHost_B200(N_Vector v, double* host)
{
  const double* d = N_VGetDeviceArrayPointer_B200(v);
  if (!d) return -1;
  return b200_d2h(C(v)->sh->ctx, host, d, C(v)->sh->nloc);
}
int N_VDeviceFailed_B200(void) { return g_failed ? 1 : 0; }

int N_VSetDeferredRhs_B200(N_Vector f, const B200RhsOp* op, N_Vector y)
{
  Content* fc = C(f);
  Content* yc = C(y);
  if (fc->sh != yc->sh && fc->sh->nloc != yc->sh->nloc) return -1;
  if (g_failed) return -1;
  try
  {
    sync_from_host(yc);
    Value* L = value_new(fc->sh, false);
    L->op    = op;
    L->src   = yc->val;
    yc->val->refs++;
    assign(fc, L);
    if (!g_lazy) materialise(fc->sh, L);
  }
  catch (const DeviceFailure&) { return -1; }
  return 0;
}

int N_VIsDeferred_B200(N_Vector v)
{
  const Value* val = C(v)->val;
  return (val && !val->d && (val->op || val->st)) ? 1 : 0;
}
void N_VSetLazyFusion_B200(int on) { g_lazy = (on != 0); }
void N_VSetStageChain_B200(int depth)
{
  if (depth < 1) depth = 1;
  if (depth > B200_MAX_CHAIN) depth = B200_MAX_CHAIN;
  g_chain_max = depth;
}
int N_VGetStageChain_B200(void) { return g_chain_max; }
void N_VGetStats_B200(B200VecStats* s) { *s = g_stats; }

static int g_live_sessions = 0;
static int g_session_set[4] = {-1, -1, -1, -1}; // lazy, chain depth, fma, variant of the sessions that are alive

int N_VAcquireSettings_B200(int lazy, int chain_depth, int fma_arithmetic, int chain_variant)
{
  const int want[4] = {lazy, chain_depth, fma_arithmetic, chain_variant};
  if (g_live_sessions > 0)
    for (int k = 0; k < 4; k++)
      if (want[k] >= 0 && g_session_set[k] >= 0 && want[k] != g_session_set[k])
      {
        static const char* what[4] = {"lazy fusion", "stage chain depth", "arithmetic (exact / fma)", "chain kernel variant"};
        fprintf(stderr, "nvector_b200: another session is alive with %s = %d; this one asks for %d (process-wide setting)\n",
                what[k], g_session_set[k], want[k]);
        return -1;
      }
  if (g_live_sessions == 0)
    for (int k = 0; k < 4; k++) g_session_set[k] = -1;
  for (int k = 0; k < 4; k++)
    if (want[k] >= 0) g_session_set[k] = want[k];
  if (lazy >= 0) N_VSetLazyFusion_B200(lazy);
  if (chain_depth >= 1) N_VSetStageChain_B200(chain_depth);
  if (fma_arithmetic >= 0) b200_set_contract(fma_arithmetic);
  if (chain_variant >= 0) b200_set_chain_variant(chain_variant);
  g_live_sessions++;
  return 0;
}

void N_VReleaseSettings_B200(void)
{
  if (g_live_sessions > 0) g_live_sessions--;
}

} // extern "C"
