/* b200_sts.h -- C-ABI of libb200sts.so: the B200 (sm_100a) device layer of the
 * explicit super-time-stepping hot path of ceda-demonstrations' diffusion_2D / adr.
 *
 * Plain C: raw DEVICE pointers (double*), sizes (int64_t), scalars, an opaque
 * context handle and a CUDA stream passed as void*.  No torch types, no CUDA
 * headers, no SUNDIALS headers -- host code (C++, Python/ctypes, cgo ...) binds
 * these directly.  Every function returns 0 on success and a non-zero
 * cudaError_t / ncclResult_t-derived code otherwise (b200_last_error() gives text).
 * There is NO CPU fallback: without a CUDA device every entry fails loudly.
 *
 * Each entry cites the reference interface it stands in for
 * (paths relative to /root/reference; SUN = deps/sundials).
 *
 * Arithmetic contract: IEEE-754 binary64, round-to-nearest, NO fused
 * multiply-add, and the reference's own association order, so that elementwise
 * results are bit-identical to the reference's CPU build (gcc -O2, baseline
 * x86-64).  Reductions are deterministic (fixed tree) but not sequential.
 * The one opt-in exception is b200_set_contract(1) (chain kernels only, see there).
 */
#ifndef B200_STS_H
#define B200_STS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_MAX_TERMS 8

typedef struct b200_ctx b200_ctx; /* opaque: device, stream, scratch, NCCL comm */

/* ------------------------------------------------------------------ context */
/* Create a context on CUDA device `device` (one process per GPU).  `stream` may
   be NULL (the context creates its own non-blocking stream) or an existing
   cudaStream_t (e.g. torch's current stream) so that callers can bracket work
   with their own events. */
int b200_ctx_create(int device, void* stream, b200_ctx** out);
int b200_ctx_destroy(b200_ctx* ctx);
void* b200_ctx_stream(b200_ctx* ctx);
int b200_ctx_sync(b200_ctx* ctx);
const char* b200_last_error(void);
/* number of kernels this library has launched since load (bench "gpu_launches") */
uint64_t b200_launch_count(void);
/* modelled ("algorithmic") bytes of all launches since load: 8 B x vector length x the
   number of full vectors each kernel reads and writes (halo re-reads, coefficient tables and
   scalars not counted).  bench.py divides its growth over the timed region by the device
   time to get the achieved-bandwidth side of the roofline. */
uint64_t b200_algorithmic_bytes(void);

/* device / pinned-host memory (stands in for N_VNew_Parallel's malloc,
   SUN/src/nvector/parallel/nvector_parallel.c:189-221) */
int b200_malloc(b200_ctx* ctx, int64_t n_doubles, double** dptr);
int b200_free(b200_ctx* ctx, double* dptr);
int b200_host_alloc(int64_t n_doubles, double** hptr);
int b200_host_free(double* hptr);
/* n doubles of pinned host memory mapped into the device's address space: *dptr is what kernels write through
   (e.g. the wrms_result of b200_stage_extras), *hptr what the host reads after b200_ctx_sync.  Free with b200_host_free. */
int b200_mapped_alloc(b200_ctx* ctx, int64_t n, double** hptr, double** dptr);
int b200_h2d(b200_ctx* ctx, double* dst_dev, const double* src_host, int64_t n);
/* n <= 64: read back through mapped pinned memory written by a 1-block kernel (no DMA copy, so a
   scalar never queues behind a bulk transfer on the copy engine); larger n: cudaMemcpyAsync + sync */
int b200_d2h(b200_ctx* ctx, double* dst_host, const double* src_dev, int64_t n);

/* ----------------------------------------- pipelined host <-> device staging */
/* For a stream of INDEPENDENT states (ensemble members, sweeps, bench.py's end-to-end leg): the
   upload of state seq+1 and the download of result seq-1 overlap the integration of state seq.
   Two input and two output staging buffers of n doubles in HBM, one H2D and one D2H copy stream;
   slot = seq % 2.  Host buffers must be pinned (b200_host_alloc) for the copies to be asynchronous.
   Call order per state: upload(seq) [may run ahead by two], take(seq, vec), ...work on the
   context's stream..., put(seq, vec, host); b200_pipe_drain before reading the host results.
   (The reference keeps its state in host memory, diffusion_2D/main.cpp:176-190: this replaces the
   plain b200_h2d / b200_d2h round trip when there is more than one state to process.) */
typedef struct b200_pipe b200_pipe;
int b200_pipe_create(b200_ctx* ctx, int64_t n_doubles, b200_pipe** out);
int b200_pipe_destroy(b200_pipe* p);
int b200_pipe_upload(b200_pipe* p, int64_t seq, const double* host_src);
int b200_pipe_take(b200_pipe* p, int64_t seq, double* dst_dev);
int b200_pipe_put(b200_pipe* p, int64_t seq, const double* src_dev, double* host_dst);
int b200_pipe_drain(b200_pipe* p);

/* ------------------------------------------------- elementwise vector kernels */
/* z = sum_k c[k]*v[k] evaluated left to right:  acc = c0*v0; acc = acc + ck*vk.
   This is SUNDIALS' generic N_VLinearCombination fallback
   (SUN/src/sundials/sundials_nvector.c:557-565 -> N_VScale + Vaxpy,
   SUN/src/nvector/parallel/nvector_parallel.c:568-592,1909-1926) and, with
   nterms<=2, every N_VLinearSum / N_VScale case except a==+-b
   (nvector_parallel.c:424-517).  z may alias any v[k].  1 <= nterms <= 8. */
int b200_lincomb(b200_ctx* ctx, int nterms, const double* c,
                 const double* const* v, double* z, int64_t n);
/* z = a*(x + y)  (sign=+1, VScaleSum nvector_parallel.c:1839) or
   z = a*(x - y)  (sign=-1, VScaleDiff :1855) */
int b200_scale_sumdiff(b200_ctx* ctx, double a, const double* x, const double* y,
                       int sign, double* z, int64_t n);
int b200_const(b200_ctx* ctx, double c, double* z, int64_t n);  /* N_VConst :519 */
int b200_prod(b200_ctx* ctx, const double* x, const double* y, double* z, int64_t n); /* N_VProd :534 */
int b200_div(b200_ctx* ctx, const double* x, const double* y, double* z, int64_t n);  /* N_VDiv :551 */
int b200_abs(b200_ctx* ctx, const double* x, double* z, int64_t n);   /* N_VAbs :594 */
int b200_inv(b200_ctx* ctx, const double* x, double* z, int64_t n);   /* N_VInv :610 */
int b200_addconst(b200_ctx* ctx, const double* x, double b, double* z, int64_t n); /* N_VAddConst :626 */
/* ewt = 1 / (rtol*|y| + atol): the four-op sequence of arkEwtSetSS
   (SUN/src/arkode/arkode.c:2932-2944) in one pass, same rounding sequence. */
int b200_ewt_ss(b200_ctx* ctx, const double* y, double rtol, double atol,
                double* ewt, int64_t n);

/* ------------------------------------------------------------- reductions */
/* Each writes the LOCAL result to *result (host) after a stream sync; with an
   initialised communicator (b200_comm_init) the value is all-reduced over the
   ranks first (SUM / MAX / MIN), standing in for the MPI_Allreduce in
   nvector_parallel.c:658-730. */
int b200_dot(b200_ctx* ctx, const double* x, const double* y, int64_t n, double* result);      /* N_VDotProd :658 */
int b200_wsqrsum(b200_ctx* ctx, const double* x, const double* w, int64_t n, double* result);  /* N_VWSqrSumLocal :700 + Allreduce :728 */
/* the same with ONE weight w for every entry (a constant-valued weight vector, e.g. the ewt = N_VConst(SUN_SMALL_REAL)
   of fixed-step explicit runs, SUN/src/arkode/arkode.c:2985-2990): w is never read from memory */
int b200_wsqrsum_scalar(b200_ctx* ctx, const double* x, double w, int64_t n, double* result);
/* B200_TRACE_LAUNCHES=1 in the environment: per-kernel launch counts and wall times (each launch bracketed by stream
   synchronisations) and the host time between launches, printed to stderr by this call and at b200_ctx_destroy. */
void b200_trace_report(void);

/* Fused forms of the implicit path's vector work (SUN/src/sunlinsol/pcg/sunlinsol_pcg.c:499-601).  Each stores z exactly
   as the separate N_VLinearSum / N_VProd would and returns the reduction N_VDotProd / N_VWrmsNorm takes of it next:
     b200_lin2_wsqrsum  z = ca*a + cb*b ; result = sum_i (z_i w_i)^2   (w == NULL: the scalar weight wscalar)
                        -- r = r - alpha*Ap with rho = <r.*s, r.*s> (:543-559), p = z + beta*p with the WRMS norm of
                           arkLsDQJtimes (arkode_ls.c:2852)
     b200_prod_dot      z = a .* b ; result = sum_i c_i z_i              -- z = P^-1 r (Jacobi), rz = <r, z> (:571-589) */
int b200_lin2_wsqrsum(b200_ctx* ctx, double ca, const double* a, double cb, const double* b, const double* w,
                      double wscalar, double* z, int64_t n, double* result);
int b200_prod_dot(b200_ctx* ctx, const double* a, const double* b, const double* c, double* z, int64_t n, double* result);
/* ewt = 1/(rtol*|y| + atol) as b200_ewt_ss, and in the same pass result = sum_i (y_i ewt_i)^2: the new error weights
   (arkEwtSetSS, SUN/src/arkode/arkode.c:2932-2944) together with the norm ARKODE takes of y_n with them at the top of
   the next step (arkode.c:835) */
int b200_ewt_ss_wsqrsum(b200_ctx* ctx, const double* y, double rtol, double atol, double* ewt, int64_t n, double* result);
int b200_maxnorm(b200_ctx* ctx, const double* x, int64_t n, double* result);                   /* N_VMaxNorm :689 */
int b200_min(b200_ctx* ctx, const double* x, int64_t n, double* result);                       /* N_VMin :780 */
int b200_l1norm(b200_ctx* ctx, const double* x, int64_t n, double* result);                    /* N_VL1Norm :815 */

/* ------------------------------------------------- diffusion_2D RHS stencil */
/* Local sub-domain of the anisotropic / inhomogeneous 5-point operator
   f = d/dx(Dx du/dx) + d/dy(Dy du/dy) of diffusion_2D/diffusion.cpp:9-209.
   Fields are row-major nx*ny doubles, x fastest (IDX, diffusion_2D.hpp:55).
   The four face-coefficient tables hold exactly the values diffusion.cpp:36-46
   computes per cell -- Diffusion_Coeff_X(xl+(is+i-+0.5)dx)/(dx*dx) etc. -- and
   are computed ON THE HOST (libm sin) and uploaded, so coefficients are
   bit-identical to the reference.  halo_* are the neighbour columns / rows
   received from the W/E/S/N ranks (the Wrecv..Nrecv buffers of
   diffusion_2D.hpp:156-159); a NULL halo pointer means "this direction is
   periodic onto my own field" (one rank in that direction), and the kernel
   wraps the index instead of reading a buffer. */
typedef struct b200_stencil_geom
{
  int64_t nx, ny;          /* local extents nx_loc, ny_loc */
  const double* cxw;       /* [nx] Dx_w(i) */
  const double* cxe;       /* [nx] Dx_e(i) */
  const double* cys;       /* [ny] Dy_s(j) */
  const double* cyn;       /* [ny] Dy_n(j) */
  const double* halo_w;    /* [ny] or NULL */
  const double* halo_e;    /* [ny] or NULL */
  const double* halo_s;    /* [nx] or NULL */
  const double* halo_n;    /* [nx] or NULL */
  /* Optional promise (0 = none): every entry of cxw / cxe / cys / cyn (margins of the halo flavour
     included) is bitwise equal to u_cxw / u_cxe / u_cys / u_cyn -- the homogeneous problem,
     Diffusion_Coeff_X/Y = kx / ky (diffusion_2D.cpp:887-897).  b200_stencil_chain[_halo] then takes the
     coefficients from kernel parameters instead of loading tables; results are bit-identical. */
  int uniform;
  double u_cxw, u_cxe, u_cys, u_cyn;
} b200_stencil_geom;

/* term sources for b200_stencil_lincomb */
#define B200_SRC_VECTOR  0 /* v[k] is a device vector */
#define B200_SRC_CENTRE  1 /* term k is the stencil input x itself (no second load) */
#define B200_SRC_STENCIL 2 /* term k is L(x), computed on the fly */

/* What to do besides z (all optional, NULL = off) */
typedef struct b200_stage_extras
{
  double* f_out;        /* also store L(x) (the plain ARKRhsFn result) */
  double* send_w;       /* [ny] pack z's west column   (buffers.cpp:28-30)  */
  double* send_e;       /* [ny] pack z's east column   (buffers.cpp:32-34)  */
  double* send_s;       /* [nx] pack z's south row     (buffers.cpp:36-38)  */
  double* send_n;       /* [nx] pack z's north row     (buffers.cpp:40-42)  */
  const double* wrms_w; /* fuse sum_i (z_i*w_i)^2 (N_VWSqrSumLocal) ...      */
  double* wrms_result;  /* ... into this DEVICE double                       */
  /* With wrms_w: also the error weights OF THE STENCIL INPUT x, ewt_i = 1/(rtol*|x_i| + atol) -- the N_VAbs, N_VScale,
     N_VAddConst, N_VInv sequence of arkEwtSetSS (SUN/src/arkode/arkode.c:2932-2944), same roundings -- stored to ewt_out,
     and sum_i (x_i*ewt_i)^2 into the device double ewt_result: what ARKODE computes next (arkode.c:835, :2985) if x is
     accepted as y_{n+1}, i.e. when this launch is the closing stage of an adaptive STS step. */
  double* ewt_out;
  double ewt_rtol, ewt_atol;
  double* ewt_result;
} b200_stage_extras;

/* The fused STS stage:  z = sum_k c[k]*T_k, left to right, where T_k is a
   device vector, the stencil input x, or L(x) (exactly one term should be
   B200_SRC_STENCIL).  One HBM pass realises the reference's
   `fe(x -> F)` + `N_VLinearCombination(5, ...)` of
   SUN/src/arkode/arkode_lsrkstep.c:686-717 (RKC), :985-1020 (RKL), the closing
   `fe` + embedding LC4 of :768-789, the SSP `fe` + `N_VLinearSum(1,y,c,F,y)`
   of :1213-1231, and plain f = L(x) (nterms=1, c=1).
   z must NOT alias x (neighbours are read while z is written); it may alias v[k].
   region: 0 = whole sub-domain, 1 = boundary ring only, 2 = interior only
   (ring first / interior after lets the halo exchange overlap the interior). */
int b200_stencil_lincomb(b200_ctx* ctx, const b200_stencil_geom* g,
                         const double* x, int nterms, const double* c,
                         const int* src, const double* const* v, double* z,
                         const b200_stage_extras* extras, int region);

/* Temporal blocking: `nstages` (2..B200_MAX_CHAIN) consecutive
   RKC/RKL stages of SUN/src/arkode/arkode_lsrkstep.c:674-750 (RKC) / :960-1050 (RKL)
   in one pass.  Stage l (1-based) computes
       z_l = c[l][0]*L(z_{l-1}) + c[l][1]*z_{l-2} + c[l][2]*yn + c[l][3]*z_{l-1} + c[l][4]*fn
   with z_0 = x and z_{-1} = prev2, evaluated left to right exactly as
   b200_stencil_lincomb would, so results are bit-identical to nstages separate
   launches.  coeffs is [nstages][5] row-major; z_out[l] receives z_{l+1} or may be
   NULL for an intermediate stage nobody reads (the last stage must be stored).
   One periodic rank only (all halo_* NULL), nx even >= 128, ny >= 16. */
#define B200_MAX_CHAIN 6
int b200_stencil_chain(b200_ctx* ctx, const b200_stencil_geom* g, int nstages,
                       const double* x, const double* prev2, const double* yn,
                       const double* fn, const double* coeffs, double* const* z_out);
/* The same on a rank of a 2-D block decomposition: rows / columns outside the local
   sub-domain come from per-operand deep-halo buffers filled by b200_deep_halo_exchange
   (layout below), halos[] = { of x, of prev2, of yn, of fn }.  The coefficient tables of
   `g` must then be valid for indices -16 .. n+15 (global periodic index, i.e. what the
   neighbouring ranks use for those cells).  halo_rows >= nstages, halo_cols even >=
   2*ceil(nstages/2). */
int b200_stencil_chain_halo(b200_ctx* ctx, const b200_stencil_geom* g, int nstages,
                            const double* x, const double* prev2, const double* yn,
                            const double* fn, const double* coeffs, double* const* z_out,
                            const double* const* halos, int halo_rows, int halo_cols);
/* The chain that BEGINS an STS step (stage 1 folded in): z_1 = x + c_1 L(x) with x = y_n (N_VLinearSum(ONE, yn, h*mus, fn,
   ..), arkode_lsrkstep.c:640 / :930; coeffs[0][0] = c_1, the other four entries of row 0 are ignored), stages 2..nstages
   as above with z_0 = y_n = x and f_n = L(x).  f_n is stored to f_out (it is the vector ARKODE keeps as fn; NULL when
   the caller holds it already -- an adaptive step whose f_n came out of the previous step's closing stage: the kernel
   recomputes the same bits rather than stream them) and the later stages take it from the kernel's own ring: y_n is
   streamed once, there is no z_{-1} and no f_n stream.
   halo_x: deep halo of x on a rank of a decomposition (NULL on one periodic rank). */
int b200_stencil_chain_head(b200_ctx* ctx, const b200_stencil_geom* g, int nstages, const double* x,
                            const double* coeffs, double* const* z_out, double* f_out,
                            const double* halo_x, int halo_rows, int halo_cols);
/* Load (and opt into their shared memory) all instantiations of the chain kernel a problem with these properties can
   reach: depths 2..B200_MAX_CHAIN, with and without the stage-1 head, for the current arithmetic (b200_set_contract).
   CUDA loads kernels lazily at their first launch; an adaptive run may meet a new depth in any step. */
int b200_stencil_chain_preload(b200_ctx* ctx, int halo, int uniform);
/* Deep halo of one nx*ny field, `rows` deep in y and `cols` deep in x, corners included:
     [ S: rows x nx | N: rows x nx | W: (ny+2 rows) x cols | E: (ny+2 rows) x cols ]
   S = rows -rows..-1, N = rows ny..ny+rows-1, W / E = columns -cols..-1 / nx..nx+cols-1
   of rows -rows..ny+rows-1.  b200_deep_halo_doubles gives its size. */
int64_t b200_deep_halo_doubles(int64_t nx, int64_t ny, int rows, int cols);
/* Fill the deep halos of `nfields` (<= 4) fields: two NCCL phases on the compute stream
   (S/N blocks, then W/E strips that carry the corners), replacing the one-deep
   start_exchange/end_exchange of diffusion_2D.cpp:400-584 for temporally blocked
   launches.  peers = ranks of the W,E,S,N neighbours; x_split / y_split = 0 means one
   rank in that direction (the periodic neighbour is this rank: local copies). */
int b200_deep_halo_exchange(b200_ctx* ctx, const int peers[4], int x_split, int y_split,
                            int64_t nx, int64_t ny, int rows, int cols, int nfields,
                            const double* const* fields, double* const* halos);
/* ---- the same halos WITHOUT NCCL: peer-mapped slots written by the neighbours over NVLink --------------------
   b200_peer_halo_create allocates a ring of `nslots` deep-halo slots (each sized for a block nx x ny_max) on this
   rank, publishes it to the other ranks (CUDA IPC handles all-gathered over the communicator of b200_comm_init) and
   maps the rings of the eight neighbours nbr[] = { W, E, S, N, SW, SE, NW, NE } (a rank may appear several times, and
   may be this rank itself: periodic wrap with one rank in a direction).  ny_south / ny_north are the heights of the
   blocks below / above this one (the strips of the diagonal neighbours are laid out for THEIR height).
   b200_peer_halo_exchange(fields, slots): ONE kernel per rank -- it stores the edge bands and corners of each field
   into slot slots[f] of the neighbours (same slot number on every rank: ranks call the allocator in lockstep), raises
   one flag per neighbour and waits for the neighbours' flags; after it, slots[f] on this rank holds the deep halo of
   fields[f] in the layout above.  It replaces b200_deep_halo_exchange (two dependent NCCL groups plus pack kernels)
   and with it MPI_Isend / Irecv / Waitall of diffusion_2D.cpp:400-584 for temporally blocked launches. */
typedef struct b200_peer_halo b200_peer_halo;
int b200_peer_halo_create(b200_ctx* ctx, const int nbr[8], int64_t nx, int64_t ny, int64_t ny_south, int64_t ny_north,
                          int64_t ny_max, int rows, int cols, int nslots, b200_peer_halo** out);
int b200_peer_halo_destroy(b200_peer_halo* ph);
double* b200_peer_halo_slot_alloc(b200_peer_halo* ph);            /* NULL when the ring is exhausted */
int b200_peer_halo_slot_free(b200_peer_halo* ph, double* slot);
int b200_peer_halo_exchange(b200_peer_halo* ph, int nfields, const double* const* fields, double* const* slots);
int b200_peer_halo_stats(const b200_peer_halo* ph, uint64_t* exchanges, uint64_t* doubles_pushed);
/* The matrix-free linear operator of the implicit path in ONE stencil pass (one periodic rank, even nx):
     outer = 1:  z = ca*v + cb*( siginv*( L(sigma*v + y) - fy ) )   = arkLsATimes o arkLsDQJtimes with ca = 1, cb = -gamma
                 (SUN/src/arkode/arkode_ls.c:2316-2372, :2839-2877)
     outer = 0:  z = siginv*( L(sigma*v + y) - fy )                  = arkLsDQJtimes / lsrkStep_DQJtimes
                 (SUN/src/arkode/arkode_lsrkstep.c:2408-2431)
   element by element the instruction sequence of the separate N_VLinearSum / RHS / N_VLinearSum calls.  With
   dot_result != NULL also returns sum_i z_i v_i (PCG's <Ap, p>, sunlinsol_pcg.c:519). */
int b200_stencil_dq(b200_ctx* ctx, const b200_stencil_geom* g, const double* v, const double* y, const double* fy,
                    double sigma, double siginv, int outer, double ca, double cb, double* z, double* dot_result);
/* rows of output each thread block of the chain kernel marches over; 0 (default) = automatic: 256
   where that leaves at least 4 waves of blocks, else 128, else 64, else 32 (8 or 16 on grids so small that every
   block is resident at once).  Results do not depend on it. */
int b200_set_chain_rows(int rows);
/* the automatic choice for an nx x ny block and a chain of nstages stages on a device with sm_count multiprocessors
   (0: the device of ctx, which may then not be NULL); launches nothing.  -1: bad arguments. */
int b200_chain_rows_query(b200_ctx* ctx, int64_t nx, int64_t ny, int nstages, int sm_count);
/* Which kernel runs a chain: 0 (default) = k_chain_march, two cells per thread; 1 = k_chain_quad,
   four cells per thread (two 64-cell halves per warp window, 120 of 128 cells useful at depth 4).
   Same shape requirements, bit-identical results; at depth 4 on 16384^2 both sit at 85-90 % of the
   measured HBM copy bandwidth on the chain's own 48 B per cell (profiles/).  The environment
   variable B200_CHAIN_VARIANT sets the initial value. */
int b200_set_chain_variant(int variant);
int b200_get_chain_variant(void);
/* SPLIT flavour of k_chain_march (depth 4, exact arithmetic): the upper half of the levels runs one row late and
   first in a row step, so the two halves are independent instruction streams (more work in flight for the FP64 pipe);
   bit-identical results.  B200_CHAIN_SPLIT sets the initial value. */
int b200_set_chain_split(int on);
int b200_get_chain_split(void);
/* How k_chain_march (depth 4, exact arithmetic) fills its operand ring in shared memory.  0: one 16-byte cp.async per
   thread, operand and row; 1 / 2: bulk asynchronous copies (cp.async.bulk, the TMA unit's 1-D path: one 512-byte copy
   per warp, operand and row, issued by one lane from warp-uniform pointers and completed on an mbarrier) with 3 / 4
   rows in flight.  Default 2 (111 KB of shared memory per block; measured in DESIGN.md 4.1b); bit-identical results.
   B200_CHAIN_BULK sets the initial value; a negative argument returns to it. */
int b200_set_chain_bulk(int on);
int b200_get_chain_bulk(void);
/* Rows of operands the plain flavour of k_chain_march (depth 4, exact arithmetic) keeps in flight: 3 or 4 (other
   values: back to the initial one, B200_CHAIN_PF or the default).  Results do not depend on it. */
int b200_set_chain_pf(int pf);
int b200_get_chain_pf(void);
/* 1 (default): honour b200_stencil_geom.uniform; 0: always load the coefficient tables (A/B tests) */
int b200_set_chain_uniform(int on);
/* name of the kernel the most recent chain launch used ("k_chain_quad" / "k_chain_march", "" if none) */
const char* b200_last_chain_kernel(void);
/* Arithmetic of the chain kernels.  0 (default) = the arithmetic contract stated at the top of this
   header: every multiply and add rounded separately, bit-identical to the reference's baseline
   x86-64 build.  1 = each c + a*b of diffusion.cpp:48-53 / sundials_nvector.c:557-565 is contracted
   to one fused multiply-add, i.e. what gcc's default -ffp-contract=fast makes of the reference on an
   FMA-baseline ISA (aarch64, ppc64le, x86-64-v3): results then agree with the baseline build to
   rounding (~1e-15 relative per stage; the parity bar of 1e-10 on the final state holds), not bit
   for bit, and the FP64 pipe has 22 instead of 38 instructions per two cell-updates to issue. */
int b200_set_contract(int on);
int b200_get_contract(void);

/* Diagnostics.  B200_TRACE_LAUNCHES=1 serialises every launch and times it per kernel; B200_HOST_PROFILE=1 only
   accumulates the host time spent inside launch calls and in waits for a reduction result; both are printed by
   b200_trace_report (declared above; also called when a context is destroyed).  b200_host_profile_wait adds a wait that
   happened in a caller. */
int b200_host_profile_on(void);
void b200_host_profile_wait(double ms);

/* Tuning knob: rows of the sub-domain each thread block of the fused kernel marches over
   (default 8).  Results do not depend on it. */
int b200_set_rows_per_block(int rows);

/* Standalone halo pack of a materialised field (buffers.cpp:20-43). */
int b200_pack_halo(b200_ctx* ctx, const double* u, int64_t nx, int64_t ny,
                   double* send_w, double* send_e, double* send_s, double* send_n);

/* Jacobi preconditioner setup: diag = 1/(1 - gamma*d_ij),
   d_ij = -((pxw[i]+pxe[i]) + (pys[j]+pyn[j])), tables host-computed with the
   reference's (different) coordinates (preconditioner_jacobi.cpp:9-46). */
int b200_jacobi_setup(b200_ctx* ctx, int64_t nx, int64_t ny, const double* pxw,
                      const double* pxe, const double* pys, const double* pyn,
                      double gamma, double* diag);

/* ------------------------------------------------ adr 2-D Brusselator kernels */
/* Periodic nx*ny grid, two interleaved species y[2*(i+j*nx)+s]
   (adr/advection_diffusion_reaction_2d.hpp:44-45). */
typedef struct b200_adr_params
{
  int64_t nx, ny;
  double dx, dy;
  double cux, cuy, cvx, cvy; /* advection speeds        (…2d.hpp UserData) */
  double d;                  /* diffusion coefficient                      */
  double A, B;               /* Brusselator parameters                     */
} b200_adr_params;
/* mode bits: 1 = advection (f_advection …2d.cpp:1406-1445),
              2 = diffusion (f_diffusion :1448-1491),
              4 = reaction  (f_reaction  :1494-1520);
   3/5/6/7 realise the composite callbacks (f_adv_react :1602-1619 etc.) in the
   reference's summation order.  f must not alias y. */
int b200_adr_rhs(b200_ctx* ctx, const b200_adr_params* p, int mode,
                 const double* y, double* f);
/* fused  z = sum_k c[k]*T_k  with T_k in { v[k], y, F_mode(y) } per src[k], F_mode the
   (composite) callback selected by the mode bits above; f_out != NULL also stores
   F_mode(y).  Realises the RHS call + the N_VLinearCombination / N_VLinearSum that
   consumes it (LSRKStep stages of the diffusion partition, ARKStep/ERKStep stages of
   the advection-reaction partition) in one pass. */
int b200_adr_lincomb(b200_ctx* ctx, const b200_adr_params* p, int mode, const double* y,
                     int nterms, const double* c, const int* src,
                     const double* const* v, double* z, double* f_out);
/* mode = 2 (f_diffusion) shorthand */
int b200_adr_diffusion_lincomb(b200_ctx* ctx, const b200_adr_params* p,
                               const double* y, int nterms, const double* c,
                               const int* src, const double* const* v, double* z,
                               double* f_out);

/* Temporal blocking of the adr diffusion partition: `nstages` (2..B200_MAX_CHAIN) consecutive RKC / RKL
   stages  z_l = c[l][0] F(z_{l-1}) + c[l][1] z_{l-2} + c[l][2] yn + c[l][3] z_{l-1} + c[l][4] fn  with
   F = f_diffusion (adr/advection_diffusion_reaction_2d.cpp:1448-1491), z_0 = x, z_{-1} = prev2, in one
   pass; bit-identical to nstages b200_adr_lincomb launches (mode 2, sources {2,0,0,1,0}).  coeffs is
   [nstages][5]; z_out[l] may be NULL for a stage nobody reads (the last one must be stored).
   nx >= 64, ny >= 16.  (k_adr_chain, csrc/adr_chain.cuh; verified on the host emulator, opt-in in the
   adr driver through --sts_chain K until measured on the GPU.) */
int b200_adr_chain(b200_ctx* ctx, const b200_adr_params* p, int nstages, const double* x,
                   const double* prev2, const double* yn, const double* fn, const double* coeffs,
                   double* const* z_out);

/* ---- adr with IMPLICIT reaction (--implicit-reaction; SetupStrang / SetupExtSTS, ...2d.cpp:1207-1246, :820-854) ----
   The reference keeps J_reaction (...2d.cpp:1523-1551) in a SUNBandMatrix(neq, 2, 2) and solves the Newton systems
   with SUNLinSol_Band.  With the species interleaved the matrix is block diagonal (one 2 x 2 block per grid point), and
   the banded LU with partial pivoting (SUN/src/sundials/sundials_band.c) never leaves a block; these entry points run
   its operations per block, in its order and with its roundings.  Matrix storage: 4 doubles per grid point
   { dU/du, dV/du, dU/dv, dV/dv }; piv: one double per grid point (1 = rows swapped). */
int b200_adr_jac_reaction(b200_ctx* ctx, const b200_adr_params* p, const double* y, double* J);
/* A = c*A + I (SUNMatScaleAddI_Band, SUN/src/sunmatrix/band/sunmatrix_band.c) */
int b200_blk2_scale_add_i(b200_ctx* ctx, double c, double* A, int64_t npts);
/* in-place LU (bandGBTRF); *info = 0 or the 1-based column of the first zero pivot.  Synchronises the stream. */
int b200_blk2_factor(b200_ctx* ctx, double* A, double* piv, int64_t npts, long long* info);
/* x = A^-1 b from the factors (bandGBTRS); x may alias b */
int b200_blk2_solve(b200_ctx* ctx, const double* A, const double* piv, const double* b, double* x, int64_t npts);

/* --------------------------------------------------- multi-GPU (NCCL, NVLink) */
/* One process per GPU.  rank 0 calls b200_comm_unique_id and distributes the 128
   bytes out of band (torch.distributed / a file); every rank then calls
   b200_comm_init.  Replaces MPI_Init + MPI_Cart_create of
   diffusion_2D/diffusion_2D.cpp:223-394. */
int b200_comm_unique_id(unsigned char id[128]);
int b200_comm_init(b200_ctx* ctx, int rank, int nranks, const unsigned char id[128]);
int b200_comm_rank(b200_ctx* ctx, int* rank, int* nranks);
/* Halo exchange: one grouped ncclSend/ncclRecv per neighbour and direction on the
   context's communication stream (start_exchange/end_exchange,
   diffusion_2D.cpp:400-584).  peers: ranks of the W,E,S,N neighbours.  The call
   makes the comm stream wait for everything enqueued so far on the compute
   stream, and b200_halo_wait makes the compute stream wait for the exchange. */
int b200_halo_exchange(b200_ctx* ctx, const int peers[4], const double* send_w,
                       const double* send_e, const double* send_s,
                       const double* send_n, double* recv_w, double* recv_e,
                       double* recv_s, double* recv_n, int64_t nx, int64_t ny);
int b200_halo_wait(b200_ctx* ctx);
/* all-reduce n doubles in place on the device (op: 0 sum, 1 max, 2 min) */
int b200_allreduce(b200_ctx* ctx, double* dev_buf, int n, int op);

#ifdef __cplusplus
}
#endif
#endif /* B200_STS_H */
