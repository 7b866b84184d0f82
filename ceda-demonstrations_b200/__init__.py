"""ceda-demonstrations_b200 -- B200-native explicit super-time-stepping hot path.

Python host-side bindings (ctypes) over the two in-tree shared libraries

    lib/libb200sts.so            CUDA kernels + C-ABI           (include/b200_sts.h)
    lib/libb200sts_sundials.so   N_Vector_B200 + diffusion_2D   (include/nvector_b200.h,
                                 problem layer and driver        include/b200_diffusion2d.h)

The package directory name contains a hyphen, so import it with

    import importlib; b200 = importlib.import_module("ceda-demonstrations_b200")

There is no CPU fallback anywhere: the libraries must be built (``make`` or
``__graft_entry__.build()``) and every compute entry needs a CUDA device.
torch is used only for device memory, streams, pinned host memory and
torch.distributed rendezvous.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
BIN_DIR = os.path.join(_HERE, "bin")
KERNEL_LIB = os.path.join(LIB_DIR, "libb200sts.so")
SUNDIALS_LIB = os.path.join(LIB_DIR, "libb200sts_sundials.so")
DRIVER_BIN = os.path.join(BIN_DIR, "diffusion_2D_b200")
ADR_DRIVER_BIN = os.path.join(BIN_DIR, "adr2d_b200")

MAX_TERMS = 8
SRC_VECTOR, SRC_CENTRE, SRC_STENCIL = 0, 1, 2

c_double_p = ctypes.POINTER(ctypes.c_double)


class StencilGeom(ctypes.Structure):
    """b200_stencil_geom (include/b200_sts.h)."""

    _fields_ = [
        ("nx", ctypes.c_int64), ("ny", ctypes.c_int64),
        ("cxw", ctypes.c_void_p), ("cxe", ctypes.c_void_p),
        ("cys", ctypes.c_void_p), ("cyn", ctypes.c_void_p),
        ("halo_w", ctypes.c_void_p), ("halo_e", ctypes.c_void_p),
        ("halo_s", ctypes.c_void_p), ("halo_n", ctypes.c_void_p),
        ("uniform", ctypes.c_int),
        ("u_cxw", ctypes.c_double), ("u_cxe", ctypes.c_double), ("u_cys", ctypes.c_double), ("u_cyn", ctypes.c_double),
    ]


class StageExtras(ctypes.Structure):
    """b200_stage_extras (include/b200_sts.h)."""

    _fields_ = [
        ("f_out", ctypes.c_void_p),
        ("send_w", ctypes.c_void_p), ("send_e", ctypes.c_void_p),
        ("send_s", ctypes.c_void_p), ("send_n", ctypes.c_void_p),
        ("wrms_w", ctypes.c_void_p), ("wrms_result", ctypes.c_void_p),
        ("ewt_out", ctypes.c_void_p), ("ewt_rtol", ctypes.c_double), ("ewt_atol", ctypes.c_double),
        ("ewt_result", ctypes.c_void_p),
    ]


class AdrParams(ctypes.Structure):
    """b200_adr_params (include/b200_sts.h)."""

    _fields_ = [
        ("nx", ctypes.c_int64), ("ny", ctypes.c_int64),
        ("dx", ctypes.c_double), ("dy", ctypes.c_double),
        ("cux", ctypes.c_double), ("cuy", ctypes.c_double),
        ("cvx", ctypes.c_double), ("cvy", ctypes.c_double),
        ("d", ctypes.c_double), ("A", ctypes.c_double), ("B", ctypes.c_double),
    ]


class D2DStats(ctypes.Structure):
    """b200_d2d_stats (include/b200_diffusion2d.h)."""

    _fields_ = [
        ("t", ctypes.c_double), ("h_last", ctypes.c_double), ("urms", ctypes.c_double),
        ("evolve_seconds", ctypes.c_double), ("spectral_radius", ctypes.c_double),
        ("steps", ctypes.c_long), ("step_attempts", ctypes.c_long), ("err_test_fails", ctypes.c_long),
        ("rhs_evals", ctypes.c_long), ("dee_rhs_evals", ctypes.c_long),
        ("dom_eig_updates", ctypes.c_long), ("max_stages", ctypes.c_long),
        ("lin_iters", ctypes.c_long), ("lin_rhs_evals", ctypes.c_long),
        ("prec_solves", ctypes.c_long), ("nonlin_iters", ctypes.c_long),
        ("fused_launches", ctypes.c_long), ("plain_rhs_launches", ctypes.c_long),
        ("aliased_copies", ctypes.c_long), ("wrms_fused", ctypes.c_long),
        ("buffers_allocated", ctypes.c_long),
        ("kernel_launches", ctypes.c_uint64),
        ("nx", ctypes.c_int64), ("ny", ctypes.c_int64), ("nx_loc", ctypes.c_int64),
        ("ny_loc", ctypes.c_int64), ("is_", ctypes.c_int64), ("js", ctypes.c_int64),
        ("npx", ctypes.c_int), ("npy", ctypes.c_int), ("rank", ctypes.c_int), ("nranks", ctypes.c_int),
        ("chain_launches", ctypes.c_long), ("chain_stages", ctypes.c_long),
        ("dq_fused", ctypes.c_long), ("ew_fused", ctypes.c_long),
    ]

    def as_dict(self):
        return {name.rstrip("_"): getattr(self, name) for name, _ in self._fields_}


class AdrStats(ctypes.Structure):
    """b200_adr_stats (include/b200_adr2d.h)."""

    _fields_ = [
        ("t", ctypes.c_double), ("evolve_seconds", ctypes.c_double),
        ("steps", ctypes.c_long), ("step_attempts", ctypes.c_long),
        ("rhs_evals_explicit", ctypes.c_long), ("rhs_evals_implicit", ctypes.c_long),
        ("lsrk_steps", ctypes.c_long), ("lsrk_rhs_evals", ctypes.c_long), ("lsrk_max_stages", ctypes.c_long),
        ("ark_steps", ctypes.c_long), ("ark_rhs_evals", ctypes.c_long),
        ("fused_launches", ctypes.c_long), ("plain_rhs_launches", ctypes.c_long),
        ("aliased_copies", ctypes.c_long), ("buffers_allocated", ctypes.c_long),
        ("kernel_launches", ctypes.c_uint64),
        ("nx", ctypes.c_int64), ("ny", ctypes.c_int64), ("neq", ctypes.c_int64),
        ("ark_rhs_evals_implicit", ctypes.c_long), ("nls_iters", ctypes.c_long),
        ("ls_setups", ctypes.c_long), ("jac_evals", ctypes.c_long),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_kernel_lib = None
_sundials_lib = None


def kernel_lib():
    """Load libb200sts.so (fails loudly if it has not been built)."""
    global _kernel_lib
    if _kernel_lib is None:
        if not os.path.exists(KERNEL_LIB):
            raise RuntimeError("%s is missing: run `make` (there is no CPU fallback)" % KERNEL_LIB)
        lib = ctypes.CDLL(KERNEL_LIB, mode=ctypes.RTLD_GLOBAL)
        lib.b200_last_error.restype = ctypes.c_char_p
        lib.b200_launch_count.restype = ctypes.c_uint64
        lib.b200_ctx_stream.restype = ctypes.c_void_p
        _kernel_lib = lib
    return _kernel_lib


def sundials_lib():
    """Load libb200sts_sundials.so (N_Vector_B200 + the diffusion_2D layer)."""
    global _sundials_lib
    if _sundials_lib is None:
        kernel_lib()
        if not os.path.exists(SUNDIALS_LIB):
            raise RuntimeError("%s is missing: run `make`" % SUNDIALS_LIB)
        _sundials_lib = ctypes.CDLL(SUNDIALS_LIB, mode=ctypes.RTLD_GLOBAL)
    return _sundials_lib


def check(rc, what="b200 call"):
    if rc != 0:
        msg = kernel_lib().b200_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))


def _ptr(t):
    """Device (or pinned host) pointer of a torch tensor / numpy array / int."""
    if t is None:
        return None
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    if hasattr(t, "data_ptr"):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


class Context:
    """b200_ctx: one per process / GPU.  With ``stream=None`` kernels run on torch's
    current stream of ``device`` so torch.cuda.Event timing brackets them."""

    def __init__(self, device=0, stream="torch"):
        lib = kernel_lib()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        sptr = None
        if stream == "torch":
            import torch

            torch.cuda.set_device(device)
            sptr = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            if not sptr.value:
                # the legacy default stream: give the library an explicit torch stream instead
                self._tstream = torch.cuda.Stream(device)
                torch.cuda.set_stream(self._tstream)
                sptr = ctypes.c_void_p(self._tstream.cuda_stream)
        elif stream is not None:
            sptr = ctypes.c_void_p(stream)
        check(lib.b200_ctx_create(int(device), sptr, ctypes.byref(self.handle)), "b200_ctx_create")

    def sync(self):
        check(self._lib.b200_ctx_sync(self.handle), "b200_ctx_sync")

    def close(self):
        if self.handle:
            self._lib.b200_ctx_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    # ---- thin wrappers used by tests / smoke -------------------------------------------------
    def lincomb(self, coeffs, vecs, z):
        n = len(coeffs)
        c = (ctypes.c_double * n)(*coeffs)
        v = (ctypes.c_void_p * n)(*[x.data_ptr() for x in vecs])
        check(self._lib.b200_lincomb(self.handle, n, c, v, _ptr(z), ctypes.c_int64(z.numel())), "b200_lincomb")

    def stencil_lincomb(self, geom, x, coeffs, srcs, vecs, z, extras=None, region=0):
        n = len(coeffs)
        c = (ctypes.c_double * n)(*coeffs)
        s = (ctypes.c_int * n)(*srcs)
        v = (ctypes.c_void_p * n)(*[(t.data_ptr() if t is not None else 0) for t in vecs])
        check(self._lib.b200_stencil_lincomb(self.handle, ctypes.byref(geom), _ptr(x), n, c, s, v, _ptr(z),
                                             ctypes.byref(extras) if extras is not None else None, int(region)),
              "b200_stencil_lincomb")

    def stencil_chain(self, geom, x, prev2, yn, fn, coeffs, outs):
        """b200_stencil_chain: len(coeffs) temporally blocked stages; outs[l] may be None."""
        k = len(coeffs)
        flat = [v for row in coeffs for v in row]
        c = (ctypes.c_double * (5 * k))(*flat)
        o = (ctypes.c_void_p * k)(*[(t.data_ptr() if t is not None else 0) for t in outs])
        check(self._lib.b200_stencil_chain(self.handle, ctypes.byref(geom), k, _ptr(x), _ptr(prev2), _ptr(yn),
                                           _ptr(fn), c, o), "b200_stencil_chain")

    def reduce(self, name, x, y=None):
        out = ctypes.c_double()
        fn = getattr(self._lib, "b200_" + name)
        if y is None:
            check(fn(self.handle, _ptr(x), ctypes.c_int64(x.numel()), ctypes.byref(out)), name)
        else:
            check(fn(self.handle, _ptr(x), _ptr(y), ctypes.c_int64(x.numel()), ctypes.byref(out)), name)
        return out.value


def block_decomposition(nx, ny, rank, nranks, npx=0, npy=0):
    """Host mirror of UserData::setup (diffusion_2D.cpp:243-317): the 2-D periodic block
    decomposition.  Pure Python so it can be tested without a GPU; the C++ driver's own
    implementation (b200_d2d_local_extent) must agree with it."""
    if not (npx > 0 and npy > 0):
        if npx > 0:
            npy = nranks // npx
        elif npy > 0:
            npx = nranks // npy
        else:
            b = 1
            f = 1
            while f * f <= nranks:
                if nranks % f == 0:
                    b = f
                f += 1
            npx, npy = nranks // b, b
    if npx * npy != nranks:
        raise ValueError("npx*npy != nranks")
    idx, idy = rank // npy, rank % npy

    def extent(n, p, c):
        q, r = divmod(n, p)
        s = q * c + min(c, r)
        return s, q + (1 if c < r else 0)

    is_, nxl = extent(nx, npx, idx)
    js, nyl = extent(ny, npy, idy)

    def cart(cx, cy):
        return (cx % npx) * npy + (cy % npy)

    return {
        "npx": npx, "npy": npy, "idx": idx, "idy": idy,
        "is": is_, "nx_loc": nxl, "js": js, "ny_loc": nyl,
        "ipW": cart(idx - 1, idy), "ipE": cart(idx + 1, idy),
        "ipS": cart(idx, idy - 1), "ipN": cart(idx, idy + 1),
    }


class Diffusion2D:
    """One diffusion_2D problem + ARKODE integrator on this rank's GPU (b200_d2d).

    ``args`` are the reference driver's command-line flags."""

    def __init__(self, args, rank=0, nranks=1, nccl_id=None, device=0, stream="torch"):
        lib = sundials_lib()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        argv = (ctypes.c_char_p * len(args))(*[str(a).encode() for a in args])
        sptr = None
        if stream == "torch":
            import torch

            torch.cuda.set_device(device)
            self._tstream = torch.cuda.Stream(device)
            torch.cuda.set_stream(self._tstream)
            sptr = ctypes.c_void_p(self._tstream.cuda_stream)
        idbuf = None
        if nccl_id is not None:
            idbuf = (ctypes.c_ubyte * 128)(*bytes(nccl_id))
        rc = lib.b200_d2d_create(len(args), argv, int(rank), int(nranks), idbuf, int(device), sptr,
                                 ctypes.byref(self.handle))
        if rc != 0:
            raise RuntimeError("b200_d2d_create failed: %s" % kernel_lib().b200_last_error().decode())

    def evolve(self, tout):
        if self._lib.b200_d2d_evolve(self.handle, ctypes.c_double(tout)) != 0:
            raise RuntimeError("b200_d2d_evolve failed")

    def step(self, nsteps=1):
        if self._lib.b200_d2d_step(self.handle, int(nsteps)) != 0:
            raise RuntimeError("b200_d2d_step failed")

    def get_state(self, host):
        if self._lib.b200_d2d_get_state(self.handle, _ptr(host)) != 0:
            raise RuntimeError("b200_d2d_get_state failed")

    def set_state(self, host, t=0.0):
        if self._lib.b200_d2d_set_state(self.handle, _ptr(host), ctypes.c_double(t)) != 0:
            raise RuntimeError("b200_d2d_set_state failed")

    def run_batches(self, host_ins, host_outs, t=0.0, nsteps=1):
        """b200_d2d_run_batches: independent states host_ins[i] (pinned) -> nsteps steps from time t ->
        host_outs[i] (pinned), uploads / downloads of neighbouring batches overlapping the integration."""
        n = len(host_ins)
        assert len(host_outs) == n
        ins = (ctypes.c_void_p * n)(*[_ptr(h) for h in host_ins])
        outs = (ctypes.c_void_p * n)(*[_ptr(h) for h in host_outs])
        if self._lib.b200_d2d_run_batches(self.handle, n, ins, outs, ctypes.c_double(t), int(nsteps)) != 0:
            raise RuntimeError("b200_d2d_run_batches failed: %s" % kernel_lib().b200_last_error().decode())

    def stats(self):
        s = D2DStats()
        self._lib.b200_d2d_get_stats(self.handle, ctypes.byref(s))
        return s.as_dict()

    def close(self):
        if self.handle:
            self._lib.b200_d2d_destroy(self.handle)
            self.handle = ctypes.c_void_p()


class Adr2D:
    """One adr 2-D Brusselator problem + ARKODE integrator on one GPU (b200_adr).

    ``args`` are the reference driver's command-line flags
    (adr/advection_diffusion_reaction_2d.hpp InputHelp)."""

    def __init__(self, args, device=0, stream="torch"):
        lib = sundials_lib()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        argv = (ctypes.c_char_p * len(args))(*[str(a).encode() for a in args])
        sptr = None
        if stream == "torch":
            import torch

            torch.cuda.set_device(device)
            self._tstream = torch.cuda.Stream(device)
            torch.cuda.set_stream(self._tstream)
            sptr = ctypes.c_void_p(self._tstream.cuda_stream)
        rc = lib.b200_adr_create(len(args), argv, int(device), sptr, ctypes.byref(self.handle))
        if rc != 0:
            raise RuntimeError("b200_adr_create failed: %s" % kernel_lib().b200_last_error().decode())

    def evolve(self, tout):
        if self._lib.b200_adr_evolve(self.handle, ctypes.c_double(tout)) != 0:
            raise RuntimeError("b200_adr_evolve failed")

    def step(self, nsteps=1):
        if self._lib.b200_adr_step(self.handle, int(nsteps)) != 0:
            raise RuntimeError("b200_adr_step failed")

    def get_state(self, host):
        if self._lib.b200_adr_get_state(self.handle, _ptr(host)) != 0:
            raise RuntimeError("b200_adr_get_state failed")

    def set_state(self, host, t=0.0):
        if self._lib.b200_adr_set_state(self.handle, _ptr(host), ctypes.c_double(t)) != 0:
            raise RuntimeError("b200_adr_set_state failed")

    def stats(self):
        s = AdrStats()
        self._lib.b200_adr_get_stats(self.handle, ctypes.byref(s))
        return s.as_dict()

    def close(self):
        if self.handle:
            self._lib.b200_adr_destroy(self.handle)
            self.handle = ctypes.c_void_p()


def nccl_unique_id():
    buf = (ctypes.c_ubyte * 128)()
    check(kernel_lib().b200_comm_unique_id(buf), "b200_comm_unique_id")
    return bytes(buf)
