"""GPU parity tests of the sm_100a kernels, called through the C-ABI (include/b200_sts.h) with torch
providing device memory only.  Expected values come from the CPU oracle on the same seeded inputs.
Bar: elementwise / stencil / pack / Jacobi / ADR results are BIT-EXACT (IEEE binary64, no FMA, the
reference's association order); reductions are deterministic and agree to 1e-13 relative (tree vs
sequential summation order -- the reference itself moves by that much between 1 and 4 ranks)."""
import ctypes
import math

import numpy as np
import pytest
from conftest import OrcAdr, P, make_grid

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx(b200):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    c = b200.Context(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


# ------------------------------------------------------------------------------ elementwise ops
@pytest.mark.parametrize("n", [1, 2, 3, 255, 1000, (1 << 20) + 1])
@pytest.mark.parametrize("nterms", [1, 2, 3, 5, 8])
def test_lincomb_bit_exact(ctx, orc, n, nterms):
    rng = np.random.default_rng(n + nterms)
    X = [rng.standard_normal(n) for _ in range(nterms)]
    c = list(rng.standard_normal(nterms))
    c[0] = 1.0 if nterms % 2 else c[0]
    want = np.zeros(n)
    arr = (ctypes.c_void_p * nterms)(*[v.ctypes.data for v in X])
    orc.orc_linear_combination(nterms, (ctypes.c_double * nterms)(*c), arr, P(want), ctypes.c_int64(n))
    dX = [dev(v) for v in X]
    z = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.lincomb(c, dX, z)
    assert np.array_equal(host(z), want)
    # in place: z aliases the first operand (VScaleBy / Vaxpy forms)
    ctx.lincomb(c, dX, dX[0])
    assert np.array_equal(host(dX[0]), want)


@pytest.mark.parametrize("n", [1, 7, 4096, 100001])
def test_unary_binary_ops_bit_exact(ctx, b200, orc, n):
    lib = b200.kernel_lib()
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n) + 3.0
    dx, dy = dev(x), dev(y)
    z = torch.empty(n, dtype=torch.float64, device="cuda")
    N = ctypes.c_int64(n)
    h = ctx.handle
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    b200.check(lib.b200_prod(h, p(dx), p(dy), p(z), N)); assert np.array_equal(host(z), x * y)
    b200.check(lib.b200_div(h, p(dx), p(dy), p(z), N)); assert np.array_equal(host(z), x / y)
    b200.check(lib.b200_abs(h, p(dx), p(z), N)); assert np.array_equal(host(z), np.abs(x))
    b200.check(lib.b200_inv(h, p(dy), p(z), N)); assert np.array_equal(host(z), 1.0 / y)
    b200.check(lib.b200_addconst(h, p(dx), ctypes.c_double(1e-10), p(z), N)); assert np.array_equal(host(z), x + 1e-10)
    b200.check(lib.b200_const(h, ctypes.c_double(-2.5), p(z), N)); assert np.array_equal(host(z), np.full(n, -2.5))
    for sign in (+1, -1):
        b200.check(lib.b200_scale_sumdiff(h, ctypes.c_double(0.37), p(dx), p(dy), sign, p(z), N))
        assert np.array_equal(host(z), 0.37 * (x + sign * y))
    # ewt = 1/(rtol*|y| + atol) must equal the four-op sequence of arkEwtSetSS
    tmp, want = np.zeros(n), np.zeros(n)
    orc.orc_ewt_ss(P(x), ctypes.c_double(1e-5), ctypes.c_double(1e-10), P(tmp), P(want), N)
    b200.check(lib.b200_ewt_ss(h, p(dx), ctypes.c_double(1e-5), ctypes.c_double(1e-10), p(z), N))
    assert np.array_equal(host(z), want)


@pytest.mark.parametrize("n", [1, 2, 31, 1025, 1 << 20, (1 << 22) + 3])
def test_reductions(ctx, orc, n):
    rng = np.random.default_rng(n)
    x, w = rng.standard_normal(n), rng.random(n) + 0.5
    dx, dw = dev(x), dev(w)
    dot = ctx.reduce("dot", dx, dw)
    assert dot == pytest.approx(math.fsum(x * w), rel=1e-13, abs=1e-13)
    wsq = ctx.reduce("wsqrsum", dx, dw)
    assert wsq == pytest.approx(math.fsum((x * w) ** 2), rel=1e-13)
    assert wsq == pytest.approx(orc.orc_wsqrsum(P(x), P(w), ctypes.c_int64(n)), rel=1e-12)
    assert ctx.reduce("maxnorm", dx) == np.max(np.abs(x))
    assert ctx.reduce("min", dx) == np.min(x)
    assert ctx.reduce("l1norm", dx) == pytest.approx(math.fsum(np.abs(x)), rel=1e-13)
    # deterministic: bitwise repeatable
    assert ctx.reduce("dot", dx, dw) == dot and ctx.reduce("wsqrsum", dx, dw) == wsq


# ------------------------------------------------------------------------------- stencil stage
def geometry(b200, orc, nx, ny, inhom=True, kx=0.9, ky=1.7, halos=None):
    g = make_grid(max(nx, 2), max(ny, 2), kx=kx, ky=ky, inhom=inhom, nx_loc=nx, ny_loc=ny)
    tabs = [np.zeros(nx), np.zeros(nx), np.zeros(ny), np.zeros(ny)]
    orc.orc_coeff_tables(ctypes.byref(g), *[P(t) for t in tabs])
    dt = [dev(t) for t in tabs]
    hp = [None, None, None, None]
    dh = []
    if halos is not None:
        for k, hv in enumerate(halos):
            if hv is not None:
                t = dev(hv)
                dh.append(t)
                hp[k] = t.data_ptr()
    geom = b200.StencilGeom(nx, ny, dt[0].data_ptr(), dt[1].data_ptr(), dt[2].data_ptr(), dt[3].data_ptr(), *hp)
    return g, geom, (dt, dh)


PATTERNS = {
    "rhs": [2],
    "ssp_stage": [1, 2],
    "axpy_v": [0, 2],
    "dq": [2, 0],
    "ssp_close": [1, 0, 2],
    "sts_embed": [0, 1, 0, 2],
    "sts_stage": [2, 0, 0, 1, 0],
    "general3": [0, 2, 0],          # not a compiled pattern -> runtime-pattern kernel
    "general6": [0, 0, 2, 1, 0, 0],
    "general8": [0, 0, 0, 0, 0, 2, 0, 1],
}
SIZES = [(64, 48), (514, 33), (1024, 40), (2, 2), (4, 5), (75, 51), (33, 2)]


def expected_stage(orc, g, x, coeffs, srcs, vecs, halos):
    n = x.size
    L = np.zeros(n)
    hp = [P(h) if h is not None else None for h in (halos or [None] * 4)]
    orc.orc_laplacian(ctypes.byref(g), P(x), P(L), *hp)
    terms = [L if s == 2 else (x if s == 1 else v) for s, v in zip(srcs, vecs)]
    z = np.zeros(n)
    arr = (ctypes.c_void_p * len(terms))(*[t.ctypes.data for t in terms])
    orc.orc_linear_combination(len(terms), (ctypes.c_double * len(terms))(*coeffs), arr, P(z), ctypes.c_int64(n))
    return z, L


@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("pat", sorted(PATTERNS))
@pytest.mark.parametrize("halo_mode", ["wrap", "all", "x_only"])
def test_stencil_lincomb_bit_exact(ctx, b200, orc, size, pat, halo_mode):
    nx, ny = size
    srcs = PATTERNS[pat]
    rng = np.random.default_rng(nx * 131 + ny + len(srcs))
    x = rng.standard_normal(nx * ny)
    vecs = [rng.standard_normal(nx * ny) if s == 0 else None for s in srcs]
    coeffs = list(rng.standard_normal(len(srcs)))
    halos = None
    if halo_mode != "wrap":
        halos = [rng.standard_normal(ny), rng.standard_normal(ny), rng.standard_normal(nx), rng.standard_normal(nx)]
        if halo_mode == "x_only":
            halos[2] = halos[3] = None
    g, geom, keep = geometry(b200, orc, nx, ny, halos=halos)
    want_z, want_L = expected_stage(orc, g, x, coeffs, srcs, vecs, halos)
    dx = dev(x)
    dv = [dev(v) if v is not None else None for v in vecs]
    z = torch.full((nx * ny,), float("nan"), dtype=torch.float64, device="cuda")
    f = torch.full((nx * ny,), float("nan"), dtype=torch.float64, device="cuda")
    sw, se = torch.zeros(ny, dtype=torch.float64, device="cuda"), torch.zeros(ny, dtype=torch.float64, device="cuda")
    ss, sn = torch.zeros(nx, dtype=torch.float64, device="cuda"), torch.zeros(nx, dtype=torch.float64, device="cuda")
    ex = b200.StageExtras(f.data_ptr(), sw.data_ptr(), se.data_ptr(), ss.data_ptr(), sn.data_ptr(), None, None)
    ctx.stencil_lincomb(geom, dx, coeffs, srcs, dv, z, ex, region=0)
    assert np.array_equal(host(z), want_z)
    assert np.array_equal(host(f), want_L)
    Z = want_z.reshape(ny, nx)
    assert np.array_equal(host(sw), Z[:, 0]) and np.array_equal(host(se), Z[:, -1])
    assert np.array_equal(host(ss), Z[0, :]) and np.array_equal(host(sn), Z[-1, :])
    # interior (region 2) + ring (region 1) must tile the sub-domain exactly
    if nx >= 4 and ny >= 4:
        z2 = torch.full((nx * ny,), float("nan"), dtype=torch.float64, device="cuda")
        f2 = torch.full((nx * ny,), float("nan"), dtype=torch.float64, device="cuda")
        ex2 = b200.StageExtras(f2.data_ptr(), None, None, None, None, None, None)
        ctx.stencil_lincomb(geom, dx, coeffs, srcs, dv, z2, ex2, region=2)
        inter = host(z2).reshape(ny, nx)
        assert np.all(np.isnan(inter[0, :])) and np.all(np.isnan(inter[:, 0])) and np.all(np.isnan(inter[-1, :])) and np.all(np.isnan(inter[:, -1]))
        ctx.stencil_lincomb(geom, dx, coeffs, srcs, dv, z2, ex2, region=1)
        assert np.array_equal(host(z2), want_z) and np.array_equal(host(f2), want_L)


@pytest.mark.parametrize("size", [(64, 48), (1024, 64), (75, 51)], ids=lambda s: "%dx%d" % s)
def test_fused_wrms_matches_separate_norm(ctx, b200, orc, size):
    nx, ny = size
    rng = np.random.default_rng(99)
    x = rng.standard_normal(nx * ny)
    yn, fn, w = rng.standard_normal(nx * ny), rng.standard_normal(nx * ny), rng.random(nx * ny) + 0.1
    srcs, coeffs = [0, 1, 0, 2], [0.8, -0.8, 0.4e-3, 0.4e-3]
    g, geom, keep = geometry(b200, orc, nx, ny)
    want_z, _ = expected_stage(orc, g, x, coeffs, srcs, [yn, None, fn, None], None)
    z = torch.empty(nx * ny, dtype=torch.float64, device="cuda")
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    dw = dev(w)
    ex = b200.StageExtras(None, None, None, None, None, dw.data_ptr(), res.data_ptr())
    ctx.stencil_lincomb(geom, dev(x), coeffs, srcs, [dev(yn), None, dev(fn), None], z, ex)
    assert np.array_equal(host(z), want_z)
    want = orc.orc_wsqrsum(P(want_z), P(w), ctypes.c_int64(nx * ny))
    assert float(host(res)[0]) == pytest.approx(want, rel=1e-13)


@pytest.mark.parametrize("size", [(64, 48), (1024, 64), (75, 51), (32, 32)], ids=lambda s: "%dx%d" % s)
def test_fused_next_error_weights_match_arkEwtSetSS(ctx, b200, orc, size):
    """The closing stage of an adaptive step with the error weights of its stencil input fused in
    (b200_stage_extras.ewt_out): ewt = 1/(rtol |x| + atol) bit for bit as the N_VAbs / N_VScale / N_VAddConst / N_VInv
    sequence of arkEwtSetSS (arkode.c:2932-2944, oracle orc_ewt_ss), its norm sum (x ewt)^2 to reduction-order
    rounding, and z / the WRMS sum of z unchanged."""
    nx, ny = size
    n = nx * ny
    rng = np.random.default_rng(7 + nx)
    x = rng.standard_normal(n)
    x[::17] = 0.0
    yn, fn, w = rng.standard_normal(n), rng.standard_normal(n), rng.random(n) + 0.1
    srcs, coeffs = [0, 1, 0, 2], [0.8, -0.8, 0.4e-3, 0.4e-3]
    rtol, atol = 1e-5, 1e-10
    g, geom, keep = geometry(b200, orc, nx, ny)
    want_z, _ = expected_stage(orc, g, x, coeffs, srcs, [yn, None, fn, None], None)
    want_e, tmp = np.empty(n), np.empty(n)
    orc.orc_ewt_ss(P(x), ctypes.c_double(rtol), ctypes.c_double(atol), P(tmp), P(want_e), ctypes.c_int64(n))
    z = torch.empty(n, dtype=torch.float64, device="cuda")
    e = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    res = torch.zeros(2, dtype=torch.float64, device="cuda")
    dw = dev(w)
    ex = b200.StageExtras(None, None, None, None, None, dw.data_ptr(), res.data_ptr(), e.data_ptr(), rtol, atol,
                          res.data_ptr() + 8)
    ctx.stencil_lincomb(geom, dev(x), coeffs, srcs, [dev(yn), None, dev(fn), None], z, ex)
    assert np.array_equal(host(z), want_z)
    assert np.array_equal(host(e), want_e)
    r = host(res)
    assert float(r[0]) == pytest.approx(orc.orc_wsqrsum(P(want_z), P(w), ctypes.c_int64(n)), rel=1e-13)
    assert float(r[1]) == pytest.approx(orc.orc_wsqrsum(P(x), P(want_e), ctypes.c_int64(n)), rel=1e-13)


def test_stage_chain_equals_oracle_rkc_step(ctx, b200, orc):
    """Drive the fused kernel exactly as LSRKStep's RKC loop does (coefficients from the oracle's
    restatement of arkode_lsrkstep.c:629-717) and compare the whole step with orc_step_rkc."""
    from conftest import RHS_FN, OrcStepWs

    nx, ny, h = 128, 96, 1.0e-3
    g, geom, keep = geometry(b200, orc, nx, ny, inhom=True, kx=1.0, ky=0.5)
    n = nx * ny
    yn = np.zeros(n)
    orc.orc_initial(ctypes.byref(g), P(yn))
    fn = np.zeros(n)
    orc.orc_laplacian(ctypes.byref(g), P(yn), P(fn), None, None, None, None)
    sr = abs(orc.orc_dom_eig(ctypes.byref(g)) * 1.01)
    # oracle step, recording every linear combination's coefficients through the RHS callback order
    vec = {k: np.zeros(n) for k in ("ycur", "tempv1", "tempv2", "tempv3")}
    ewt = np.ones(n)

    def rhs(t, y, f, user):
        ya = np.ctypeslib.as_array(y, shape=(n,))
        fa = np.ctypeslib.as_array(f, shape=(n,))
        orc.orc_laplacian(ctypes.byref(g), P(ya), P(fa), None, None, None, None)
        return 0

    ws = OrcStepWs(n, n, yn.ctypes.data, fn.ctypes.data, vec["ycur"].ctypes.data, vec["tempv1"].ctypes.data,
                   vec["tempv2"].ctypes.data, vec["tempv3"].ctypes.data, ewt.ctypes.data, 1, 0)
    dsm = ctypes.c_double()
    s = orc.orc_step_rkc(ctypes.byref(ws), RHS_FN(rhs), None, ctypes.c_double(0.0), ctypes.c_double(h),
                         ctypes.c_double(sr), ctypes.byref(dsm))
    assert s >= 3
    # same recurrence on the GPU with fused launches
    w0 = 1.0 + 2.0 / (13.0 * (float(s) * float(s)))
    temp1 = w0 * w0 - 1.0
    temp2 = math.sqrt(temp1)
    arg = s * math.log(w0 + temp2)
    w1 = math.sinh(arg) * temp1 / (math.cosh(arg) * s * temp2 - w0 * math.sinh(arg))
    bjm1 = 1.0 / ((2.0 * w0) * (2.0 * w0))
    bjm2 = bjm1
    mus = w1 * bjm1
    d_yn, d_fn = dev(yn), dev(fn)
    t1 = d_yn.clone()
    t2 = torch.empty_like(d_yn)
    ctx.lincomb([1.0, h * mus], [d_yn, d_fn], t2)
    zjm1, zjm2, dzjm1, dzjm2, d2zjm1, d2zjm2 = w0, 1.0, 1.0, 0.0, 0.0, 0.0
    ycur = torch.empty_like(d_yn)
    for j in range(2, s + 1):
        zj = 2.0 * w0 * zjm1 - zjm2
        dzj = 2.0 * w0 * dzjm1 - dzjm2 + 2.0 * zjm1
        d2zj = 2.0 * w0 * d2zjm1 - d2zjm2 + 4.0 * dzjm1
        bj = d2zj / (dzj * dzj)
        ajm1 = 1.0 - zjm1 * bjm1
        mu = 2.0 * w0 * bj / bjm1
        nu = -bj / bjm2
        mus = mu * w1 / w0
        ctx.stencil_lincomb(geom, t2, [mus * h, nu, 1.0 - mu - nu, mu, -mus * ajm1 * h], [2, 0, 0, 1, 0],
                            [None, t1, d_yn, None, d_fn], ycur)
        if j < s:
            t1, t2, ycur = t2, ycur, t1
            bjm2, bjm1, zjm2, zjm1, dzjm2, dzjm1, d2zjm2, d2zjm1 = bjm1, bj, zjm1, zj, dzjm1, dzj, d2zjm1, d2zj
    assert np.array_equal(host(ycur), vec["ycur"])


def test_pack_and_jacobi_bit_exact(ctx, b200, orc):
    lib = b200.kernel_lib()
    nx, ny = 130, 37
    g = make_grid(256, 64, kx=1.3, ky=0.4, inhom=True, nx_loc=nx, ny_loc=ny, is_=17, js=5)
    rng = np.random.default_rng(4)
    u = rng.standard_normal(nx * ny)
    want = [np.zeros(ny), np.zeros(ny), np.zeros(nx), np.zeros(nx)]
    orc.orc_pack(ctypes.byref(g), P(u), *[P(w) for w in want])
    bufs = [torch.zeros(k, dtype=torch.float64, device="cuda") for k in (ny, ny, nx, nx)]
    du = dev(u)
    b200.check(lib.b200_pack_halo(ctx.handle, ctypes.c_void_p(du.data_ptr()), ctypes.c_int64(nx), ctypes.c_int64(ny),
                                  *[ctypes.c_void_p(b.data_ptr()) for b in bufs]))
    for b, w in zip(bufs, want):
        assert np.array_equal(host(b), w)
    # Jacobi: tables with the reference's own (unshifted) coordinates, preconditioner_jacobi.cpp:23-37
    gamma = 0.0123
    wantd = np.zeros(nx * ny)
    orc.orc_jacobi_setup(ctypes.byref(g), ctypes.c_double(gamma), P(wantd))
    pxw = np.array([orc.orc_coeff_x(ctypes.c_double((g.is_ + i) * g.dx), ctypes.byref(g)) / (g.dx * g.dx) for i in range(nx)])
    pxe = np.array([orc.orc_coeff_x(ctypes.c_double((g.is_ + i + 1) * g.dx), ctypes.byref(g)) / (g.dx * g.dx) for i in range(nx)])
    pys = np.array([orc.orc_coeff_y(ctypes.c_double((g.js + j) * g.dy), ctypes.byref(g)) / (g.dy * g.dy) for j in range(ny)])
    pyn = np.array([orc.orc_coeff_y(ctypes.c_double((g.js + j + 1) * g.dy), ctypes.byref(g)) / (g.dy * g.dy) for j in range(ny)])
    tabs = [dev(t) for t in (pxw, pxe, pys, pyn)]
    d = torch.empty(nx * ny, dtype=torch.float64, device="cuda")
    b200.check(lib.b200_jacobi_setup(ctx.handle, ctypes.c_int64(nx), ctypes.c_int64(ny), *[ctypes.c_void_p(t.data_ptr()) for t in tabs],
                                     ctypes.c_double(gamma), ctypes.c_void_p(d.data_ptr())))
    assert np.array_equal(host(d), wantd)


# ------------------------------------------------------------------------------------- adr 2-D
@pytest.mark.parametrize("size", [(16, 12), (64, 64), (101, 33)], ids=lambda s: "%dx%d" % s)
def test_adr_kernels_bit_exact(ctx, b200, orc, size):
    lib = b200.kernel_lib()
    nx, ny = size
    p = OrcAdr(nx, ny, 1.0 / nx, 1.0 / ny, -0.5, 1.0, 0.4, 0.7, 1e-2, 1.3, 1.0)
    bp = b200.AdrParams(nx, ny, 1.0 / nx, 1.0 / ny, -0.5, 1.0, 0.4, 0.7, 1e-2, 1.3, 1.0)
    n = 2 * nx * ny
    y = np.zeros(n)
    orc.orc_adr_ic(ctypes.byref(p), ctypes.c_double(0.0), ctypes.c_double(0.0), P(y))
    y += 0.01 * np.random.default_rng(1).standard_normal(n)
    fa, fd, fr, tmp, far = (np.zeros(n) for _ in range(5))
    orc.orc_adr_advection(ctypes.byref(p), P(y), P(fa))
    orc.orc_adr_diffusion(ctypes.byref(p), P(y), P(fd))
    orc.orc_adr_reaction(ctypes.byref(p), P(y), P(fr))
    orc.orc_adr_adv_react(ctypes.byref(p), P(y), P(tmp), P(far))
    dy = dev(y)
    f = torch.empty(n, dtype=torch.float64, device="cuda")
    for mode, want in ((1, fa), (2, fd), (4, fr), (5, far), (7, (fa + fd) + fr)):
        b200.check(lib.b200_adr_rhs(ctx.handle, ctypes.byref(bp), mode, ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(f.data_ptr())))
        assert np.array_equal(host(f), want), mode
    # fused STS stage on the diffusion partition
    rng = np.random.default_rng(2)
    v1, v2, v4 = (rng.standard_normal(n) for _ in range(3))
    c = [1e-3, -0.3, 0.2, 1.1, -2e-4]
    want = c[0] * fd
    for ck, vk in zip(c[1:], (v1, v2, y, v4)):
        want = want + ck * vk
    z = torch.empty(n, dtype=torch.float64, device="cuda")
    f2 = torch.empty(n, dtype=torch.float64, device="cuda")
    dv = [dev(v1), dev(v2), dev(v4)]
    cc = (ctypes.c_double * 5)(*c)
    ss = (ctypes.c_int * 5)(2, 0, 0, 1, 0)
    vv = (ctypes.c_void_p * 5)(0, dv[0].data_ptr(), dv[1].data_ptr(), 0, dv[2].data_ptr())
    b200.check(lib.b200_adr_diffusion_lincomb(ctx.handle, ctypes.byref(bp), ctypes.c_void_p(dy.data_ptr()), 5, cc, ss, vv,
                                              ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(f2.data_ptr())))
    assert np.array_equal(host(z), want) and np.array_equal(host(f2), fd)


# ---------------------------------------------------- size-independent properties at full size
@pytest.mark.parametrize("n", [4096, 16384])
def test_full_size_shift_equivariance_and_conservation(ctx, b200, n):
    """BASELINE sizes (4096^2, 16384^2) are too big for the CPU oracle, so check properties of the
    periodic constant-coefficient operator instead: (i) rolling the input rolls the output,
    bit for bit (exercises the wrap-around paths and every tile edge); (ii) sum(L u) = 0 up to
    rounding (discrete conservation); (iii) the fused stage equals stage-by-parts on the device."""
    nx = ny = n
    N = nx * ny
    gen = torch.Generator(device="cuda").manual_seed(1234)
    u = torch.rand(N, dtype=torch.float64, device="cuda", generator=gen)
    cx = torch.full((nx,), 1.0 / 0.00038351860508939673 ** 2, dtype=torch.float64, device="cuda")
    cy = torch.full((ny,), 0.5 / 0.00073246658121223218 ** 2, dtype=torch.float64, device="cuda")
    geom = b200.StencilGeom(nx, ny, cx.data_ptr(), cx.data_ptr(), cy.data_ptr(), cy.data_ptr(), None, None, None, None)
    f = torch.empty_like(u)
    ctx.stencil_lincomb(geom, u, [1.0], [2], [None], f)
    ur = torch.roll(u.view(ny, nx), shifts=(37, -5), dims=(0, 1)).contiguous().view(-1)
    fr = torch.empty_like(u)
    ctx.stencil_lincomb(geom, ur, [1.0], [2], [None], fr)
    torch.cuda.synchronize()
    assert torch.equal(fr.view(ny, nx), torch.roll(f.view(ny, nx), shifts=(37, -5), dims=(0, 1)))
    del ur, fr
    total, scale = float(f.sum()), float(f.abs().sum())
    assert abs(total) <= 1e-10 * scale
    # fused stage == separate ops (device-side check of the fusion itself, bit-exact)
    v1 = torch.rand(N, dtype=torch.float64, device="cuda", generator=gen)
    c = [1e-9, -0.3, 1.1]
    z = torch.empty_like(u)
    ctx.stencil_lincomb(geom, u, c, [2, 0, 1], [None, v1, None], z)
    z2 = torch.empty_like(u)
    ctx.lincomb(c, [f, v1, u], z2)
    torch.cuda.synchronize()
    assert torch.equal(z, z2)


# --------------------------------------------------------------- temporally blocked stage chains
@pytest.mark.parametrize("size", [(128, 16), (192, 70), (1024, 96), (2050, 33)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("rows", [64, 5])
@pytest.mark.parametrize("variant", [0, 1], ids=["march", "quad"])
def test_stencil_chain_equals_single_stage_launches(ctx, b200, size, k, rows, variant):
    """b200_stencil_chain (K stages in one pass) must be bit-identical to K b200_stencil_lincomb launches
    (which are themselves pinned against the oracle above), including periodic wrap in x and y, partial
    windows (nx not a multiple of 60/56, or of 120 for the four-cells-per-thread kernel) and partial row
    blocks."""
    nx, ny = size
    n = nx * ny
    rng = np.random.default_rng(nx * 7 + ny + k)
    lib = b200.kernel_lib()
    lib.b200_set_chain_rows(rows)
    lib.b200_set_chain_variant(variant)
    lib.b200_set_contract(0)
    cx = [dev(rng.random(nx) + 0.5) for _ in range(2)]
    cy = [dev(rng.random(ny) + 0.5) for _ in range(2)]
    g = b200.StencilGeom(nx, ny, cx[0].data_ptr(), cx[1].data_ptr(), cy[0].data_ptr(), cy[1].data_ptr(), None, None, None, None)
    x, p2, yn, fn = (dev(rng.standard_normal(n)) for _ in range(4))
    coeffs = [[1e-3 * (l + 1), -0.3 + 0.1 * l, 0.2, 1.1 - 0.05 * l, -2e-4] for l in range(k)]
    # reference: one launch per stage
    zs = []
    prev, cur = p2, x
    for l in range(k):
        z = torch.empty(n, dtype=torch.float64, device="cuda")
        ctx.stencil_lincomb(g, cur, coeffs[l], [2, 0, 0, 1, 0], [None, prev, yn, None, fn], z)
        zs.append(z)
        prev, cur = cur, z
    # chain: all levels stored
    outs = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(k)]
    ctx.stencil_chain(g, x, p2, yn, fn, coeffs, outs)
    ctx.sync()
    for l in range(k):
        assert np.array_equal(host(outs[l]), host(zs[l])), "level %d" % (l + 1)
    # chain: only the last two levels stored (what LSRKStep needs)
    outs2 = [None] * k
    outs2[k - 1] = torch.full((n,), np.nan, dtype=torch.float64, device="cuda")
    outs2[k - 2] = torch.full((n,), np.nan, dtype=torch.float64, device="cuda")
    ctx.stencil_chain(g, x, p2, yn, fn, coeffs, outs2)
    ctx.sync()
    assert np.array_equal(host(outs2[k - 1]), host(zs[k - 1])) and np.array_equal(host(outs2[k - 2]), host(zs[k - 2]))
    lib.b200_last_chain_kernel.restype = ctypes.c_char_p
    assert lib.b200_last_chain_kernel() == (b"k_chain_quad" if variant == 1 else b"k_chain_march")
    lib.b200_set_chain_rows(0)
    lib.b200_set_chain_variant(0)


@pytest.mark.parametrize("size", [(256, 40), (1024, 96)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("k", [2, 4, 6])
def test_stencil_chain_fma_flavour(ctx, b200, size, k):
    """b200_set_contract(1): the chained stages' multiply-adds become FMAs.  Both chain kernels contract the
    same operations, so they agree bit for bit with each other, and with the exact flavour to rounding."""
    nx, ny = size
    n = nx * ny
    rng = np.random.default_rng(nx + ny + k)
    lib = b200.kernel_lib()
    cx = [dev(rng.random(nx) + 0.5) for _ in range(2)]
    cy = [dev(rng.random(ny) + 0.5) for _ in range(2)]
    g = b200.StencilGeom(nx, ny, cx[0].data_ptr(), cx[1].data_ptr(), cy[0].data_ptr(), cy[1].data_ptr(), None, None, None, None)
    x, p2, yn, fn = (dev(rng.standard_normal(n)) for _ in range(4))
    coeffs = [[1e-3 * (l + 1), -0.3 + 0.1 * l, 0.2, 1.1 - 0.05 * l, -2e-4] for l in range(k)]
    res = {}
    for contract, variant in ((0, 1), (1, 1), (1, 0)):
        lib.b200_set_contract(contract)
        lib.b200_set_chain_variant(variant)
        outs = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(k)]
        ctx.stencil_chain(g, x, p2, yn, fn, coeffs, outs)
        ctx.sync()
        res[(contract, variant)] = [host(o) for o in outs]
    lib.b200_set_contract(0)
    lib.b200_set_chain_variant(0)
    for l in range(k):
        assert np.array_equal(res[(1, 1)][l], res[(1, 0)][l]), l
        e, f = res[(0, 1)][l], res[(1, 1)][l]
        rel = np.linalg.norm(e - f) / np.linalg.norm(e)
        assert 0.0 < rel < 1e-13, (l, rel)


@pytest.mark.parametrize("size", [(256, 40), (1030, 70)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
def test_stencil_chain_uniform_coefficient_flavour(ctx, b200, size, k):
    """b200_stencil_geom.uniform (homogeneous problem): coefficients from kernel parameters instead of
    tables; bit-identical to the table-driven launch, in both arithmetic flavours."""
    nx, ny = size
    n = nx * ny
    rng = np.random.default_rng(nx + 3 * ny + k)
    lib = b200.kernel_lib()
    u4 = [0.8125, 0.8125, 2.3, 2.3]
    cx = [torch.full((nx,), u4[0], dtype=torch.float64, device="cuda"), torch.full((nx,), u4[1], dtype=torch.float64, device="cuda")]
    cy = [torch.full((ny,), u4[2], dtype=torch.float64, device="cuda"), torch.full((ny,), u4[3], dtype=torch.float64, device="cuda")]
    base = (nx, ny, cx[0].data_ptr(), cx[1].data_ptr(), cy[0].data_ptr(), cy[1].data_ptr(), None, None, None, None)
    g_tab = b200.StencilGeom(*base)
    g_uni = b200.StencilGeom(*base, 1, *u4)
    x, p2, yn, fn = (dev(rng.standard_normal(n)) for _ in range(4))
    coeffs = [[1e-3 * (l + 1), -0.3 + 0.1 * l, 0.2, 1.1 - 0.05 * l, -2e-4] for l in range(k)]
    lib.b200_set_chain_variant(0)
    for contract in (0, 1):
        lib.b200_set_contract(contract)
        res = []
        for g in (g_tab, g_uni):
            outs = [torch.full((n,), np.nan, dtype=torch.float64, device="cuda") for _ in range(k)]
            ctx.stencil_chain(g, x, p2, yn, fn, coeffs, outs)
            ctx.sync()
            res.append([host(o) for o in outs])
        for l in range(k):
            assert np.array_equal(res[0][l], res[1][l]), (contract, l)
    lib.b200_set_contract(0)
