"""CPU tests of the stage kernels' SOURCE (csrc/chain_march.cuh, chain_quad.cuh: K temporally blocked
stages per launch; csrc/stage_kernels.cuh + reduce_prims.cuh: the fused one-stage kernels with halo pack and
fused WRMS reduction; csrc/vector_kernels.cuh, halo_kernels.cuh, adr_kernels.cuh: elementwise ops, reductions,
halo packing, Jacobi diagonal, adr Brusselator kernels) run through the host emulation harness tests/emu (one OS thread per CUDA thread, warp shuffles,
cp.async groups in eager and lazy completion order).  They check the tiling / ring / halo / wrap
indexing and the arithmetic order against a numpy restatement of the stage recurrence
(diffusion_2D/diffusion.cpp:34-55 + arkode_lsrkstep.c:706-717), which is itself checked against the
oracle's orc_laplacian here.  Bar: bit-exact (two-rounding arithmetic); the FMA flavour of the two
kernels must agree with each other bit for bit and with the exact flavour to 1e-13.

The emulation is test infrastructure: the product path never loads it (the GPU tests in
test_kernels_gpu.py run the same cases on the device)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from conftest import ROOT, P, make_grid

EMU_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
    lib = ctypes.CDLL(os.path.join(EMU_DIR, "_build", "libemu_chain.so"))
    lib.emu_stencil_chain.restype = ctypes.c_int
    return lib


def stage_np(cxw, cxe, cys, cyn, x, p2, yn, fn, c):
    """One STS stage on a periodic grid, (ny, nx) arrays, the reference's association order."""
    uw, ue = np.roll(x, 1, axis=1), np.roll(x, -1, axis=1)
    us, un = np.roll(x, 1, axis=0), np.roll(x, -1, axis=0)
    L = (-((cxw + cxe)[None, :] + (cys + cyn)[:, None])) * x
    L = L + cxw[None, :] * uw
    L = L + cxe[None, :] * ue
    L = L + cys[:, None] * us
    L = L + cyn[:, None] * un
    z = c[0] * L
    z = z + c[1] * p2
    z = z + c[2] * yn
    z = z + c[3] * x
    z = z + c[4] * fn
    return z


def chain_np(cxw, cxe, cys, cyn, x, p2, yn, fn, coeffs):
    zs, prev, cur = [], p2, x
    for c in coeffs:
        z = stage_np(cxw, cxe, cys, cyn, cur, prev, yn, fn, c)
        zs.append(z)
        prev, cur = cur, z
    return zs


def coeffs_for(k):
    return [[1e-3 * (l + 1), -0.3 + 0.1 * l, 0.2, 1.1 - 0.05 * l, -2e-4] for l in range(k)]


def run_emu(emu, variant, k, fma, lazy, nx, ny, cx, cy, ops, coeffs, rows, store, halos=None, g=0, g2=0, uniform=None):
    n = nx * ny
    outs = [np.full(n, np.nan) if store[l] else None for l in range(k)]
    optr = (ctypes.c_void_p * k)(*[o.ctypes.data if o is not None else None for o in outs])
    cf = np.ascontiguousarray(np.array(coeffs, dtype=np.float64).ravel())
    hptr = None
    if halos is not None:
        hptr = (ctypes.c_void_p * 4)(*[h.ctypes.data for h in halos])
    rc = emu.emu_stencil_chain(variant, k, int(fma), int(lazy), ctypes.c_int64(nx), ctypes.c_int64(ny),
                               ctypes.c_void_p(cx[0]), ctypes.c_void_p(cx[1]), ctypes.c_void_p(cy[0]), ctypes.c_void_p(cy[1]),
                               P(ops[0]), P(ops[1]), P(ops[2]), P(ops[3]), P(cf), optr, rows, hptr, g, g2,
                               None if uniform is None else P(np.ascontiguousarray(np.array(uniform, dtype=np.float64))))
    return rc, outs


def test_numpy_stage_matches_oracle_laplacian(orc):
    nx, ny = 64, 48
    g = make_grid(nx, ny, kx=1.0, ky=0.5, inhom=True)
    tabs = [np.zeros(nx), np.zeros(nx), np.zeros(ny), np.zeros(ny)]
    orc.orc_coeff_tables(ctypes.byref(g), *[P(t) for t in tabs])
    rng = np.random.default_rng(5)
    u = rng.standard_normal((ny, nx))
    want = np.zeros(nx * ny)
    orc.orc_laplacian(ctypes.byref(g), P(np.ascontiguousarray(u.ravel())), P(want), None, None, None, None)
    zero = np.zeros_like(u)
    got = stage_np(tabs[0], tabs[1], tabs[2], tabs[3], u, zero, zero, zero, [1.0, 0.0, 0.0, 0.0, 0.0])
    assert np.array_equal(got.ravel(), want)


CASES = [  # (nx, ny, rows)
    (128, 16, 64),   # smallest supported field: a second warp window that wraps in x
    (132, 21, 5),    # partial last window, partial row blocks, short blocks (warm-up and drain only)
    (376, 26, 8),    # > 1 block in x for the quad kernel (4 warps x 120 cells), steady-state rows
    (190, 17, 9),    # nx % 4 == 2: the halves of a lane wrap at different lanes
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_rows%d" % c)
@pytest.mark.parametrize("k", [2, 3, 4])
@pytest.mark.parametrize("variant", [0, 1], ids=["march", "quad"])
def test_chain_kernels_periodic_wrap_bit_exact(emu, case, k, variant):
    nx, ny, rows = case
    rng = np.random.default_rng(nx * 31 + ny + k)
    tabs = [rng.random(nx) + 0.5, rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(ny) + 0.5]
    ops = [np.ascontiguousarray(rng.standard_normal(nx * ny)) for _ in range(4)]
    coeffs = coeffs_for(k)
    want = chain_np(*tabs, *[o.reshape(ny, nx) for o in ops], coeffs)
    cx = [tabs[0].ctypes.data, tabs[1].ctypes.data]
    cy = [tabs[2].ctypes.data, tabs[3].ctypes.data]
    for lazy in (0, 1):
        # all levels stored
        rc, outs = run_emu(emu, variant, k, 0, lazy, nx, ny, cx, cy, ops, coeffs, rows, [True] * k)
        assert rc == 0
        for l in range(k):
            assert np.array_equal(outs[l], want[l].ravel()), "level %d (lazy=%d)" % (l + 1, lazy)
    # only the last two levels stored (what LSRKStep needs)
    store = [l >= k - 2 for l in range(k)]
    rc, outs = run_emu(emu, variant, k, 0, 1, nx, ny, cx, cy, ops, coeffs, rows, store)
    assert rc == 0
    assert np.array_equal(outs[k - 1], want[k - 1].ravel()) and np.array_equal(outs[k - 2], want[k - 2].ravel())


@pytest.mark.parametrize("k", [5, 6])
@pytest.mark.parametrize("variant", [0, 1], ids=["march", "quad"])
def test_chain_kernels_deep_levels_bit_exact(emu, k, variant):
    nx, ny, rows = 192, 24, 7
    rng = np.random.default_rng(k)
    tabs = [rng.random(nx) + 0.5, rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(ny) + 0.5]
    ops = [np.ascontiguousarray(rng.standard_normal(nx * ny)) for _ in range(4)]
    coeffs = coeffs_for(k)
    want = chain_np(*tabs, *[o.reshape(ny, nx) for o in ops], coeffs)
    rc, outs = run_emu(emu, variant, k, 0, 1, nx, ny, [tabs[0].ctypes.data, tabs[1].ctypes.data],
                       [tabs[2].ctypes.data, tabs[3].ctypes.data], ops, coeffs, rows, [True] * k)
    assert rc == 0
    for l in range(k):
        assert np.array_equal(outs[l], want[l].ravel()), "level %d" % (l + 1)


def deep_halo(field, i0, j0, nx, ny, g, g2):
    """[S | N | W | E] deep halo of the block [j0, j0+ny) x [i0, i0+nx) of a periodic global field
    (layout of b200_deep_halo_exchange: S/N g rows x nx; W/E (ny+2g) rows x g2, corners inside W/E)."""
    NY, NX = field.shape
    rows = lambda a, b: np.arange(a, b) % NY
    cols = lambda a, b: np.arange(a, b) % NX
    S = field[np.ix_(rows(j0 - g, j0), cols(i0, i0 + nx))]
    N = field[np.ix_(rows(j0 + ny, j0 + ny + g), cols(i0, i0 + nx))]
    W = field[np.ix_(rows(j0 - g, j0 + ny + g), cols(i0 - g2, i0))]
    E = field[np.ix_(rows(j0 - g, j0 + ny + g), cols(i0 + nx, i0 + nx + g2))]
    return np.ascontiguousarray(np.concatenate([S.ravel(), N.ravel(), W.ravel(), E.ravel()]))


@pytest.mark.parametrize("block", [(0, 0), (1, 1)], ids=["block00", "block11"])
@pytest.mark.parametrize("k", [2, 4, 5])
@pytest.mark.parametrize("variant", [0, 1], ids=["march", "quad"])
def test_chain_kernels_halo_flavour_bit_exact(emu, block, k, variant):
    """A 2 x 2 block decomposition of a periodic field: every block, fed with deep halos and the
    coefficient tables extended by the global periodic index, must reproduce the global result."""
    nx, ny, g, g2, M, rows = 132, 20, 6, 6, 16, 6
    if k > 4:
        g2 = 8  # the quad kernel needs 4 * ceil(K/4) deep-halo columns
    NX, NY = 2 * nx, 2 * ny
    rng = np.random.default_rng(100 + k)
    T = [rng.random(NX) + 0.5, rng.random(NX) + 0.5, rng.random(NY) + 0.5, rng.random(NY) + 0.5]
    G = [rng.standard_normal((NY, NX)) for _ in range(4)]
    coeffs = coeffs_for(k)
    want = chain_np(*T, *G, coeffs)
    bi, bj = block
    i0, j0 = bi * nx, bj * ny
    ext_x = [np.ascontiguousarray(t[np.arange(i0 - M, i0 + nx + M) % NX]) for t in T[:2]]
    ext_y = [np.ascontiguousarray(t[np.arange(j0 - M, j0 + ny + M) % NY]) for t in T[2:]]
    cx = [t.ctypes.data + 8 * M for t in ext_x]
    cy = [t.ctypes.data + 8 * M for t in ext_y]
    ops = [np.ascontiguousarray(f[j0:j0 + ny, i0:i0 + nx].ravel()) for f in G]
    halos = [deep_halo(f, i0, j0, nx, ny, g, g2) for f in G]
    for lazy in (0, 1):
        rc, outs = run_emu(emu, variant, k, 0, lazy, nx, ny, cx, cy, ops, coeffs, rows, [True] * k, halos, g, g2)
        assert rc == 0
        for l in range(k):
            assert np.array_equal(outs[l], want[l][j0:j0 + ny, i0:i0 + nx].ravel()), "level %d" % (l + 1)


def test_fma_flavour_consistent_between_kernels_and_close_to_exact(emu):
    nx, ny, rows, k = 256, 20, 8, 4
    rng = np.random.default_rng(77)
    tabs = [rng.random(nx) + 0.5, rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(ny) + 0.5]
    ops = [np.ascontiguousarray(rng.standard_normal(nx * ny)) for _ in range(4)]
    coeffs = coeffs_for(k)
    cx = [tabs[0].ctypes.data, tabs[1].ctypes.data]
    cy = [tabs[2].ctypes.data, tabs[3].ctypes.data]
    want = chain_np(*tabs, *[o.reshape(ny, nx) for o in ops], coeffs)
    _, m = run_emu(emu, 0, k, 1, 0, nx, ny, cx, cy, ops, coeffs, rows, [True] * k)
    _, q = run_emu(emu, 1, k, 1, 0, nx, ny, cx, cy, ops, coeffs, rows, [True] * k)
    for l in range(k):
        assert np.array_equal(m[l], q[l])
        rel = np.linalg.norm(m[l] - want[l].ravel()) / np.linalg.norm(want[l])
        assert 0 < rel < 1e-13  # contracted arithmetic differs, by rounding only


@pytest.mark.parametrize("k", [3, 4])
@pytest.mark.parametrize("halo", [False, True], ids=["wrap", "halo"])
def test_uniform_coefficient_flavour_bit_exact(emu, k, halo):
    """b200_stencil_geom.uniform: coefficients from kernel parameters, centre coefficient summed on the
    host.  Must equal both the numpy restatement and the table-driven flavour bit for bit."""
    nx, ny, rows, g, g2, M = 132, 20, 6, 6, 6, 16
    u4 = [1.7, 1.7, 0.45, 0.45]
    rng = np.random.default_rng(9 + k)
    coeffs = coeffs_for(k)
    if not halo:
        tabs = [np.full(nx, u4[0]), np.full(nx, u4[1]), np.full(ny, u4[2]), np.full(ny, u4[3])]
        ops = [np.ascontiguousarray(rng.standard_normal(nx * ny)) for _ in range(4)]
        want = [w.ravel() for w in chain_np(*tabs, *[o.reshape(ny, nx) for o in ops], coeffs)]
        cx = [tabs[0].ctypes.data, tabs[1].ctypes.data]
        cy = [tabs[2].ctypes.data, tabs[3].ctypes.data]
        halos = None
    else:
        NX, NY = 2 * nx, 2 * ny
        T = [np.full(NX, u4[0]), np.full(NX, u4[1]), np.full(NY, u4[2]), np.full(NY, u4[3])]
        G = [rng.standard_normal((NY, NX)) for _ in range(4)]
        i0, j0 = nx, 0
        want = [w[j0:j0 + ny, i0:i0 + nx].ravel() for w in chain_np(*T, *G, coeffs)]
        ext = [np.full(nx + 2 * M, u4[0]), np.full(nx + 2 * M, u4[1]), np.full(ny + 2 * M, u4[2]), np.full(ny + 2 * M, u4[3])]
        cx = [ext[0].ctypes.data + 8 * M, ext[1].ctypes.data + 8 * M]
        cy = [ext[2].ctypes.data + 8 * M, ext[3].ctypes.data + 8 * M]
        ops = [np.ascontiguousarray(f[j0:j0 + ny, i0:i0 + nx].ravel()) for f in G]
        halos = [deep_halo(f, i0, j0, nx, ny, g, g2) for f in G]
    rc, tab = run_emu(emu, 0, k, 0, 0, nx, ny, cx, cy, ops, coeffs, rows, [True] * k, halos, g, g2)
    assert rc == 0
    rc, uni = run_emu(emu, 0, k, 0, 1, nx, ny, cx, cy, ops, coeffs, rows, [True] * k, halos, g, g2, uniform=u4)
    assert rc == 0
    for l in range(k):
        assert np.array_equal(uni[l], want[l]), "level %d" % (l + 1)
        assert np.array_equal(uni[l], tab[l]), "level %d" % (l + 1)


def test_quad_kernel_rejects_unsupported_shapes(emu):
    rng = np.random.default_rng(1)
    nx, ny = 131, 16  # odd nx: no 16-byte row alignment
    tabs = [rng.random(nx), rng.random(nx), rng.random(ny), rng.random(ny)]
    ops = [rng.standard_normal(nx * ny) for _ in range(4)]
    rc, _ = run_emu(emu, 1, 4, 0, 0, nx, ny, [tabs[0].ctypes.data, tabs[1].ctypes.data],
                    [tabs[2].ctypes.data, tabs[3].ctypes.data], ops, coeffs_for(4), 8, [True] * 4)
    assert rc == -1


# ------------------------------------------------------------------ the fused one-stage kernels
# (csrc/stage_kernels.cuh: k_stage_march / k_stage_generic / k_stage_ring), against the oracle's
# orc_laplacian + orc_linear_combination + orc_wsqrsum -- the CPU twin of
# test_kernels_gpu.py::test_stencil_lincomb_bit_exact / test_fused_wrms_matches_separate_norm.
STAGE_PATTERNS = {
    "rhs": [2], "ssp_stage": [1, 2], "sts_embed": [0, 1, 0, 2], "sts_stage": [2, 0, 0, 1, 0],
    "general3": [0, 2, 0],  # not a compiled pattern -> runtime-pattern instantiation
}


def emu_stage(emu, nx, ny, tabs, halos, x, coeffs, srcs, vecs, region=0, rows=8, wrms_w=None, force_generic=0,
              z=None, f=None):
    n = nx * ny
    z = np.full(n, np.nan) if z is None else z
    f = np.full(n, np.nan) if f is None else f
    send = [np.zeros(ny), np.zeros(ny), np.zeros(nx), np.zeros(nx)]
    res = np.zeros(1)
    nt = len(srcs)
    varr = (ctypes.c_void_p * nt)(*[v.ctypes.data if v is not None else None for v in vecs])
    hp = [P(h) if h is not None else None for h in (halos or [None] * 4)]
    emu.emu_stencil_lincomb.restype = ctypes.c_int
    rc = emu.emu_stencil_lincomb(
        ctypes.c_int64(nx), ctypes.c_int64(ny), P(tabs[0]), P(tabs[1]), P(tabs[2]), P(tabs[3]), *hp, P(x), nt,
        (ctypes.c_double * nt)(*coeffs), (ctypes.c_int * nt)(*srcs), varr, P(z), P(f),
        *([P(s) for s in send] if region == 0 else [None] * 4),
        P(wrms_w) if wrms_w is not None else None, P(res) if wrms_w is not None else None, rows, region, force_generic)
    assert rc == 0
    return z, f, send, float(res[0])


def oracle_stage(orc, g, x, coeffs, srcs, vecs, halos):
    n = x.size
    L = np.zeros(n)
    hp = [P(h) if h is not None else None for h in (halos or [None] * 4)]
    orc.orc_laplacian(ctypes.byref(g), P(x), P(L), *hp)
    terms = [L if s == 2 else (x if s == 1 else v) for s, v in zip(srcs, vecs)]
    z = np.zeros(n)
    arr = (ctypes.c_void_p * len(terms))(*[t.ctypes.data for t in terms])
    orc.orc_linear_combination(len(terms), (ctypes.c_double * len(terms))(*coeffs), arr, P(z), ctypes.c_int64(n))
    return z, L


@pytest.mark.parametrize("size", [(64, 20), (514, 9), (4, 5), (75, 11)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("pat", sorted(STAGE_PATTERNS))
@pytest.mark.parametrize("halo_mode", ["wrap", "all"])
def test_stage_kernels_bit_exact(emu, orc, size, pat, halo_mode):
    nx, ny = size
    srcs = STAGE_PATTERNS[pat]
    rng = np.random.default_rng(nx * 131 + ny + len(srcs))
    x = rng.standard_normal(nx * ny)
    vecs = [rng.standard_normal(nx * ny) if s == 0 else None for s in srcs]
    coeffs = list(rng.standard_normal(len(srcs)))
    halos = None
    if halo_mode == "all":
        halos = [rng.standard_normal(ny), rng.standard_normal(ny), rng.standard_normal(nx), rng.standard_normal(nx)]
    g = make_grid(nx, ny, kx=1.0, ky=0.5, inhom=True)
    tabs = [np.zeros(nx), np.zeros(nx), np.zeros(ny), np.zeros(ny)]
    orc.orc_coeff_tables(ctypes.byref(g), *[P(t) for t in tabs])
    want_z, want_L = oracle_stage(orc, g, x, coeffs, srcs, vecs, halos)
    z, f, send, _ = emu_stage(emu, nx, ny, tabs, halos, x, coeffs, srcs, vecs)
    assert np.array_equal(z, want_z) and np.array_equal(f, want_L)
    Z = want_z.reshape(ny, nx)
    assert np.array_equal(send[0], Z[:, 0]) and np.array_equal(send[1], Z[:, -1])  # buffers.cpp:20-43
    assert np.array_equal(send[2], Z[0, :]) and np.array_equal(send[3], Z[-1, :])
    if nx % 2 == 0:  # the any-width kernel must agree with the fast path
        zg, fg, _, _ = emu_stage(emu, nx, ny, tabs, halos, x, coeffs, srcs, vecs, force_generic=1)
        assert np.array_equal(zg, want_z) and np.array_equal(fg, want_L)
    if nx >= 4 and ny >= 4:  # interior (region 2) + ring (region 1) tile the sub-domain exactly
        z2, f2, _, _ = emu_stage(emu, nx, ny, tabs, halos, x, coeffs, srcs, vecs, region=2)
        inter = z2.reshape(ny, nx)
        assert np.all(np.isnan(inter[0, :])) and np.all(np.isnan(inter[:, 0])) and np.all(np.isnan(inter[-1, :])) \
            and np.all(np.isnan(inter[:, -1]))
        emu_stage(emu, nx, ny, tabs, halos, x, coeffs, srcs, vecs, region=1, z=z2, f=f2)
        assert np.array_equal(z2, want_z) and np.array_equal(f2, want_L)


@pytest.mark.parametrize("size", [(64, 20), (1024, 12), (75, 11)], ids=lambda s: "%dx%d" % s)
def test_stage_kernel_fused_wrms(emu, orc, size):
    """Fused sum((z*w)^2): deterministic block tree + last-ticket finish (grid_finish) on the emulator."""
    nx, ny = size
    rng = np.random.default_rng(99)
    x = rng.standard_normal(nx * ny)
    yn, fn, w = rng.standard_normal(nx * ny), rng.standard_normal(nx * ny), rng.random(nx * ny) + 0.1
    srcs, coeffs = [0, 1, 0, 2], [0.8, -0.8, 0.4e-3, 0.4e-3]
    g = make_grid(nx, ny)
    tabs = [np.zeros(nx), np.zeros(nx), np.zeros(ny), np.zeros(ny)]
    orc.orc_coeff_tables(ctypes.byref(g), *[P(t) for t in tabs])
    want_z, _ = oracle_stage(orc, g, x, coeffs, srcs, [yn, None, fn, None], None)
    z, _, _, res = emu_stage(emu, nx, ny, tabs, None, x, coeffs, srcs, [yn, None, fn, None], wrms_w=w, rows=4)
    assert np.array_equal(z, want_z)
    want = orc.orc_wsqrsum(P(want_z), P(w), ctypes.c_int64(nx * ny))
    assert res == pytest.approx(want, rel=1e-13)


@pytest.mark.parametrize("size", [(64, 20), (1024, 12), (75, 11), (32, 32)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("pat", ["sts_embed", "general3"])
def test_stage_kernel_fused_next_error_weights(emu, orc, size, pat):
    """StageArgs::ewt_out: the closing stage of an adaptive step also writes ewt = 1/(rtol |x| + atol) of its stencil input
    (bit for bit the N_VAbs / N_VScale / N_VAddConst / N_VInv sequence of arkEwtSetSS, arkode.c:2932-2944 = orc_ewt_ss) and
    reduces sum (x ewt)^2 next to sum (z w)^2 (grid_finish2: two sums, one ticket); z itself is untouched."""
    nx, ny = size
    n = nx * ny
    rng = np.random.default_rng(5 + nx)
    x = rng.standard_normal(n)
    x[::13] = 0.0
    srcs = STAGE_PATTERNS[pat]
    vecs = [rng.standard_normal(n) if s == 0 else None for s in srcs]
    coeffs = list(rng.standard_normal(len(srcs)))
    w = rng.random(n) + 0.1
    rtol, atol = 1e-4, 1e-11
    g = make_grid(nx, ny)
    tabs = [np.zeros(nx), np.zeros(nx), np.zeros(ny), np.zeros(ny)]
    orc.orc_coeff_tables(ctypes.byref(g), *[P(t) for t in tabs])
    want_z, _ = oracle_stage(orc, g, x, coeffs, srcs, vecs, None)
    want_e, tmp = np.empty(n), np.empty(n)
    orc.orc_ewt_ss(P(x), ctypes.c_double(rtol), ctypes.c_double(atol), P(tmp), P(want_e), ctypes.c_int64(n))
    for generic in ((0, 1) if nx % 2 == 0 else (0,)):
        e, r2 = np.full(n, np.nan), np.zeros(1)
        emu.emu_set_next_ewt(P(e), ctypes.c_double(rtol), ctypes.c_double(atol), P(r2))
        z, _, _, res = emu_stage(emu, nx, ny, tabs, None, x, coeffs, srcs, vecs, wrms_w=w, rows=4, force_generic=generic)
        assert np.array_equal(z, want_z) and np.array_equal(e, want_e)
        assert res == pytest.approx(orc.orc_wsqrsum(P(want_z), P(w), ctypes.c_int64(n)), rel=1e-13)
        assert float(r2[0]) == pytest.approx(orc.orc_wsqrsum(P(x), P(want_e), ctypes.c_int64(n)), rel=1e-13)


# ------------------------------------------------- the vector / halo / Jacobi / adr kernels on the emulator
# (csrc/vector_kernels.cuh, halo_kernels.cuh, adr_kernels.cuh; entry points tests/emu/emu_kernels.cpp)
def _aligned(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ctypes.data % 16 == 0
    return a


@pytest.mark.parametrize("n", [1, 2, 3, 1025, 4099])
def test_elementwise_kernels_bit_exact(emu, orc, n):
    rng = np.random.default_rng(n)
    x, y = _aligned(rng.standard_normal(n)), _aligned(rng.standard_normal(n) + 3.0)
    N = ctypes.c_int64(n)

    def run(op, a=0.0, b=0.0, terms=None):
        z = _aligned(np.full(n, np.nan))
        nt, cf, vv = 0, None, None
        if terms:
            nt = len(terms)
            cf = (ctypes.c_double * nt)(*[t[0] for t in terms])
            vv = (ctypes.c_void_p * nt)(*[t[1].ctypes.data for t in terms])
        assert emu.emu_elementwise(op, N, P(x), P(y), ctypes.c_double(a), ctypes.c_double(b), nt, cf, vv, P(z), 3) == 0
        return z

    want = np.zeros(n)
    vs = [_aligned(rng.standard_normal(n)) for _ in range(5)]
    cs = list(rng.standard_normal(5))
    arr = (ctypes.c_void_p * 5)(*[v.ctypes.data for v in vs])
    orc.orc_linear_combination(5, (ctypes.c_double * 5)(*cs), arr, P(want), N)
    assert np.array_equal(run(0, terms=list(zip(cs, vs))), want)
    assert np.array_equal(run(1, a=0.75), 0.75 * (x + y)) and np.array_equal(run(2, a=-1.5), -1.5 * (x - y))
    assert np.array_equal(run(3, a=2.5), np.full(n, 2.5))
    for op, fn in ((4, orc.orc_prod), (5, orc.orc_div)):
        fn(P(x), P(y), P(want), N)
        assert np.array_equal(run(op), want)
    for op, fn in ((6, orc.orc_abs), (7, orc.orc_inv)):
        fn(P(x), P(want), N)
        assert np.array_equal(run(op), want)
    orc.orc_addconst(P(x), ctypes.c_double(0.3), P(want), N)
    assert np.array_equal(run(8, b=0.3), want)
    tmp = np.zeros(n)
    orc.orc_ewt_ss(P(x), ctypes.c_double(1e-5), ctypes.c_double(1e-10), P(tmp), P(want), N)  # arkode.c:2932-2944
    assert np.array_equal(run(9, a=1e-5, b=1e-10), want)


@pytest.mark.parametrize("n", [1, 2, 5, 10001])
@pytest.mark.parametrize("blocks", [1, 5])
def test_reduction_kernels(emu, orc, n, blocks):
    """Deterministic tree reductions: equal to the oracle's sequential sums to 1e-13 (summation order), max / min
    exactly; the same input reduced twice gives the same bits."""
    rng = np.random.default_rng(n + blocks)
    x, w = _aligned(rng.standard_normal(n)), _aligned(rng.random(n) + 0.1)
    emu.emu_reduce.restype = ctypes.c_double
    N = ctypes.c_int64(n)
    got = [emu.emu_reduce(k, N, P(x), P(w), blocks) for k in range(5)]
    assert got[0] == pytest.approx(orc.orc_dot(P(x), P(w), N), rel=1e-13, abs=1e-15)
    assert got[1] == pytest.approx(orc.orc_wsqrsum(P(x), P(w), N), rel=1e-13)
    assert got[2] == orc.orc_maxnorm(P(x), N)
    assert got[3] == orc.orc_min(P(x), N)
    assert got[4] == pytest.approx(orc.orc_l1norm(P(x), N), rel=1e-13)
    assert got == [emu.emu_reduce(k, N, P(x), P(w), blocks) for k in range(5)]


def test_pack_jacobi_and_strip_kernels(emu, orc):
    nx, ny, g, g2 = 70, 37, 6, 6
    rng = np.random.default_rng(4)
    u = _aligned(rng.standard_normal(nx * ny))
    grid = make_grid(nx, ny)
    want = [np.zeros(ny), np.zeros(ny), np.zeros(nx), np.zeros(nx)]
    orc.orc_pack(ctypes.byref(grid), P(u), *[P(w) for w in want])  # buffers.cpp:20-43
    got = [np.full(ny, np.nan), np.full(ny, np.nan), np.full(nx, np.nan), np.full(nx, np.nan)]
    emu.emu_pack(P(u), ctypes.c_int64(nx), ctypes.c_int64(ny), *[P(b) for b in got])
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # Jacobi diagonal, preconditioner_jacobi.cpp:41-46, from given face tables
    t = [rng.random(nx) + 0.5, rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(ny) + 0.5]
    gamma = 0.0123
    diag = np.full(nx * ny, np.nan)
    emu.emu_jacobi(ctypes.c_int64(nx), ctypes.c_int64(ny), *[P(a) for a in t], ctypes.c_double(gamma), P(diag))
    d = -((t[0] + t[1])[None, :] + (t[2] + t[3])[:, None])
    assert np.array_equal(diag.reshape(ny, nx), 1.0 / (1.0 - gamma * d))
    # W / E strips of the deep halo: columns [0, g2) and [nx-g2, nx) of rows -g .. ny+g-1, the rows outside the
    # field taken from the S / N halo blocks (corners travel with the second exchange phase)
    F = u.reshape(ny, nx)
    S, Nn = rng.standard_normal((g, nx)), rng.standard_normal((g, nx))
    halo = _aligned(np.concatenate([S.ravel(), Nn.ravel()]))
    ws, es = np.full((ny + 2 * g) * g2, np.nan), np.full((ny + 2 * g) * g2, np.nan)
    emu.emu_pack_strips(P(u), P(halo), ctypes.c_int64(nx), ctypes.c_int64(ny), g, g2, P(ws), P(es))
    tall = np.vstack([S, F, Nn])
    assert np.array_equal(ws.reshape(ny + 2 * g, g2), tall[:, :g2])
    assert np.array_equal(es.reshape(ny + 2 * g, g2), tall[:, nx - g2:])


@pytest.mark.parametrize("size", [(16, 12), (300, 9), (101, 33)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5, 6, 7])
def test_adr_kernel_every_composite_bit_exact(emu, orc, size, mode):
    """k_adr_march<MODE>: z = c0*v0 + c1*y + c2*F_mode(y) + c3*v3 in one pass vs the oracle's callbacks summed in
    the reference's order (adr/advection_diffusion_reaction_2d.cpp:1406-1520, :1602-1619), and the plain RHS."""
    from conftest import OrcAdr

    class AdrParams(ctypes.Structure):  # b200_adr_params, include/b200_sts.h
        _fields_ = OrcAdr._fields_

    nx, ny = size
    vals = (nx, ny, 1.0 / nx, 1.0 / ny, -0.5, 1.0, 0.4, 0.7, 1e-2, 1.3, 1.0)
    p, bp = OrcAdr(*vals), AdrParams(*vals)
    n = 2 * nx * ny
    y = _aligned(np.zeros(n))
    orc.orc_adr_ic(ctypes.byref(p), ctypes.c_double(0.0), ctypes.c_double(0.0), P(y))
    y += 0.01 * np.random.default_rng(mode).standard_normal(n)
    parts = {}
    for bit, fn in ((1, orc.orc_adr_advection), (2, orc.orc_adr_diffusion), (4, orc.orc_adr_reaction)):
        parts[bit] = np.zeros(n)
        fn(ctypes.byref(p), P(y), P(parts[bit]))
    F = None
    for bit in (1, 2, 4):
        if mode & bit:
            F = parts[bit].copy() if F is None else F + parts[bit]
    rng = np.random.default_rng(100 + mode)
    v0, v3 = _aligned(rng.standard_normal(n)), _aligned(rng.standard_normal(n))
    c = [0.7, -0.25, 3e-3, 1.5]
    want = ((c[0] * v0 + c[1] * y) + c[2] * F) + c[3] * v3
    z, f = _aligned(np.full(n, np.nan)), _aligned(np.full(n, np.nan))
    vv = (ctypes.c_void_p * 4)(v0.ctypes.data, None, None, v3.ctypes.data)
    rc = emu.emu_adr(ctypes.byref(bp), mode, P(y), 4, (ctypes.c_double * 4)(*c), (ctypes.c_int * 4)(0, 1, 2, 0), vv,
                     P(z), P(f), 8)
    assert rc == 0
    assert np.array_equal(z, want) and np.array_equal(f, F)
    f2 = _aligned(np.full(n, np.nan))
    assert emu.emu_adr(ctypes.byref(bp), mode, P(y), 0, None, None, None, None, P(f2), 5) == 0
    assert np.array_equal(f2, F)


@pytest.mark.parametrize("size", [(64, 16, 64), (100, 21, 5), (250, 26, 8)], ids=lambda s: "%dx%d_rows%d" % s)
@pytest.mark.parametrize("k", [2, 3, 4, 6])
def test_adr_chain_kernel_equals_single_stage_launches(emu, orc, size, k):
    """k_adr_chain (csrc/adr_chain.cuh): K stages of the adr diffusion partition in one pass must be bit-identical
    to K launches of k_adr_march<2> with the LSRKStep stage pattern [F(y), z_{j-2}, yn, y, fn] (which is pinned
    against the oracle above), including the periodic wrap in both directions and partial windows / row blocks."""
    from conftest import OrcAdr

    class AdrParams(ctypes.Structure):
        _fields_ = OrcAdr._fields_

    nx, ny, rows = size
    bp = AdrParams(nx, ny, 1.0 / nx, 1.0 / ny, -0.5, 1.0, 0.4, 0.7, 2e-2, 1.3, 1.0)
    n = 2 * nx * ny
    rng = np.random.default_rng(nx + 7 * ny + k)
    x, p2, yn, fn = (_aligned(rng.standard_normal(n)) for _ in range(4))
    coeffs = [[1e-4 * (l + 1), -0.3 + 0.1 * l, 0.2, 1.1 - 0.05 * l, -2e-4] for l in range(k)]
    # reference: one fused launch per stage
    want, prev, cur = [], p2, x
    for l in range(k):
        z = _aligned(np.full(n, np.nan))
        vv = (ctypes.c_void_p * 5)(None, prev.ctypes.data, yn.ctypes.data, None, fn.ctypes.data)
        assert emu.emu_adr(ctypes.byref(bp), 2, P(cur), 5, (ctypes.c_double * 5)(*coeffs[l]),
                           (ctypes.c_int * 5)(2, 0, 0, 1, 0), vv, P(z), None, 8) == 0
        want.append(z)
        prev, cur = cur, z
    cf = np.ascontiguousarray(np.array(coeffs).ravel())
    for lazy, store in ((0, [True] * k), (1, [l >= k - 2 for l in range(k)])):
        outs = [_aligned(np.full(n, np.nan)) if s else None for s in store]
        optr = (ctypes.c_void_p * k)(*[o.ctypes.data if o is not None else None for o in outs])
        assert emu.emu_adr_chain(ctypes.byref(bp), k, P(x), P(p2), P(yn), P(fn), P(cf), optr, rows, lazy) == 0
        for l in range(k):
            if store[l]:
                assert np.array_equal(outs[l], want[l]), "level %d (lazy=%d)" % (l + 1, lazy)


def _random_shapes(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        nx = int(rng.integers(64, 350)) * 2
        ny = int(rng.integers(16, 90))
        rows = int(rng.integers(1, 48))
        k = int(rng.integers(2, 7))
        out.append((nx, ny, rows, k, int(rng.integers(0, 2)), int(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("shape", _random_shapes(40, 2026), ids=lambda s: "%dx%d_rows%d_k%d_v%d_l%d" % s)
def test_chain_kernels_random_shapes_bit_exact(emu, shape):
    """Seeded random widths (even, 128..698), heights, rows per block (1..47, mostly not dividing ny), depths,
    both kernels, both cp.async completion orders: every level against the numpy restatement."""
    nx, ny, rows, k, variant, lazy = shape
    rng = np.random.default_rng(nx * 1009 + ny * 31 + rows + k)
    tabs = [rng.random(nx) + 0.5, rng.random(nx) + 0.5, rng.random(ny) + 0.5, rng.random(ny) + 0.5]
    ops = [np.ascontiguousarray(rng.standard_normal(nx * ny)) for _ in range(4)]
    coeffs = coeffs_for(k)
    want = chain_np(*tabs, *[o.reshape(ny, nx) for o in ops], coeffs)
    store = [bool(rng.integers(0, 2)) or l == k - 1 for l in range(k)]
    rc, outs = run_emu(emu, variant, k, 0, lazy, nx, ny, [tabs[0].ctypes.data, tabs[1].ctypes.data],
                       [tabs[2].ctypes.data, tabs[3].ctypes.data], ops, coeffs, rows, store)
    assert rc == 0
    for l in range(k):
        if store[l]:
            assert np.array_equal(outs[l], want[l].ravel()), "level %d" % (l + 1)


@pytest.mark.parametrize("shape", _random_shapes(16, 777), ids=lambda s: "%dx%d_rows%d_k%d_v%d_l%d" % s)
def test_chain_kernels_random_decompositions_bit_exact(emu, shape):
    """The halo flavour on a random block of a random 2 x 2 / 3 x 2 decomposition of a periodic field."""
    nx, ny, rows, k, variant, lazy = shape
    g, g2, M = 6, 6, 16
    rng = np.random.default_rng(nx + ny + rows + k)
    npx, npy = int(rng.integers(2, 4)), 2
    bi, bj = int(rng.integers(0, npx)), int(rng.integers(0, npy))
    NX, NY = npx * nx, npy * ny
    T = [rng.random(NX) + 0.5, rng.random(NX) + 0.5, rng.random(NY) + 0.5, rng.random(NY) + 0.5]
    G = [rng.standard_normal((NY, NX)) for _ in range(4)]
    coeffs = coeffs_for(k)
    want = chain_np(*T, *G, coeffs)
    i0, j0 = bi * nx, bj * ny
    ext_x = [np.ascontiguousarray(t[np.arange(i0 - M, i0 + nx + M) % NX]) for t in T[:2]]
    ext_y = [np.ascontiguousarray(t[np.arange(j0 - M, j0 + ny + M) % NY]) for t in T[2:]]
    ops = [np.ascontiguousarray(f[j0:j0 + ny, i0:i0 + nx].ravel()) for f in G]
    halos = [deep_halo(f, i0, j0, nx, ny, g, g2) for f in G]
    rc, outs = run_emu(emu, variant, k, 0, lazy, nx, ny, [t.ctypes.data + 8 * M for t in ext_x],
                       [t.ctypes.data + 8 * M for t in ext_y], ops, coeffs, rows, [True] * k, halos, g, g2)
    assert rc == 0
    for l in range(k):
        assert np.array_equal(outs[l], want[l][j0:j0 + ny, i0:i0 + nx].ravel()), "level %d" % (l + 1)


@pytest.mark.parametrize("shape", _random_shapes(12, 4242), ids=lambda s: "%dx%d_rows%d_k%d_v%d_l%d" % s)
def test_adr_chain_random_shapes_bit_exact(emu, shape):
    from conftest import OrcAdr

    class AdrParams(ctypes.Structure):
        _fields_ = OrcAdr._fields_

    nx, ny, rows, k, _, lazy = shape
    nx //= 2  # grid points (each is two doubles); 64..349, odd widths included
    bp = AdrParams(nx, ny, 1.0 / nx, 1.0 / ny, -0.5, 1.0, 0.4, 0.7, 3e-2, 1.3, 1.0)
    n = 2 * nx * ny
    rng = np.random.default_rng(nx * 17 + ny + k)
    x, p2, yn, fn = (_aligned(rng.standard_normal(n)) for _ in range(4))
    coeffs = [[1e-4 * (l + 1), -0.3 + 0.1 * l, 0.2, 1.1 - 0.05 * l, -2e-4] for l in range(k)]
    want, prev, cur = [], p2, x
    for l in range(k):
        z = _aligned(np.full(n, np.nan))
        vv = (ctypes.c_void_p * 5)(None, prev.ctypes.data, yn.ctypes.data, None, fn.ctypes.data)
        assert emu.emu_adr(ctypes.byref(bp), 2, P(cur), 5, (ctypes.c_double * 5)(*coeffs[l]),
                           (ctypes.c_int * 5)(2, 0, 0, 1, 0), vv, P(z), None, 8) == 0
        want.append(z)
        prev, cur = cur, z
    outs = [_aligned(np.full(n, np.nan)) for _ in range(k)]
    optr = (ctypes.c_void_p * k)(*[o.ctypes.data for o in outs])
    cf = np.ascontiguousarray(np.array(coeffs).ravel())
    assert emu.emu_adr_chain(ctypes.byref(bp), k, P(x), P(p2), P(yn), P(fn), P(cf), optr, rows, lazy) == 0
    for l in range(k):
        assert np.array_equal(outs[l], want[l]), "level %d" % (l + 1)


@pytest.mark.parametrize("k", [2, 4, 6])
def test_uniform_flavour_with_fma_arithmetic(emu, k):
    """The <FMA, UNI> instantiations: uniform coefficients from kernel parameters with contracted multiply-adds
    must equal the table-driven FMA launch bit for bit (same operations, same operands)."""
    nx, ny, rows = 196, 23, 7
    u4 = [0.9, 0.9, 1.6, 1.6]
    rng = np.random.default_rng(50 + k)
    tabs = [np.full(nx, u4[0]), np.full(nx, u4[1]), np.full(ny, u4[2]), np.full(ny, u4[3])]
    ops = [np.ascontiguousarray(rng.standard_normal(nx * ny)) for _ in range(4)]
    coeffs = coeffs_for(k)
    cx = [tabs[0].ctypes.data, tabs[1].ctypes.data]
    cy = [tabs[2].ctypes.data, tabs[3].ctypes.data]
    rc, tab = run_emu(emu, 0, k, 1, 0, nx, ny, cx, cy, ops, coeffs, rows, [True] * k)
    assert rc == 0
    rc, uni = run_emu(emu, 0, k, 1, 1, nx, ny, cx, cy, ops, coeffs, rows, [True] * k, uniform=u4)
    assert rc == 0
    for l in range(k):
        assert np.array_equal(uni[l], tab[l]), "level %d" % (l + 1)


# ---- peer-mapped deep-halo exchange (k_peer_exchange + peer_dst_pointers) on a simulated process grid ----------
@pytest.mark.parametrize("npx,npy,nx_loc,ny_glob", [(1, 1, 16, 12), (2, 1, 8, 9), (1, 2, 16, 17), (2, 2, 10, 21), (4, 2, 8, 19), (3, 3, 6, 31)])
@pytest.mark.parametrize("g,g2", [(2, 2), (6, 6), (3, 4)])
def test_peer_exchange_fills_every_deep_halo_of_a_process_grid(emu, npx, npy, nx_loc, ny_glob, g, g2):
    """Every rank of a periodic npx x npy grid (uneven block heights: remainder rows to the low coordinates,
    diffusion_2D.cpp:286-317) pushes its edges and corners; afterwards every rank's slot must hold the deep halo
    [S | N | W | E] of its block, i.e. the periodic continuation of the global field, corners included."""
    nx_glob = nx_loc * npx
    rng = np.random.default_rng(npx * 100 + npy * 10 + g)
    U = rng.standard_normal((ny_glob, nx_glob))
    q, r = divmod(ny_glob, npy)
    ny_loc = [q + (1 if c < r else 0) for c in range(npy)]
    js = [sum(ny_loc[:c]) for c in range(npy)]
    if min(ny_loc) < g:
        pytest.skip("block lower than the halo")
    blocks, slots = [], []
    for rank in range(npx * npy):
        cx, cy = rank // npy, rank % npy
        blocks.append(np.ascontiguousarray(U[js[cy]:js[cy] + ny_loc[cy], cx * nx_loc:(cx + 1) * nx_loc]))
        slots.append(np.full(2 * g * nx_loc + 2 * (ny_loc[cy] + 2 * g) * g2, np.nan))
    bp = (ctypes.c_void_p * len(blocks))(*[b.ctypes.data for b in blocks])
    sp = (ctypes.c_void_p * len(slots))(*[s.ctypes.data for s in slots])
    nyl = (ctypes.c_int64 * npy)(*ny_loc)
    assert emu.emu_peer_exchange_grid(npx, npy, ctypes.c_int64(nx_loc), nyl, g, g2, bp, sp) == 0
    for rank in range(npx * npy):
        cx, cy = rank // npy, rank % npy
        i0, j0, ny = cx * nx_loc, js[cy], ny_loc[cy]
        rows = lambda a, b: np.arange(j0 + a, j0 + b) % ny_glob  # noqa: E731
        cols = lambda a, b: np.arange(i0 + a, i0 + b) % nx_glob  # noqa: E731
        want = np.concatenate([
            U[np.ix_(rows(-g, 0), cols(0, nx_loc))].ravel(),
            U[np.ix_(rows(ny, ny + g), cols(0, nx_loc))].ravel(),
            U[np.ix_(rows(-g, ny + g), cols(-g2, 0))].ravel(),
            U[np.ix_(rows(-g, ny + g), cols(nx_loc, nx_loc + g2))].ravel()])
        assert np.array_equal(slots[rank], want), rank


# ---- BULK flavour of k_chain_march: the operand ring filled by cp.async.bulk + mbarrier (emulated in cuda_emu.h:
# eager = the copy lands at issue time, lazy = when somebody waits for its barrier).  Same checks as above, the switch on.
@pytest.fixture(params=[1, 2], ids=["pf", "pf_plus_1"])
def emu_bulk(emu, request):
    emu.emu_set_chain_bulk(request.param)  # 1: the plain flavour's prefetch depth, 2: one row deeper
    yield emu
    emu.emu_set_chain_bulk(0)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_rows%d" % c)
@pytest.mark.parametrize("k", [2, 3, 4])
def test_bulk_ring_periodic_wrap_bit_exact(emu_bulk, case, k):
    test_chain_kernels_periodic_wrap_bit_exact(emu_bulk, case, k, 0)


@pytest.mark.parametrize("k", [5, 6])
def test_bulk_ring_deep_levels_bit_exact(emu_bulk, k):
    test_chain_kernels_deep_levels_bit_exact(emu_bulk, k, 0)


@pytest.mark.parametrize("block", [(0, 0), (1, 1)], ids=["block00", "block11"])
@pytest.mark.parametrize("k", [2, 4, 5])
def test_bulk_ring_halo_flavour_bit_exact(emu_bulk, block, k):
    test_chain_kernels_halo_flavour_bit_exact(emu_bulk, block, k, 0)


@pytest.mark.parametrize("k", [3, 4])
@pytest.mark.parametrize("halo", [False, True], ids=["wrap", "halo"])
def test_bulk_ring_uniform_coefficients_bit_exact(emu_bulk, k, halo):
    test_uniform_coefficient_flavour_bit_exact(emu_bulk, k, halo)


@pytest.mark.parametrize("shape", [s for s in _random_shapes(40, 2026) if s[4] == 0],
                         ids=lambda s: "%dx%d_rows%d_k%d_v%d_l%d" % s)
def test_bulk_ring_random_shapes_bit_exact(emu_bulk, shape):
    test_chain_kernels_random_shapes_bit_exact(emu_bulk, shape)


@pytest.mark.parametrize("shape", [s for s in _random_shapes(16, 777) if s[4] == 0],
                         ids=lambda s: "%dx%d_rows%d_k%d_v%d_l%d" % s)
def test_bulk_ring_random_decompositions_bit_exact(emu_bulk, shape):
    test_chain_kernels_random_decompositions_bit_exact(emu_bulk, shape)
