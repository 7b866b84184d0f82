"""CPU-side checks: the C-ABI libraries load and export every declared symbol, the host mirror of
the reference's decomposition agrees with the oracle and with the C++ driver, the oracle's vector
ops honour the reference's special cases, and the product refuses to run without a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
from conftest import ROOT, P, has_gpu, make_grid

INCLUDE = os.path.join(ROOT, "include")


def declared_functions(header):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:b200|N_V)\w*)\s*\(", text))
    return sorted(n for n in names if not n.isupper())


def exported(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    return {ln.split()[-1] for ln in out.splitlines() if ln.strip()}


def test_kernel_library_exports_every_declared_symbol(b200):
    syms = exported(b200.KERNEL_LIB)
    missing = [f for f in declared_functions("b200_sts.h") if f not in syms]
    assert not missing, missing
    b200.kernel_lib()  # dlopen succeeds without a GPU and without NCCL being loaded


def test_chain_rows_are_fitted_to_whole_waves(b200):
    """Rows per block of the chain kernel on a 148-SM device (host logic only, nothing is launched): the measured choices
    of DESIGN 4.1b -- 256 rows at 16384^2 (8.0 waves), 128 at 8192^2, 32 at 4096^2, and at 2048^2 not 32 rows (320 blocks
    for 296 slots: a straggler wave) but 35 (295 blocks); tiny grids fit one wave with the fewest rows."""
    import ctypes

    lib = b200.kernel_lib()
    lib.b200_chain_rows_query.restype = ctypes.c_int
    q = lambda nx, ny, k: lib.b200_chain_rows_query(None, ctypes.c_int64(nx), ctypes.c_int64(ny), k, 148)
    assert q(16384, 16384, 4) == 256
    assert q(8192, 8192, 4) == 128
    assert q(4096, 4096, 4) == 32
    assert q(2048, 2048, 4) == 35
    assert q(128, 128, 4) == 8
    assert q(1024, 1024, 4) == 16
    slots = 2 * 148
    for n in (1536, 2048, 2304, 2560, 3072):
        r = q(n, n, 4)
        gx = ((n + 55) // 56 + 7) // 8
        blocks = gx * ((n + r - 1) // r)
        waves, rest = divmod(blocks, slots)
        assert 8 <= r <= 256
        assert waves == 0 or waves > 3 or rest == 0 or 4 * rest > slots, (n, r, blocks)  # no straggler quarter-wave
    assert q(4096, 4096, 1) == -1 and q(4096, 4096, 7) == -1


def test_sundials_library_exports_every_declared_symbol(b200):
    if not os.path.exists(b200.SUNDIALS_LIB):
        pytest.skip("libb200sts_sundials.so not built (no SUNDIALS host library)")
    syms = exported(b200.SUNDIALS_LIB)
    want = (declared_functions("nvector_b200.h") + declared_functions("b200_diffusion2d.h") +
            declared_functions("b200_adr2d.h") + declared_functions("b200_callbacks.h"))
    assert "b200_adr_f_diffusion_forcing" in want and "b200_adr_create" in want and "b200_diffusion_rhs" in want
    missing = [f for f in want if f not in syms and not f.startswith("b200_ctx") and f in want]
    missing = [f for f in missing if f.startswith(("N_V", "b200_d2d", "b200_diffusion", "b200_adr2d", "b200_adr_"))]
    assert not missing, missing
    b200.sundials_lib()


def test_kernel_library_has_sm100a_code_and_no_nccl_link_dependency(b200):
    out = subprocess.run(["cuobjdump", "-lelf", b200.KERNEL_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", b200.KERNEL_LIB], capture_output=True, text=True).stdout
    assert "nccl" not in ldd  # bound with dlopen at first multi-GPU use


def test_default_chain_kernel_uses_the_tma_unit(b200):
    """The shipped library's default instantiation of the dominant kernel -- k_chain_march<4, 4, wrap, exact, uniform,
    body, plain order, BULK> -- fills its operand ring with bulk asynchronous copies: its SASS must hold the bulk-copy,
    transaction-barrier and elect instructions (and, for the edge windows, the per-thread LDGSTS path), and no local
    memory traffic (spills)."""
    import shutil

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    fun = "_Z13k_chain_marchILi4ELi4ELb0ELb0ELb1ELb0ELb0ELb1EEv9ChainArgs"
    r = subprocess.run(["cuobjdump", "-sass", "-fun", fun, b200.KERNEL_LIB], capture_output=True, text=True)
    sass = r.stdout
    assert "Function : " + fun in sass, r.stderr[-500:]
    for mnemonic in ("UBLKCP.S.G", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "ELECT", "LDGSTS.E.BYPASS.128"):
        assert mnemonic in sass, mnemonic
    assert sass.count("UBLKCP.S.G") >= 4 * 9  # 4 operands x (3 prologue groups + 3 unrolled phases x >= 2 loops)
    assert "STL" not in sass and "LDL" not in sass
    assert sass.count("DMUL") + sass.count("DADD") > 1000 and "DFMA" not in sass  # exact arithmetic: no contraction


@pytest.mark.skipif(has_gpu(), reason="only meaningful where there is no GPU")
def test_product_fails_loudly_without_a_gpu(b200):
    lib = b200.kernel_lib()
    h = ctypes.c_void_p()
    rc = lib.b200_ctx_create(0, None, ctypes.byref(h))
    assert rc != 0 and lib.b200_last_error()
    if os.path.exists(b200.DRIVER_BIN):
        r = subprocess.run([b200.DRIVER_BIN, "--nx", "32", "--ny", "32", "--integrator", "rkc"], capture_output=True, text=True)
        assert r.returncode != 0


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may mention it."""
    bad = []
    for base in ("ceda-demonstrations_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "_sundials" in dp or dp.endswith(("lib", "bin", "__pycache__")):
                continue
            for fn in files:
                if fn.endswith((".py", ".cpp", ".cu", ".h", ".hpp")):
                    txt = open(os.path.join(dp, fn)).read()
                    if re.search(r"liboracle|sts_oracle|orc_|oracle/", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_product_is_never_built_with_the_host_emulator(b200):
    """tests/emu compiles the kernel headers with -DB200_HOST_EMU to run them on CPU threads (test
    infrastructure).  The product must not: the define appears in no product build recipe, the only places
    that react to it are the guarded include in kernel_prims.cuh and the launch / runtime switch at the top of
    b200_kernels.cu (klaunch, cuda_runtime_emu.h), and the shipped libraries carry neither the emulator's entry
    points nor its runtime."""
    for recipe in ("Makefile", "__graft_entry__.py", os.path.join("scripts", "sundials_host.mk")):
        assert "B200_HOST_EMU" not in open(os.path.join(ROOT, recipe)).read(), recipe
    csrc = os.path.join(ROOT, "ceda-demonstrations_b200", "csrc")
    users = sorted(fn for fn in os.listdir(csrc) if "#ifdef B200_HOST_EMU" in open(os.path.join(csrc, fn)).read())
    assert users == ["b200_kernels.cu", "kernel_prims.cuh"], users
    for lib in (b200.KERNEL_LIB, b200.SUNDIALS_LIB):
        if os.path.exists(lib):
            syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
            assert "emu_" not in syms and "cuda_emu" not in syms and "_ZN3emu" not in syms, lib
            allsyms = subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
            assert "_ZN3emu" not in allsyms and "emu_switch" not in allsyms, lib


# ---- decomposition: diffusion_2D.cpp:243-317 ----------------------------------------------------
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 6, 8, 12, 16])
@pytest.mark.parametrize("nx,ny", [(128, 128), (101, 77), (16384, 16384), (10, 7)])
def test_block_decomposition_matches_oracle_and_driver(b200, orc, nranks, nx, ny):
    dims = (ctypes.c_int * 2)()
    orc.orc_dims_create(nranks, dims)
    covered = np.zeros((ny, nx), dtype=np.int32)
    for rank in range(nranks):
        d = b200.block_decomposition(nx, ny, rank, nranks)
        assert (d["npx"], d["npy"]) == (dims[0], dims[1])
        s, c = ctypes.c_int64(), ctypes.c_int64()
        orc.orc_decompose(ctypes.c_int64(nx), d["npx"], d["idx"], ctypes.byref(s), ctypes.byref(c))
        assert (d["is"], d["nx_loc"]) == (s.value, c.value)
        orc.orc_decompose(ctypes.c_int64(ny), d["npy"], d["idy"], ctypes.byref(s), ctypes.byref(c))
        assert (d["js"], d["ny_loc"]) == (s.value, c.value)
        covered[d["js"] : d["js"] + d["ny_loc"], d["is"] : d["is"] + d["nx_loc"]] += 1
        # periodic neighbours are mutual
        for a, bk in (("ipW", "ipE"), ("ipS", "ipN")):
            nb = b200.block_decomposition(nx, ny, d[a], nranks)
            assert nb[bk] == rank
        if os.path.exists(b200.SUNDIALS_LIB):
            lib = b200.sundials_lib()
            out = [ctypes.c_int64() for _ in range(4)]
            px, py = ctypes.c_int(), ctypes.c_int()
            rc = lib.b200_d2d_local_extent(ctypes.c_int64(nx), ctypes.c_int64(ny), rank, nranks, 0, 0,
                                           *[ctypes.byref(o) for o in out], ctypes.byref(px), ctypes.byref(py))
            assert rc == 0
            assert [o.value for o in out] == [d["is"], d["nx_loc"], d["js"], d["ny_loc"]]
            assert (px.value, py.value) == (d["npx"], d["npy"])
    assert np.all(covered == 1)


def test_mpi_dims_create_known_values(orc):
    want = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2), 16: (4, 4), 64: (8, 8), 6: (3, 2), 7: (7, 1)}
    dims = (ctypes.c_int * 2)()
    for n, d in want.items():
        orc.orc_dims_create(n, dims)
        assert (dims[0], dims[1]) == d


# ---- oracle vector ops: the special cases of N_VLinearSum (nvector_parallel.c:424-517) ------------
def test_oracle_linear_sum_special_cases(orc):
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(257), rng.standard_normal(257)
    cases = [(1.0, 1.0), (1.0, -1.0), (-1.0, 1.0), (1.0, 0.37), (0.37, 1.0), (-1.0, 0.37), (0.37, -1.0),
             (0.37, 0.37), (0.37, -0.37), (0.37, 2.5)]
    for a, b in cases:
        z = np.zeros_like(x)
        orc.orc_linear_sum(ctypes.c_double(a), P(x), ctypes.c_double(b), P(y), P(z), ctypes.c_int64(x.size))
        if a == b and abs(a) != 1.0:
            want = a * (x + y)
        elif a == -b and abs(a) != 1.0:
            want = a * (x - y)
        else:
            want = (a * x) + (b * y)
        assert np.array_equal(z, want), (a, b)
    # in-place axpy forms
    z = y.copy()
    orc.orc_linear_sum(ctypes.c_double(0.37), P(x), ctypes.c_double(1.0), P(z), P(z), ctypes.c_int64(x.size))
    assert np.array_equal(z, y + 0.37 * x)


def test_oracle_linear_combination_is_left_to_right(orc):
    rng = np.random.default_rng(3)
    X = [rng.standard_normal(100) for _ in range(5)]
    c = [1e-4, -0.31, 0.2, 1.11, -3e-5]
    z = np.zeros(100)
    arr = (ctypes.c_void_p * 5)(*[v.ctypes.data for v in X])
    orc.orc_linear_combination(5, (ctypes.c_double * 5)(*c), arr, P(z), ctypes.c_int64(100))
    want = c[0] * X[0]
    for k in range(1, 5):
        want = want + c[k] * X[k]
    assert np.array_equal(z, want)


def test_oracle_laplacian_halo_equals_periodic_wrap(orc):
    """One periodic rank: feeding the rank's own opposite edges as halos must equal the wrap path."""
    nx, ny = 23, 17
    g = make_grid(nx, ny, kx=0.9, ky=1.7, inhom=True)
    rng = np.random.default_rng(11)
    u = rng.standard_normal(nx * ny)
    f1, f2 = np.zeros(nx * ny), np.zeros(nx * ny)
    orc.orc_laplacian(ctypes.byref(g), P(u), P(f1), None, None, None, None)
    W, E, S, N = np.zeros(ny), np.zeros(ny), np.zeros(nx), np.zeros(nx)
    orc.orc_pack(ctypes.byref(g), P(u), P(W), P(E), P(S), P(N))
    # what I send west arrives as my east neighbour's "from west"... with one rank: Wrecv = Esend
    orc.orc_laplacian(ctypes.byref(g), P(u), P(f2), P(E), P(W), P(N), P(S))
    assert np.array_equal(f1, f2)


def test_python_and_header_struct_sizes_agree(b200, tmp_path):
    """ctypes mirrors must match the C structs: sizes and the offsets of the last fields as gcc lays the
    headers out."""
    import subprocess

    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "b200_sts.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(b200_stencil_geom), offsetof(b200_stencil_geom, uniform),'
        ' offsetof(b200_stencil_geom, u_cyn), sizeof(b200_stage_extras), sizeof(b200_adr_params));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    geom, off_uniform, off_ucyn, extras, adr = (int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True,
                                                                                check=True).stdout.split())
    assert ctypes.sizeof(b200.StencilGeom) == geom
    assert b200.StencilGeom.uniform.offset == off_uniform and b200.StencilGeom.u_cyn.offset == off_ucyn
    assert ctypes.sizeof(b200.StageExtras) == extras
    assert ctypes.sizeof(b200.AdrParams) == adr