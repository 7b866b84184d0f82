"""The whole product stack on a CPU (no GPU needed): unchanged ARKODE -> N_Vector_B200 (lazy stage fusion, pending-stage
chains, fused WRMS) -> problem layers (diffusion_2D / adr callbacks) -> the C-ABI of csrc/b200_kernels.cu -> the
kernel SOURCES executed by the host emulator (tests/emu: one fiber per CUDA thread, emulated CUDA runtime).

This is test infrastructure (tests/emu/Makefile `fullstack`; built lazily here): it exists so that the host logic
above the C-ABI -- which only ever runs against a device in the product -- is covered by the `-m "not gpu"` suite
against the SAME reference fixtures the GPU suite uses (tests/golden, written by the unmodified reference build),
with the same bars: equal integrator statistics, fixed-step states equal to every printed digit, adaptive states
within 1e-10 relative L2 (or the reference's own np=1 / np=4 spread).  The emulated library has its own file name,
binds -Bsymbolic and is loaded RTLD_LOCAL only by this module; the product never sees it.
"""
import ctypes
import importlib
import json
import os
import subprocess

import numpy as np
import pytest
from conftest import GOLDEN, ROOT, fmt16, load_golden

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libb200_fullstack_emu.so")
REL_L2_TOL = 1e-10

STAT_MAP = {"steps": "steps", "attempts": "step_attempts", "err_fails": "err_test_fails", "rhs_evals": "rhs_evals",
            "rhs_evals_e": "rhs_evals", "rhs_evals_i": "rhs_evals", "dom_eig_updates": "dom_eig_updates",
            "max_stages": "max_stages", "dee_evals": "dee_rhs_evals", "lin_iters": "lin_iters",
            "nls_iters": "nonlin_iters", "prec_solves": "prec_solves"}


@pytest.fixture(scope="module")
def emu(b200):
    """The package's ctypes classes bound to the emulated library for the duration of this module."""
    if not os.path.exists(os.path.join(ROOT, "ceda-demonstrations_b200", "_sundials", "lib", "libsundials_host.so")):
        pytest.skip("SUNDIALS host library not built")
    subprocess.run(["make", "-s", "-C", EMU_DIR, "fullstack"], check=True)
    lib = ctypes.CDLL(EMU_LIB)  # RTLD_LOCAL
    lib.b200_last_error.restype = ctypes.c_char_p
    lib.b200_launch_count.restype = ctypes.c_uint64
    lib.b200_last_chain_kernel.restype = ctypes.c_char_p
    saved = (b200._kernel_lib, b200._sundials_lib)
    b200._kernel_lib = b200._sundials_lib = lib
    yield b200
    b200._kernel_lib, b200._sundials_lib = saved


def run_d2d(b200, args):
    args = [str(a) for a in args] + ["--nout", "1", "--output", "0"]
    prob = b200.Diffusion2D(args, device=0, stream=None)
    tf = float(args[args.index("--tf") + 1]) if "--tf" in args else 1.0
    prob.evolve(tf)
    st = prob.stats()
    u = np.empty(st["nx_loc"] * st["ny_loc"])
    prob.get_state(u)
    prob.close()
    return st, u.reshape(st["ny_loc"], st["nx_loc"])


def check_stats(meta, st, implicit):
    np4 = meta.get("stats_np4", {})
    for k, v in meta["stats"].items():
        if k not in STAT_MAP or (k == "rhs_evals_e" and implicit) or (k == "rhs_evals_i" and not implicit):
            continue
        v4 = np4.get(k, v)
        d = abs(v4 - v)
        if d:  # (Krylov iteration counts move with the summation order of the dot products: never less than 0.1 % slack)
            d = max(d, int(np.ceil(1e-3 * max(v, v4))))
        assert min(v, v4) - d <= st[STAT_MAP[k]] <= max(v, v4) + d, (k, st[STAT_MAP[k]], v, v4)


D2D_CASES = ["c1_rkc_128", "rkc_fixed_aniso_inhom_96x64", "rkl_fixed_aniso_inhom_96x64", "rkl_adaptive_inhom_64",
             "rkl_internaleig_64", "rkc_odd_75x51", "ssp2_fixed_64", "ssp3_fixed_64", "ssp104_fixed_64",
             "ssp104_adaptive_64", "erk3_adaptive_64", "dirk3_pcg_64", "dirk2_pcg_inhom_noprec_48"]


@pytest.mark.parametrize("name", D2D_CASES)
def test_diffusion_fixture_parity_through_the_emulated_stack(emu, name):
    meta, ref = load_golden(name)
    st, u = run_d2d(emu, meta["args"])
    check_stats(meta, st, "dirk" in meta["args"])
    assert st["t"] == meta["t_final"]
    rel = float(np.linalg.norm(u - ref) / np.linalg.norm(ref))
    if "--fixedstep" in meta["args"]:
        assert np.array_equal(fmt16(u), fmt16(ref)), rel
    else:
        assert rel <= max(REL_L2_TOL, 3.0 * meta["ref_np1_vs_np4_rel_l2"]), rel
    assert st["kernel_launches"] > 0
    if any(m in meta["args"] for m in ("rkc", "rkl")):
        assert st["fused_launches"] >= 0.9 * (st["rhs_evals"] - 4)


# minutes on the emulator (thousands of steps); they stay in the GPU suite
SWEEP_SLOW = ("sweep_erk4_32_kx0.1_h0.00015625", "sweep_dirk3_32_kx1_h0.0025", "sweep_dirk2_64_kx1_rtol0.0001",
              "sweep_rkc_64_kx1_h0.00125")
SWEEP_SMALL = sorted(n for n in (os.path.basename(p)[4:-5] for p in __import__("glob").glob(os.path.join(GOLDEN, "d2d_sweep_*.json")))
                     if "_256_" not in n and "_128_" not in n and n not in SWEEP_SLOW)


@pytest.mark.parametrize("name", SWEEP_SMALL)
def test_reference_sweep_sample_through_the_emulated_stack(emu, name):
    """The 32^2 / 64^2 rows of the sample of the reference's evaluation matrix (runtests-diffusion2d.py)."""
    meta, ref = load_golden(name)
    st, u = run_d2d(emu, meta["args"])
    check_stats(meta, st, "dirk" in meta["args"])
    rel = float(np.linalg.norm(u - ref) / np.linalg.norm(ref))
    assert rel <= max(REL_L2_TOL, 3.0 * meta["ref_np1_vs_np4_rel_l2"]), rel


CHAIN_ARGS = [
    ["--nx", 128, "--ny", 48, "--kx", "1.0", "--ky", "0.5", "--inhomogeneous", "--integrator", "rkc",
     "--fixedstep", "0.0009765625", "--tf", "0.001953125"],
    ["--nx", 130, "--ny", 40, "--integrator", "rkl", "--fixedstep", "0.001953125", "--tf", "0.00390625"],
    ["--nx", 128, "--ny", 64, "--integrator", "rkc", "--tf", "0.05"],
]


@pytest.mark.parametrize("args", CHAIN_ARGS, ids=["rkc_fixed_inhom_128x48", "rkl_fixed_uniform_130x40", "rkc_adaptive_128x64"])
def test_temporal_blocking_and_deep_halo_paths_do_not_change_a_bit(emu, args):
    """--chain K (pending-stage chains -> k_chain_march), the k_chain_quad variant and the deep-halo flavour
    (--force-halo: halo buffers, extended tables, local-copy exchange) against one launch per stage."""
    st1, u1 = run_d2d(emu, args + ["--chain", "1"])
    assert st1["chain_launches"] == 0
    for extra in (["--chain", "2"], ["--chain", "4"], ["--chain", "6"], ["--chain", "4", "--force-halo"],
                  ["--chain", "3", "--chain-variant", "1"]):
        st, u = run_d2d(emu, args + extra)
        assert st["chain_launches"] > 0, extra
        for k in ("steps", "step_attempts", "err_test_fails", "rhs_evals", "max_stages"):
            assert st[k] == st1[k], (extra, k)
        assert np.array_equal(u, u1), extra
        if "--force-halo" not in extra:  # (the local-copy halo exchange adds pack / copy launches)
            assert st["kernel_launches"] < st1["kernel_launches"]
    emu.kernel_lib().b200_set_chain_variant(0)
    emu.sundials_lib().N_VSetStageChain_B200(4)


@pytest.mark.parametrize("extra", [[], ["--force-halo"]], ids=["wrap", "deep_halo"])
def test_stage_one_is_the_head_of_the_first_chain(emu, monkeypatch, extra):
    """Fixed-step STS: stage 1 (z_1 = y_n + c L(y_n), arkode_lsrkstep.c:640 / :930) and f_n = L(y_n) come out of the
    first chain launch of the step (HEAD flavour of k_chain_march) -- one launch less per step, every stage of the
    step inside a chain, and not a bit changes against the separate stage-1 launch (B200_NO_CHAIN_HEAD)."""
    for args in CHAIN_ARGS[:2]:
        monkeypatch.setenv("B200_NO_CHAIN_HEAD", "1")
        st0, u0 = run_d2d(emu, args + extra)
        monkeypatch.delenv("B200_NO_CHAIN_HEAD")
        st1, u1 = run_d2d(emu, args + extra)
        assert np.array_equal(u0, u1)
        assert st0["rhs_evals"] == st1["rhs_evals"] and st0["steps"] == st1["steps"]
        assert st0["chain_stages"] == st0["steps"] * (st0["max_stages"] - 1)  # stages 2..s
        assert st1["chain_stages"] == st1["steps"] * st1["max_stages"]        # stages 1..s: the whole step
        assert st1["chain_launches"] == st0["chain_launches"]
        if not extra:
            assert st1["kernel_launches"] == st0["kernel_launches"] - st0["steps"]


@pytest.mark.parametrize("extra", [[], ["--force-halo"]], ids=["wrap", "deep_halo"])
def test_adaptive_steps_start_their_first_chain_from_yn_alone(emu, monkeypatch, extra):
    """Adaptive STS: f_n is stored (the previous step's closing stage wrote it) and the vector knows it is L(y_n)
    (provenance), so stage 1 -- an elementwise y_n + c f_n in the reference (arkode_lsrkstep.c:640) -- joins the first
    chain, which recomputes f_n from y_n: no elementwise launch, same bits, same statistics."""
    args = CHAIN_ARGS[2] + extra
    monkeypatch.setenv("B200_NO_CHAIN_HEAD", "1")
    st0, u0 = run_d2d(emu, args)
    monkeypatch.delenv("B200_NO_CHAIN_HEAD")
    st1, u1 = run_d2d(emu, args)
    for k in ("steps", "step_attempts", "err_test_fails", "rhs_evals", "max_stages"):
        assert st0[k] == st1[k], k
    assert np.array_equal(u0, u1)
    assert st1["chain_stages"] >= st0["chain_stages"] + st0["steps"] - 2  # stage 1 of (nearly) every step is in a chain
    if not extra:
        assert st1["kernel_launches"] <= st0["kernel_launches"] - (st0["steps"] - 2)


@pytest.mark.parametrize("args", [
    ["--nx", 128, "--ny", 64, "--integrator", "rkc", "--tf", "0.05"],
    ["--nx", 32, "--ny", 32, "--integrator", "rkl", "--inhomogeneous", "--kx", "1", "--ky", "0.1", "--rtol", "1e-4", "--tf", "0.02"],
    ["--nx", 75, "--ny", 51, "--integrator", "rkc", "--tf", "0.03"],
], ids=["rkc_128x64", "rkl_inhom_32", "rkc_odd_75x51"])
def test_next_step_error_weights_come_out_of_the_closing_stage(emu, monkeypatch, args):
    """Adaptive STS: the closing-stage launch of a step also produces ewt(y_{n+1}) and ||y_{n+1}||_wrms
    (arkode.c:2985 / :835), so an accepted step costs no k_ewt_wsqr launch and one host synchronisation less.  The
    weights are the same bits and the norm only feeds the tolsf > 1 test, so the run is bit-identical to the one
    with the speculation off (B200_NO_SPEC_EWT)."""
    monkeypatch.setenv("B200_NO_SPEC_EWT", "1")
    st0, u0 = run_d2d(emu, args)
    monkeypatch.delenv("B200_NO_SPEC_EWT")
    st1, u1 = run_d2d(emu, args)
    for k in ("steps", "step_attempts", "err_test_fails", "rhs_evals", "max_stages", "dom_eig_updates"):
        assert st0[k] == st1[k], k
    assert np.array_equal(u0, u1)
    # the first step learns the pattern; every later accepted step saves the ewt launch
    assert st1["kernel_launches"] <= st0["kernel_launches"] - (st0["steps"] - 3)


def test_implicit_path_vector_work_is_fused(emu, monkeypatch):
    """DIRK3 + PCG + Jacobi (BASELINE configs[4] at 64^2): arkLsATimes o arkLsDQJtimes runs as ONE stencil launch that
    also returns <Ap, p>; r -= alpha*Ap with its weighted norm, z = P^-1 r with <r, z>, and p = z + beta*p with the WRMS
    norm of the next matvec are one kernel each -- about 5 launches per PCG iteration instead of ~18, same
    statistics as the reference fixture, state within its bar; with the fusion off the old launch count is back."""
    meta, ref = load_golden("dirk3_pcg_64")
    st, u = run_d2d(emu, meta["args"])
    check_stats(meta, st, True)
    assert st["lin_iters"] <= st["dq_fused"] <= st["lin_iters"] + st["nonlin_iters"] and st["ew_fused"] >= 2 * st["lin_iters"]
    assert st["kernel_launches"] <= 7.5 * st["lin_iters"], (st["kernel_launches"], st["lin_iters"])
    rel = float(np.linalg.norm(u - ref) / np.linalg.norm(ref))
    assert rel <= max(REL_L2_TOL, 3.0 * meta["ref_np1_vs_np4_rel_l2"]), rel
    st0, u0 = run_d2d(emu, meta["args"] + ["--no-fusion"])
    emu.sundials_lib().N_VSetLazyFusion_B200(1)
    check_stats(meta, st0, True)
    assert st0["dq_fused"] == 0 and st0["kernel_launches"] > 2 * st["kernel_launches"]
    assert float(np.linalg.norm(u - u0) / np.linalg.norm(u0)) <= max(REL_L2_TOL, 3.0 * meta["ref_np1_vs_np4_rel_l2"])
    # the power iteration's difference quotients (--internaleig, lsrkStep_DQJtimes) take the same kernel
    meta, ref = load_golden("rkl_internaleig_64")
    st, u = run_d2d(emu, meta["args"])
    check_stats(meta, st, False)
    assert st["dq_fused"] >= st["dee_rhs_evals"] > 0


def test_lazy_fusion_off_is_the_same_bits(emu):
    args = ["--nx", 96, "--ny", 64, "--kx", "1.0", "--ky", "0.5", "--inhomogeneous", "--integrator", "rkc",
            "--fixedstep", "0.0009765625", "--tf", "0.00390625"]
    st1, u1 = run_d2d(emu, args)
    st0, u0 = run_d2d(emu, args + ["--no-fusion"])
    emu.sundials_lib().N_VSetLazyFusion_B200(1)
    assert np.array_equal(u0, u1) and st0["rhs_evals"] == st1["rhs_evals"]
    assert st0["fused_launches"] == 0 and st1["fused_launches"] > 0


def test_pipelined_batches_equal_sequential_round_trips(emu):
    args = ["--nx", 128, "--ny", 32, "--integrator", "rkc", "--fixedstep", "0.000244140625", "--tf", "1.0",
            "--nout", "1", "--output", "0"]
    prob = emu.Diffusion2D([str(a) for a in args], device=0, stream=None)
    n, nb = 128 * 32, 3
    rng = np.random.default_rng(3)
    ins = [rng.random(n) for _ in range(nb)]
    want = []
    for i in range(nb):
        prob.set_state(ins[i], 0.125)
        prob.step(2)
        w = np.empty(n)
        prob.get_state(w)
        want.append(w)
    outs = [np.full(n, np.nan) for _ in range(nb)]
    prob.run_batches(ins, outs, 0.125, 2)
    prob.close()
    for i in range(nb):
        assert np.array_equal(outs[i], want[i]), i


# ------------------------------------------------------------------------------------------------ adr
ADR_CASES = ["strang_rkc_64", "strang_rkl_96x48_d", "strang_rkc_noadv_64", "extsts_ars_rkc_64",
             "extsts_ralston_rkl_fixed_48", "extsts_heun_rkc_fixed_64x48", "erk3_fixed_64x32"]


def fmt15(a):
    return np.array(["%.15g" % v for v in np.asarray(a).ravel()])


def test_block_solver_kernels_repeat_the_band_lu(emu):
    """b200_adr_jac_reaction / b200_blk2_scale_add_i / b200_blk2_factor / b200_blk2_solve against the reference's own
    sequence on the FULL banded matrix: J_reaction (...2d.cpp:1523-1551), SUNMatScaleAddI_Band, bandGBTRF, bandGBTRS
    (tests/band_lu_restatement.py) -- bit for bit, including blocks whose rows are swapped and a zero pivot."""
    from band_lu_restatement import band_gbtrf, band_gbtrs

    lib = emu.kernel_lib()
    ctx = ctypes.c_void_p()
    assert lib.b200_ctx_create(0, None, ctypes.byref(ctx)) == 0
    nx, ny = 6, 5
    npts, n = nx * ny, 2 * nx * ny
    rng = np.random.default_rng(11)
    y = rng.standard_normal(n) * 2.0
    B, gamma = 1.0, 0.37
    prm = emu.AdrParams(nx, ny, 1.0 / nx, 1.0 / ny, -0.5, 1.0, 0.4, 0.7, 1e-2, 1.3, B)
    J = np.full(4 * npts, np.nan)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.b200_adr_jac_reaction(ctx, ctypes.byref(prm), P(y), P(J)) == 0
    dense = np.zeros((n, n))
    for p_ in range(npts):
        u, v = y[2 * p_], y[2 * p_ + 1]
        dense[2 * p_, 2 * p_] = 2.0 * u * v - (B + 1.0)
        dense[2 * p_ + 1, 2 * p_] = B - 2.0 * u * v
        dense[2 * p_, 2 * p_ + 1] = u * u
        dense[2 * p_ + 1, 2 * p_ + 1] = -u * u
    want = np.stack([dense[0::2, 0::2].diagonal(), dense[1::2, 0::2].diagonal(), dense[0::2, 1::2].diagonal(),
                     dense[1::2, 1::2].diagonal()], axis=1).ravel()
    assert np.array_equal(J, want)
    # A = I - gamma*J (SUNMatScaleAddI(-gamma, A)): every stored band entry times c, then the diagonal + 1
    assert lib.b200_blk2_scale_add_i(ctx, ctypes.c_double(-gamma), P(J), ctypes.c_int64(npts)) == 0
    A = dense * (-gamma)
    A[np.arange(n), np.arange(n)] += 1.0
    got = np.stack([A[0::2, 0::2].diagonal(), A[1::2, 0::2].diagonal(), A[0::2, 1::2].diagonal(), A[1::2, 1::2].diagonal()],
                   axis=1).ravel()
    assert np.array_equal(J, got)
    piv = np.full(npts, np.nan)
    info = ctypes.c_longlong(-1)
    assert lib.b200_blk2_factor(ctx, P(J), P(piv), ctypes.c_int64(npts), ctypes.byref(info)) == 0
    pv, winfo = band_gbtrf(A, 2, 4)
    assert info.value == winfo == 0
    swapped = pv[0::2] != np.arange(0, n, 2)
    assert swapped.any() and not swapped.all()  # both branches of the pivot search are exercised
    assert np.array_equal(piv != 0.0, swapped) and np.all(pv[1::2] == np.arange(1, n, 2))
    got = np.stack([A[0::2, 0::2].diagonal(), A[1::2, 0::2].diagonal(), A[0::2, 1::2].diagonal(), A[1::2, 1::2].diagonal()],
                   axis=1).ravel()
    assert np.array_equal(J, got)
    off = A.copy()
    for p_ in range(npts):
        off[2 * p_:2 * p_ + 2, 2 * p_:2 * p_ + 2] = 0.0
    assert not off.any()  # the band elimination never left the blocks
    b = rng.standard_normal(n)
    x = np.full(n, np.nan)
    assert lib.b200_blk2_solve(ctx, P(J), P(piv), P(b), P(x), ctypes.c_int64(npts)) == 0
    assert np.array_equal(x, band_gbtrs(A, 2, 4, pv, b.copy()))
    # a zero pivot: bandGBTRF's 1-based column number comes back
    Z = J.copy()
    Z[4 * 7:4 * 7 + 2] = 0.0
    Zd = np.zeros((n, n))
    for p_ in range(npts):
        Zd[2 * p_, 2 * p_], Zd[2 * p_ + 1, 2 * p_], Zd[2 * p_, 2 * p_ + 1], Zd[2 * p_ + 1, 2 * p_ + 1] = Z[4 * p_:4 * p_ + 4]
    assert lib.b200_blk2_factor(ctx, P(Z), P(piv), ctypes.c_int64(npts), ctypes.byref(info)) == 0
    assert info.value == band_gbtrf(Zd, 2, 4)[1] == 15
    lib.b200_ctx_destroy(ctx)


ADR_IMPLICIT_CASES = ["strang_rkc_implreact_48", "strang_rkl_implreact_noadv_40x32", "extsts_ars_implreact_fixed_48",
                      "extsts_giraldo_implreact_noadv_48", "extsts_sdirk_implreact_fixed_64x32",
                      "extsts_giraldo_implreact_fixed_64"]


@pytest.mark.parametrize("name", ADR_IMPLICIT_CASES)
def test_adr_implicit_reaction_fixture_parity_through_the_emulated_stack(emu, name):
    """--implicit-reaction (SetupStrang / SetupExtSTS with fi = f_reaction): ARKODE's Newton iteration over the
    block-diagonal device matrix / direct solver of b200_blockdiag.h, which repeat the reference's band LU
    (SUNBandMatrix(neq, 2, 2) + SUNLinSol_Band) block by block -- same Newton / RHS / step counters as the unmodified
    reference, fixed-step states equal to every printed digit."""
    with open(os.path.join(GOLDEN, "adr_%s.json" % name)) as f:
        meta = json.load(f)
    ref = np.load(os.path.join(GOLDEN, "adr_%s.npy" % name))
    args = [str(a) for a in meta["args"]]
    tf = float(args[args.index("--tf") + 1])
    prob = emu.Adr2D(args + ["--nout", "1", "--output", "0"], device=0, stream=None)
    prob.evolve(tf)
    st = prob.stats()
    y = np.empty(st["neq"])
    prob.get_state(y)
    prob.close()
    ms = meta["stats"]
    assert st["steps"] == ms["outer.steps"] and st["lsrk_rhs_evals"] == ms["lsrk.rhs_evals"]
    if "ark.rhs_evals_i" in ms:  # Strang: the ARKStep partition owns the reaction
        assert st["ark_rhs_evals_implicit"] == ms["ark.rhs_evals_i"] and st["ark_rhs_evals"] == ms["ark.rhs_evals_e"]
        assert st["nls_iters"] == ms["ark.nls_iters"]
    else:
        assert st["rhs_evals_implicit"] == ms["outer.rhs_evals_i"] and st["rhs_evals_explicit"] == ms["outer.rhs_evals_e"]
        assert st["nls_iters"] == ms["outer.nls_iters"]
    assert st["jac_evals"] >= 1 and st["ls_setups"] >= 1
    if "--fixed_h" in args:
        assert np.array_equal(fmt15(y), fmt15(ref))
    else:
        assert float(np.linalg.norm(y - ref) / np.linalg.norm(ref)) <= REL_L2_TOL


@pytest.mark.parametrize("name", ADR_CASES)
def test_adr_fixture_parity_through_the_emulated_stack(emu, name):
    with open(os.path.join(GOLDEN, "adr_%s.json" % name)) as f:
        meta = json.load(f)
    ref = np.load(os.path.join(GOLDEN, "adr_%s.npy" % name))
    args = [str(a) for a in meta["args"]]
    tf = float(args[args.index("--tf") + 1]) if "--tf" in args else 1.0
    prob = emu.Adr2D(args + ["--nout", "1", "--output", "0"], device=0, stream=None)
    prob.evolve(tf)
    st = prob.stats()
    y = np.empty(st["neq"])
    prob.get_state(y)
    prob.close()
    assert st["steps"] == meta["stats"]["outer.steps"]
    if "lsrk.rhs_evals" in meta["stats"]:
        assert st["lsrk_rhs_evals"] == meta["stats"]["lsrk.rhs_evals"]
        assert st["lsrk_max_stages"] == meta["stats"]["lsrk.max_stages"]
    assert st["fused_launches"] > 0
    assert np.array_equal(fmt15(y), fmt15(ref))


FAILURE_SCRIPT = r"""
import ctypes, importlib, os, sys
sys.path.insert(0, {root!r})
b200 = importlib.import_module("ceda-demonstrations_b200")
lib = ctypes.CDLL({emu!r})
lib.b200_last_error.restype = ctypes.c_char_p
lib.b200_launch_count.restype = ctypes.c_uint64
b200._kernel_lib = b200._sundials_lib = lib
prob = b200.Diffusion2D({args!r}, device=0, stream=None)
try:
    prob.evolve(0.05)
    print("EVOLVE-RETURNED-OK")
except RuntimeError as exc:
    print("EVOLVE-ERROR", exc)
print("FAILED-FLAG", lib.N_VDeviceFailed_B200())
prob.close()
print("CLOSED")
"""


@pytest.mark.parametrize("args", [["--nx", "128", "--ny", "32", "--integrator", "rkc", "--tf", "0.05"],
                                  ["--nx", "64", "--ny", "32", "--integrator", "dirk", "--order", "3", "--tf", "0.05"]],
                         ids=["rkc", "dirk_pcg"])
def test_a_device_failure_comes_back_as_an_error_code_not_an_abort(emu, args):
    """A failing launch in the middle of an integration (fault injection: B200_FAIL_AFTER_LAUNCHES) is printed once,
    turns into SUN_ERR_EXT_FAIL / NaN at the ops table, ARKodeEvolve returns an error, the session can be closed, the
    process lives -- and nothing is computed on the host instead."""
    import sys

    code = FAILURE_SCRIPT.format(root=ROOT, emu=EMU_LIB, args=args + ["--nout", "1", "--output", "0"])
    env = dict(os.environ, B200_FAIL_AFTER_LAUNCHES="60")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "EVOLVE-ERROR" in r.stdout and "FAILED-FLAG 1" in r.stdout and "CLOSED" in r.stdout, r.stdout
    assert "injected failure" in r.stderr and r.stderr.count("nvector_b200: FATAL") == 1, r.stderr[-2000:]


def set_ring_flavour(lib, level):
    """How k_chain_march (depth 4) fills its operand ring: -1 = plain cp.async, 3 rows ahead; 1 / 2 = cp.async.bulk +
    mbarrier, 3 / 4 rows ahead; 3 = plain cp.async, 4 rows ahead; 0 = back to the library's defaults."""
    lib.b200_set_chain_bulk({-1: 0, 0: -1, 1: 1, 2: 2, 3: 0}[level])
    lib.b200_set_chain_pf({-1: 3, 0: 0, 1: 3, 2: 3, 3: 4}[level])


@pytest.mark.parametrize("args", CHAIN_ARGS, ids=["rkc_fixed_inhom_128x48", "rkl_fixed_uniform_130x40", "rkc_adaptive_128x64"])
@pytest.mark.parametrize("extra", [[], ["--force-halo"]], ids=["wrap", "deep_halo"])
@pytest.mark.parametrize("level", [1, 2, 3], ids=["bulk_pf3", "bulk_pf4", "plain_pf4"])
def test_bulk_copy_ring_does_not_change_a_bit(emu, args, extra, level):
    """BULK flavour of k_chain_march (depth 4): the operand ring filled by cp.async.bulk + mbarrier (interior windows)
    and by per-thread cp.async (the windows on the block's first and last columns); whole integrations through the
    launchers -- tables and uniform coefficients, wrap and deep halos, head and body chains -- agree bit for bit."""
    lib = emu.kernel_lib()
    set_ring_flavour(lib, -1)
    st0, u0 = run_d2d(emu, args + ["--chain", "4"] + extra)
    if level >= 2 and args is not CHAIN_ARGS[0]:
        pytest.skip("the deeper prefetch is covered by the first case (suite run time)")
    set_ring_flavour(lib, level)
    try:
        st1, u1 = run_d2d(emu, args + ["--chain", "4"] + extra)
    finally:
        set_ring_flavour(lib, 0)
    assert st0["chain_launches"] == st1["chain_launches"] > 0 and st0["chain_stages"] == st1["chain_stages"]
    for k in ("steps", "step_attempts", "err_test_fails", "rhs_evals", "max_stages"):
        assert st0[k] == st1[k], k
    assert np.array_equal(u0, u1)


@pytest.mark.parametrize("args", CHAIN_ARGS, ids=["rkc_fixed_inhom_128x48", "rkl_fixed_uniform_130x40", "rkc_adaptive_128x64"])
@pytest.mark.parametrize("extra", [[], ["--force-halo"]], ids=["wrap", "deep_halo"])
def test_split_level_groups_do_not_change_a_bit(emu, args, extra):
    """SPLIT flavour of k_chain_march (depth 4): the upper half of the levels runs one row late and first in a row step
    (two independent instruction streams); every cell still sees the same instruction sequence, so whole integrations
    -- table-driven and uniform coefficients, index wrap and deep halos, chains that begin a step (HEAD) and chains that
    do not, full and partial row blocks -- agree bit for bit with the plain order."""
    lib = emu.kernel_lib()
    lib.b200_set_chain_split(0)
    st0, u0 = run_d2d(emu, args + ["--chain", "4"] + extra)
    lib.b200_set_chain_split(1)
    try:
        st1, u1 = run_d2d(emu, args + ["--chain", "4"] + extra)
    finally:
        lib.b200_set_chain_split(0)
    assert st0["chain_launches"] == st1["chain_launches"] > 0 and st0["chain_stages"] == st1["chain_stages"]
    for k in ("steps", "step_attempts", "err_test_fails", "rhs_evals", "max_stages"):
        assert st0[k] == st1[k], k
    assert np.array_equal(u0, u1)
