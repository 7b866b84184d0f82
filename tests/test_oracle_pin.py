"""Pin the CPU oracle (oracle/sts_oracle.c) against the reference's own known answers.

(a) SUNDIALS' golden stage logs for RKC2 / RKL2 / SSP(s,2) / SSP(s,3)
    (tests/golden/lsrk_logging_golden.json, extracted from
    deps/sundials/test/unit_tests/logging/test_logging_arkode_lsrkstep_lvl5_{0..5}.out);
(b) final states written by the unmodified reference driver (tests/golden/d2d_*.npy);
(c) when the reference binary is present (build container), a live run of it;
(e) SUNDIALS' known answers for the power iteration
    (deps/sundials/test/answers/.../test_sundomeigest_power_{1000,10000,100000}_100_0_0.out:
    eigenvalue estimate, iteration count and residual).
"""
import ctypes
import json
import math
import os

import numpy as np
import pytest
from conftest import GOLDEN, ROOT, OrcStepWs, P, RHS_FN, fmt16, load_golden, make_grid

# ---- the scalar Prothero-Robinson problem of the golden logs (problems/prv.hpp) ----------------
LAM, ALPHA = -1000.0, 10.0


def l_coef(t):
    return LAM - ALPHA * math.cos((10.0 - t) / 10.0 * math.acos(-1.0))


def prv_rhs(t, y):
    return l_coef(t) * (y - math.atan(t)) + 1.0 / (1.0 + t * t)


def run_scalar_step(orc, method, st, rtol, atol):
    calls = []

    def rhs(t, y, f, user):
        f[0] = prv_rhs(t, y[0])
        calls.append(f[0])
        return 0

    cb = RHS_FN(rhs)
    vec = {k: np.zeros(1) for k in ("yn", "fn", "ycur", "tempv1", "tempv2", "tempv3", "ewt")}
    vec["yn"][0] = st["z0"]
    vec["fn"][0] = prv_rhs(st["tn"], st["z0"])
    tmp = np.zeros(1)
    orc.orc_ewt_ss(P(vec["yn"]), ctypes.c_double(rtol), ctypes.c_double(atol), P(tmp), P(vec["ewt"]), 1)
    ws = OrcStepWs(1, 1, *[vec[k].ctypes.data for k in ("yn", "fn", "ycur", "tempv1", "tempv2", "tempv3", "ewt")], 0, 0)
    dsm = ctypes.c_double()
    tn, h = ctypes.c_double(st["tn"]), ctypes.c_double(st["h"])
    if method in ("rkc", "rkl"):
        # spectral radius = |1.01 * lambda(tn)| (arkode_lsrkstep.c:2340-2343); the log prints it rounded
        sr = abs(1.01 * l_coef(st["tn"]))
        assert sr == pytest.approx(st["spectral_radius"], rel=1e-12)
        fn = orc.orc_step_rkc if method == "rkc" else orc.orc_step_rkl
        s = fn(ctypes.byref(ws), cb, None, tn, h, ctypes.c_double(sr), ctypes.byref(dsm))
    elif method in ("ssp43", "ssp104"):
        fn = orc.orc_step_ssp43 if method == "ssp43" else orc.orc_step_ssp104
        s = fn(ctypes.byref(ws), cb, None, tn, h, ctypes.byref(dsm))
    else:
        nst = len(st["F"])  # F_0 .. F_{s-1}
        fn = orc.orc_step_ssps2 if method == "ssps2" else orc.orc_step_ssps3
        s = fn(ctypes.byref(ws), cb, None, tn, h, nst, ctypes.byref(dsm))
    # which work vector holds the result differs per method; ycur is the solution in all of them
    return s, vec["ycur"][0], calls, dsm.value


@pytest.mark.parametrize("method", ["rkc", "rkl", "ssps2", "ssps3", "ssp43", "ssp104"])
def test_lsrk_golden_stage_logs(orc, method):
    with open(os.path.join(GOLDEN, "lsrk_logging_golden.json")) as f:
        gold = json.load(f)
    steps = gold["methods"][method]
    assert len(steps) == 3
    for st in steps:
        s, ycur, calls, dsm = run_scalar_step(orc, method, st, gold["rtol"], gold["atol"])
        if st["stages"] is not None:
            assert s == st["stages"]
        # h and tn are printed with 15 digits in the log -> compare at 1e-12
        assert ycur == pytest.approx(st["ycur"], rel=1e-12, abs=1e-300) if "ycur" in st else True
        if method in ("rkc", "rkl"):
            # F_1 .. F_{s-1} are the stage RHS values, then F_n closes the step
            want = st["F"][1:] + [st["Fn"]]
            assert len(calls) == len(want)
            np.testing.assert_allclose(calls, want, rtol=1e-12)
        else:
            want = st["F"][1:]
            np.testing.assert_allclose(calls[: len(want)], want, rtol=1e-12)
        # the error estimate is a difference of nearly equal numbers: ~1e-4 relative is what the
        # 15-digit inputs of the log allow on this scalar problem
        if st.get("dsm", 0.0) > 0.0:
            assert dsm == pytest.approx(st["dsm"], rel=5e-3)


# ---- (b) reference driver outputs -----------------------------------------------------------------
FIXED = {
    "rkc_fixed_aniso_inhom_96x64": (0, 0.0009765625, 4),
    "rkl_fixed_aniso_inhom_96x64": (1, 0.0009765625, 8),
}


@pytest.mark.parametrize("name", sorted(FIXED))
def test_oracle_fixed_step_run_matches_reference_output(orc, name):
    """Fixed step + analytic dom_eig: no reduction feeds back into the state, so the oracle must
    reproduce the reference's final state digit for digit (the reference prints 16 digits)."""
    meta, ref = load_golden(name)
    a = meta["args"]
    nx, ny = int(a[a.index("--nx") + 1]), int(a[a.index("--ny") + 1])
    g = make_grid(nx, ny, kx=float(a[a.index("--kx") + 1]), ky=float(a[a.index("--ky") + 1]), inhom="--inhomogeneous" in a)
    method, h, nsteps = FIXED[name]
    u = np.zeros(nx * ny)
    nfe = orc.orc_diffusion_fixed_run(ctypes.byref(g), method, ctypes.c_double(h), nsteps, P(u))
    assert nfe == meta["stats"]["rhs_evals"]
    assert meta["stats"]["steps"] == nsteps
    assert np.array_equal(fmt16(u), fmt16(ref))


def test_oracle_initial_condition_and_urms(orc):
    """t = 0 line of the reference's C1 run: ||u||_rms = 1.212147970328282e-01."""
    g = make_grid(128, 128)
    u = np.zeros(128 * 128)
    orc.orc_initial(ctypes.byref(g), P(u))
    urms = math.sqrt(orc.orc_dot(P(u), P(u), ctypes.c_int64(u.size)) / 128 / 128)
    assert "%.15e" % urms == "1.212147970328282e-01"


def test_oracle_dom_eig_matches_reference_stats(orc):
    meta, _ = load_golden("c1_rkc_128")
    g = make_grid(128, 128)
    lam = orc.orc_dom_eig(ctypes.byref(g)) * 1.01
    assert abs(lam) == pytest.approx(meta["stats"]["sr_max"], rel=1e-14)
    assert "%.15g" % abs(lam) == "3301.10292935388"


# ---- (c) live reference binary (build container only) -----------------------------------------------
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "diffusion_2D_ref")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("method,flag", [(0, "rkc"), (1, "rkl")])
def test_oracle_against_live_reference_binary(orc, method, flag):
    import compare_runs as cr

    nx, ny, h, nsteps = 40, 36, 2.0 ** -9, 3
    args = ["--nx", str(nx), "--ny", str(ny), "--kx", "0.7", "--ky", "1.3", "--inhomogeneous", "--integrator", flag,
            "--fixedstep", repr(h), "--tf", repr(nsteps * h), "--nout", "1", "--output", "2"]
    wd, text = cr.run(cr.REF_BIN, args, 1)
    _, ref = cr.read_solution(wd, nx, ny)
    g = make_grid(nx, ny, kx=0.7, ky=1.3, inhom=True)
    u = np.zeros(nx * ny)
    nfe = orc.orc_diffusion_fixed_run(ctypes.byref(g), method, ctypes.c_double(h), nsteps, P(u))
    assert nfe == cr.parse_stats(text)["rhs_evals"]
    assert np.array_equal(fmt16(u), fmt16(ref))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_multirank_shim_matches_single_rank():
    """The in-tree MPI shim runs the reference's real halo-exchange path on several ranks; a fixed-step
    run must not depend on the decomposition at all."""
    import compare_runs as cr

    args = ["--nx", "50", "--ny", "34", "--inhomogeneous", "--integrator", "rkc", "--fixedstep", "0.001953125",
            "--tf", "0.0078125", "--nout", "1", "--output", "2"]
    _, u1 = cr.read_solution(cr.run(cr.REF_BIN, args, 1)[0], 50, 34)
    for np_ranks in (2, 3, 4, 6):
        _, up = cr.read_solution(cr.run(cr.REF_BIN, args, np_ranks)[0], 50, 34)
        assert np.array_equal(u1, up), np_ranks


# ---- (e) SUNDIALS' known answers for the power iteration ---------------------------------------
# test/unit_tests/sundomeigest/Power/test_sundomeigest_power.c: A = diag(-100*[3..N]) plus a 2x2 block
# [[-30000, -10000], [-10000, -30000]] on the last two rows, initial guess q_i = rand()/RAND_MAX (glibc, default
# seed), rel_tol 0.01, no warm-ups; answers in test/answers/linux-ubuntu20.04-x86_64/gcc-9.4.0/double/
# test_sundomeigest_power_<N>_100_0_0.out (printed with 15 significant digits).
POWER_GOLDEN = {1000: (-93929.2359849011, 8, 0.00959956382300218),
                10000: (-936780.175307443, 8, 0.00966387730674346),
                100000: (-9374195.86189535, 8, 0.00950100649833683)}


@pytest.mark.parametrize("n", sorted(POWER_GOLDEN))
def test_oracle_power_iteration_matches_sundials_known_answers(orc, n):
    from conftest import ATIMES_FN

    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)  # the C default seed
    rand_max = 2147483647
    q = np.array([libc.rand() / rand_max for _ in range(n)])
    diag = -100.0 * (np.arange(n - 2) + 3.0)
    lambdas = []

    def atimes(user, v, z):
        va = np.ctypeslib.as_array(v, shape=(n,))
        za = np.ctypeslib.as_array(z, shape=(n,))
        za[:n - 2] = diag * va[:n - 2]
        za[n - 2] = va[n - 2] * -30000.0 + va[n - 1] * -10000.0
        za[n - 1] = va[n - 1] * -30000.0 + va[n - 2] * -10000.0
        lambdas.append(float(np.dot(va, za)))
        return 0

    V = q / math.sqrt(orc.orc_dot(P(q), P(q), ctypes.c_int64(n)))  # SUNDomEigEstimator_Initialize_Power :181-186
    V = np.ascontiguousarray(V)
    work = np.zeros(n)
    lam, iters = ctypes.c_double(), ctypes.c_int()
    rc = orc.orc_power_iteration(ATIMES_FN(atimes), None, P(V), P(work), ctypes.c_int64(n), 0, 100,
                                 ctypes.c_double(0.01), ctypes.byref(lam), ctypes.byref(iters))
    want_lam, want_iters, want_res = POWER_GOLDEN[n]
    assert rc == 0 and iters.value == want_iters
    assert lam.value == pytest.approx(want_lam, rel=2e-14)
    res = abs(lambdas[-1] - lambdas[-2]) / abs(lambdas[-1])
    assert res == pytest.approx(want_res, rel=1e-10)
