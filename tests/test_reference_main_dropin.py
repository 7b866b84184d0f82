"""The drop-in boundary, end to end: the reference's OWN diffusion_2D/main.cpp -- option parsing, UserData::setup,
Initial(), the whole ARKODE call sequence, UserOutput, ARKodePrintAllStats -- with the 12-line patch of INTEGRATION.md
section 1 (tests/native/patch_reference_main.py: vector constructor, callbacks, user_data), compiled together with the
reference's other unmodified sources by `make -C oracle refmain`, must reproduce the fixtures the unmodified
reference build wrote (tests/golden): equal statistics, fixed-step states equal to every printed digit, adaptive
states within 1e-10 (or the reference's own np=1 / np=4 spread).

  * CPU suite: the binary linked against the full-stack emulation library (tests/emu)
  * GPU suite: the binary linked against the product libraries, on the B200
"""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest
from conftest import ROOT, fmt16, load_golden

import compare_runs as cr

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
BIN_GPU = os.path.join(REF_DIR, "diffusion_2D_refmain_b200")
BIN_EMU = os.path.join(REF_DIR, "diffusion_2D_refmain_b200_emu")
CASES = ["c1_rkc_128", "rkl_fixed_aniso_inhom_96x64", "rkl_internaleig_64", "ssp104_fixed_64", "erk3_adaptive_64", "dirk3_pcg_64",
         "rkc_odd_75x51"]
STATS = ("steps", "attempts", "err_fails", "rhs_evals", "rhs_evals_e", "rhs_evals_i", "max_stages", "dom_eig_updates", "dee_evals",
         "lin_iters", "nls_iters")


def _build():
    if os.path.isdir("/root/reference/diffusion_2D"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "refmain"], check=True)


def _run_case(binary, name):
    meta, ref = load_golden(name)
    args = [str(a) for a in meta["args"]] + ["--nout", "1", "--output", "2"]
    wd = tempfile.mkdtemp(prefix="refmain_")
    try:
        res = subprocess.run([binary] + args, cwd=wd, capture_output=True, text=True, timeout=900, env=dict(os.environ, MPISHIM_NP="1"))
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        stats = cr.parse_stats(res.stdout)
        _, u = cr.read_solution(wd, ref.shape[1], ref.shape[0])
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    np4 = meta.get("stats_np4", {})
    for k in STATS:
        if k in meta["stats"] and k in stats:
            v, v4 = meta["stats"][k], np4.get(k, meta["stats"][k])
            d = abs(v4 - v)
            assert min(v, v4) - d <= stats[k] <= max(v, v4) + d, (k, stats[k], v, v4)
    rel = float(np.linalg.norm(u - ref) / np.linalg.norm(ref))
    if "--fixedstep" in meta["args"]:
        assert np.array_equal(fmt16(u), fmt16(ref)), rel
    else:
        assert rel <= max(1e-10, 3.0 * meta["ref_np1_vs_np4_rel_l2"]), rel
    # the reference's own report lines are there: it IS the reference's main()
    assert "Final integrator statistics:" in res.stdout and "Total simulation time" in res.stdout


@pytest.mark.parametrize("name", CASES)
def test_reference_main_on_the_emulated_stack(name):
    _build()
    if not os.path.exists(BIN_EMU):
        pytest.skip("oracle/_ref/diffusion_2D_refmain_b200_emu not built (needs /root/reference)")
    _run_case(BIN_EMU, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_reference_main_on_the_b200(name):
    if not os.path.exists(BIN_GPU):
        pytest.skip("oracle/_ref/diffusion_2D_refmain_b200 not shipped (make -C oracle refmain)")
    _run_case(BIN_GPU, name)


def test_the_patch_is_the_documented_handful_of_lines():
    """12 changed lines, one of them the added #include -- and every pattern must still be where INTEGRATION.md says."""
    src = "/root/reference/diffusion_2D/main.cpp"
    if not os.path.exists(src):
        pytest.skip("needs /root/reference")
    out = os.path.join(tempfile.mkdtemp(prefix="patch_"), "main_patched.cpp")
    subprocess.run(["python", os.path.join(ROOT, "tests", "native", "patch_reference_main.py"), src, out], check=True)
    a, b = open(src).read().splitlines(), open(out).read().splitlines()
    assert len(b) == len(a) + 1
    b.remove('#include "refmain_adapter.hpp"')
    assert sum(1 for x, y in zip(a, b) if x != y) == 11
