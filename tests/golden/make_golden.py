#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run HERE, where /root/reference exists).

Two sources, both the reference's own:

1. SUNDIALS' committed known-answer logs for the LSRKStep stage recurrences
   /root/reference/deps/sundials/test/unit_tests/logging/test_logging_arkode_lsrkstep_lvl5_{0..5}.out
   (RKC2, RKL2, SSP(s,2), SSP(s,3) on the scalar Prothero-Robinson problem prv.hpp).  Only the
   numbers are extracted -> lsrk_logging_golden.json.

2. Outputs of the UNMODIFIED reference driver compiled into oracle/_ref/diffusion_2D_ref
   (make -C oracle ref): integrator statistics and the final state (the reference's own
   `--output 2` text, 16 significant digits) for a handful of small configurations
   -> d2d_<name>.json / d2d_<name>.npy.   The np=1 vs np=4 spread of the reference itself is
   recorded too: it is the reference's own sensitivity to the reduction order.

    python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import compare_runs as cr  # noqa: E402

LOGDIR = "/root/reference/deps/sundials/test/unit_tests/logging"

# name -> reference command line (all --nout 1 --output 2)
D2D_CASES = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "c1_rkc_128": ["--nx", "128", "--ny", "128", "--integrator", "rkc", "--tf", "1"],
    "rkc_fixed_aniso_inhom_96x64": ["--nx", "96", "--ny", "64", "--kx", "1.0", "--ky", "0.5", "--inhomogeneous",
                                    "--integrator", "rkc", "--fixedstep", "0.0009765625", "--tf", "0.00390625"],
    "rkl_fixed_aniso_inhom_96x64": ["--nx", "96", "--ny", "64", "--kx", "1.0", "--ky", "0.1", "--inhomogeneous",
                                    "--integrator", "rkl", "--fixedstep", "0.0009765625", "--tf", "0.0078125"],
    "rkl_adaptive_inhom_64": ["--nx", "64", "--ny", "64", "--kx", "1.0", "--ky", "0.1", "--inhomogeneous",
                              "--integrator", "rkl", "--tf", "0.1"],
    "rkl_internaleig_64": ["--nx", "64", "--ny", "64", "--kx", "1.0", "--ky", "0.1", "--inhomogeneous",
                           "--integrator", "rkl", "--internaleig", "--tf", "0.1"],
    "rkc_odd_75x51": ["--nx", "75", "--ny", "51", "--integrator", "rkc", "--tf", "0.1"],
    "ssp2_fixed_64": ["--nx", "64", "--ny", "64", "--integrator", "erk", "--order", "-2",
                      "--fixedstep", "0.0001220703125", "--tf", "0.0009765625"],
    "ssp3_fixed_64": ["--nx", "64", "--ny", "64", "--integrator", "erk", "--order", "-3",
                      "--fixedstep", "0.0001220703125", "--tf", "0.0009765625"],
    "ssp104_fixed_64": ["--nx", "64", "--ny", "64", "--integrator", "erk", "--order", "-4",
                        "--fixedstep", "0.0001220703125", "--tf", "0.0009765625"],
    "ssp104_adaptive_64": ["--nx", "64", "--ny", "64", "--integrator", "erk", "--order", "-4", "--tf", "0.05"],
    "erk3_adaptive_64": ["--nx", "64", "--ny", "64", "--integrator", "erk", "--order", "3", "--tf", "0.02"],
    "dirk3_pcg_64": ["--nx", "64", "--ny", "64", "--integrator", "dirk", "--order", "3", "--tf", "0.1"],
    "dirk2_pcg_inhom_noprec_48": ["--nx", "48", "--ny", "48", "--inhomogeneous", "--integrator", "dirk",
                                  "--order", "2", "--noprec", "--tf", "0.05"],
}


# A sample of the reference's own evaluation matrix, diffusion_2D/runtests-diffusion2d.py:72-104
# (solvers x grids x kx x rtol, and the fixed-step series), with exactly its common flags :76-83.
SWEEP_COMMON = ["--inhomogeneous", "--atol", "1.e-11", "--controller", "2", "--error", "--nonlinear", "--msbp", "1",
                "--maxsteps", "100000", "--internaleig"]
SWEEP_SOLVERS = {"rkc": ["--integrator", "rkc"], "rkl": ["--integrator", "rkl"],
                 "erk2": ["--integrator", "erk", "--order", "-2"], "erk3": ["--integrator", "erk", "--order", "-3"],
                 "erk4": ["--integrator", "erk", "--order", "-4"],
                 "dirk2": ["--integrator", "dirk", "--order", "2"], "dirk3": ["--integrator", "dirk", "--order", "3"]}
SWEEP_SAMPLE = [  # (solver, grid, kx, rtol, fixed h or 0)
    ("rkc", 32, 0.1, 1e-2, 0.0), ("rkc", 64, 1.0, 1e-4, 0.0), ("rkc", 128, 10.0, 1e-6, 0.0), ("rkc", 256, 1.0, 1e-3, 0.0),
    ("rkl", 32, 10.0, 1e-3, 0.0), ("rkl", 64, 0.1, 1e-5, 0.0), ("rkl", 128, 1.0, 1e-2, 0.0), ("rkl", 256, 10.0, 1e-4, 0.0),
    ("erk2", 64, 1.0, 1e-3, 0.0), ("erk3", 32, 10.0, 1e-4, 0.0), ("erk4", 64, 0.1, 1e-5, 0.0),
    ("dirk2", 64, 1.0, 1e-4, 0.0), ("dirk3", 128, 0.1, 1e-3, 0.0),
    ("rkc", 64, 1.0, 1e-9, 1e-2 / 8), ("erk4", 32, 0.1, 1e-9, 1e-2 / 64), ("dirk3", 32, 1.0, 1e-9, 1e-2 / 4),
    # not sampled: the matrix's fixed-step RKL rows at 128^2 (kx = 10, h = 1e-2/32) and 256^2 (kx = 1,
    # h = 1e-2/2) -- the reference itself returns NaN there (h rho exceeds what 200 stages cover)
]
for _solver, _grid, _kx, _rtol, _h in SWEEP_SAMPLE:
    _name = "sweep_%s_%d_kx%g_%s" % (_solver, _grid, _kx, ("h%g" % _h) if _h > 0 else ("rtol%g" % _rtol))
    _args = ["--nx", str(_grid), "--ny", str(_grid), "--rtol", "%e" % _rtol, "--kx", "%e" % _kx, "--ky", "%e" % 0.0]
    _args += SWEEP_SOLVERS[_solver] + SWEEP_COMMON
    if _h > 0:
        _args += ["--fixedstep", "%e" % _h]
    D2D_CASES[_name] = _args


# adr 2-D (oracle/_ref/adr2d_ref): name -> reference command line (all --nout 1 --output 1)
ADR_CASES = {
    "strang_rkc_64": ["--nx", "64", "--ny", "64", "--integrator", "3", "--sts_method", "0", "--fixed_h", "0.01", "--tf", "0.1"],
    "strang_rkl_96x48_d": ["--nx", "96", "--ny", "48", "--integrator", "3", "--sts_method", "1", "--d", "0.05",
                           "--fixed_h", "0.005", "--tf", "0.05"],
    "strang_rkc_noadv_64": ["--nx", "64", "--ny", "64", "--integrator", "3", "--sts_method", "0", "--no-advection",
                            "--fixed_h", "0.01", "--tf", "0.05"],
    "extsts_ars_rkc_64": ["--nx", "64", "--ny", "64", "--integrator", "2", "--sts_method", "0", "--extsts_method", "0",
                          "--fixed_h", "0.005", "--tf", "0.05"],
    "extsts_ralston_rkl_fixed_48": ["--nx", "48", "--ny", "48", "--integrator", "2", "--sts_method", "1",
                                    "--extsts_method", "2", "--fixed_h", "0.004", "--tf", "0.04"],
    "erk2_adaptive_48": ["--nx", "48", "--ny", "48", "--integrator", "0", "--order", "2", "--rtol", "1e-4", "--tf", "0.02"],
    "erk3_fixed_64x32": ["--nx", "64", "--ny", "32", "--integrator", "0", "--order", "3", "--fixed_h", "1e-4", "--tf", "2e-3"],
    "extsts_heun_rkc_fixed_64x48": ["--nx", "64", "--ny", "48", "--integrator", "2", "--sts_method", "0",
                                    "--extsts_method", "3", "--d", "0.1", "--fixed_h", "0.002", "--tf", "0.02"],
    "ark_ars_adaptive_48": ["--nx", "48", "--ny", "48", "--integrator", "1", "--table_id", "1", "--rtol", "1e-4", "--tf", "0.02"],
    "ark_default_adaptive_48": ["--nx", "48", "--ny", "48", "--integrator", "1", "--order", "3", "--rtol", "1e-5", "--tf", "0.02"],
    # --implicit-reaction: Newton + band LU of I - gamma*J_reaction in the reference (SetupStrang / SetupExtSTS)
    "strang_rkc_implreact_48": ["--nx", "48", "--ny", "48", "--integrator", "3", "--sts_method", "0", "--fixed_h", "0.01",
                                "--tf", "0.05", "--implicit-reaction"],
    "strang_rkl_implreact_noadv_40x32": ["--nx", "40", "--ny", "32", "--integrator", "3", "--sts_method", "1", "--fixed_h", "0.005",
                                         "--tf", "0.03", "--implicit-reaction", "--no-advection"],
    "extsts_ars_implreact_fixed_48": ["--nx", "48", "--ny", "48", "--integrator", "2", "--sts_method", "0", "--extsts_method", "0",
                                      "--fixed_h", "0.005", "--tf", "0.05", "--implicit-reaction"],
    "extsts_giraldo_implreact_noadv_48": ["--nx", "48", "--ny", "48", "--integrator", "2", "--sts_method", "1", "--extsts_method", "1",
                                          "--rtol", "1e-4", "--tf", "0.05", "--implicit-reaction", "--no-advection"],
    "extsts_sdirk_implreact_fixed_64x32": ["--nx", "64", "--ny", "32", "--integrator", "2", "--sts_method", "0", "--extsts_method", "4",
                                           "--fixed_h", "0.005", "--tf", "0.05", "--implicit-reaction", "--no-advection"],
    "extsts_giraldo_implreact_fixed_64": ["--nx", "64", "--ny", "64", "--integrator", "2", "--sts_method", "0", "--extsts_method", "1",
                                          "--fixed_h", "0.004", "--tf", "0.04", "--implicit-reaction"],
}


def parse_lsrk_log(path):
    steps = []
    cur = None
    lines = open(path).read().split("\n")
    k = 0

    def nextval():
        return float(lines[k + 1].strip())

    while k < len(lines):
        ln = lines[k]
        m = re.search(r"begin-step-attempt\] step = (\d+), tn = ([-+0-9.eE]+), h = ([-+0-9.eE]+)", ln)
        if m:
            cur = {"step": int(m.group(1)), "tn": float(m.group(2)), "h": float(m.group(3)), "F": {}, "stages": None,
                   "spectral_radius": None}
            steps.append(cur)
        m = re.search(r"spectral radius = ([-+0-9.eE]+), num stages = (\d+)", ln)
        if m and cur is not None:
            cur["spectral_radius"] = float(m.group(1))
            cur["stages"] = int(m.group(2))
        if cur is not None:
            if "z_0(:) =" in ln:
                cur["z0"] = nextval()
            m = re.search(r"F_(\d+)\(:\) =", ln)
            if m:
                cur["F"][int(m.group(1))] = nextval()
            if "F_n(:) =" in ln:
                cur["Fn"] = nextval()
            if "ycur(:) =" in ln:
                cur["ycur"] = nextval()
            m = re.search(r"end-step-attempt\] status = success, dsm = ([-+0-9.eE]+)", ln)
            if m:
                cur["dsm"] = float(m.group(1))
        k += 1
    for s in steps:
        s["F"] = [s["F"][i] for i in sorted(s["F"])]
    return steps


def main(only_missing=False):
    out = {"source": "deps/sundials/test/unit_tests/logging/test_logging_arkode_lsrkstep_lvl5_{0..5}.out",
           "problem": "prv.hpp: y' = L(t)(y - atan t) + 1/(1+t^2), L(t) = -1000 - 10 cos((10-t)/10 pi); dom_eig = L(t)",
           "rtol": 1e-6, "atol": 1e-10, "methods": {}}
    for idx, name in enumerate(["rkc", "rkl", "ssps2", "ssps3", "ssp43", "ssp104"]):
        out["methods"][name] = parse_lsrk_log(os.path.join(LOGDIR, "test_logging_arkode_lsrkstep_lvl5_%d.out" % idx))
    with open(os.path.join(HERE, "lsrk_logging_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("lsrk_logging_golden.json:", {k: len(v) for k, v in out["methods"].items()})

    for name, args in D2D_CASES.items():
        if only_missing and os.path.exists(os.path.join(HERE, "d2d_%s.json" % name)):
            continue
        full = args + ["--nout", "1", "--output", "2"]
        nx, ny = int(cr.get_arg(full, "--nx", 64)), int(cr.get_arg(full, "--ny", 64))
        wd, text = cr.run(cr.REF_BIN, full, 1)
        t, u = cr.read_solution(wd, nx, ny)
        stats = cr.parse_stats(text)
        stats.pop("sim_time", None)
        m = re.search(r"Maximum relative error = ([-+0-9.eE]+)", text)
        if m:
            stats["max_rel_error"] = float(m.group(1))
        wd4, text4 = cr.run(cr.REF_BIN, full, 4)
        t4, u4 = cr.read_solution(wd4, nx, ny)
        spread = float(np.linalg.norm(u - u4) / np.linalg.norm(u))
        np.save(os.path.join(HERE, "d2d_%s.npy" % name), u)
        meta = {"args": args, "t_final": t, "stats": stats, "stats_np4": {k: v for k, v in cr.parse_stats(text4).items() if k != "sim_time"},
                "ref_np1_vs_np4_rel_l2": spread, "state_digits": 16}
        with open(os.path.join(HERE, "d2d_%s.json" % name), "w") as f:
            json.dump(meta, f, indent=1)
        print("%-32s steps=%s evals=%s np1-vs-np4=%.2e" % (name, stats.get("steps"), stats.get("rhs_evals", stats.get("rhs_evals_i")), spread))


def main_adr(only_missing=False):
    import compare_adr as ca

    for name, args in ADR_CASES.items():
        if only_missing and os.path.exists(os.path.join(HERE, "adr_%s.json" % name)):
            continue
        full = args + ["--nout", "1", "--output", "1"]
        nx, ny = int(ca.get_arg(full, "--nx", 400)), int(ca.get_arg(full, "--ny", 400))
        wd, text = ca.run(ca.REF_BIN, full)
        t, y = ca.read_solution(wd, nx, ny)
        np.save(os.path.join(HERE, "adr_%s.npy" % name), y)
        meta = {"args": args, "t_final": t, "stats": ca.parse_stats(text), "state_digits": 15}
        with open(os.path.join(HERE, "adr_%s.json" % name), "w") as f:
            json.dump(meta, f, indent=1)
        print("%-34s %s" % (name, meta["stats"]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "adr":
        main_adr()
    elif len(sys.argv) > 1 and sys.argv[1] == "adr-missing":
        main_adr(only_missing=True)
    elif len(sys.argv) > 1 and sys.argv[1] == "missing":
        main(only_missing=True)
    else:
        main()
        main_adr()
