#!/usr/bin/env python
"""Run the B200 adr driver and the reference CPU binary (oracle/_ref/adr2d_ref) on the same
command line and compare integrator statistics and the final state (solution.dat).

usage: python scripts/compare_adr.py -- <advection_diffusion_reaction_2d args...>
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200_BIN = os.path.join(ROOT, "ceda-demonstrations_b200", "bin", "adr2d_b200")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "adr2d_ref")

SECTIONS = {"Strang Integrator:": "outer", "ExtSTS Integrator:": "outer", "ARKStep Stepper:": "ark",
            "LSRKStep Stepper:": "lsrk", "Inner STS Method:": "lsrk", "Final integrator statistics:": "outer"}
KEYS = {"Steps": "steps", "Step attempts": "attempts", "Error test fails": "err_fails", "RHS fn evals": "rhs_evals",
        "RHS evals": "rhs_evals", "Explicit RHS fn evals": "rhs_evals_e", "Implicit RHS fn evals": "rhs_evals_i",
        "Explicit RHS evals": "rhs_evals_e", "Implicit RHS evals": "rhs_evals_i",
        "Explicit slow RHS fn evals": "rhs_evals_e", "Implicit slow RHS fn evals": "rhs_evals_i",
        "Number of dom_eig updates": "dom_eig_updates", "Max. num. of stages used": "max_stages",
        "Partition 1 evolves": "p1_evolves", "Partition 2 evolves": "p2_evolves",
        "NLS iters": "nls_iters", "LS iters": "lin_iters"}


def parse_stats(text):
    """{section.key: int} for the counters both drivers print with ARKodePrintAllStats."""
    out, sec = {}, "outer"
    for line in text.splitlines():
        if line.strip() in SECTIONS:
            sec = SECTIONS[line.strip()]
            continue
        m = re.match(r"^\s*([A-Za-z_. 0-9]+?)\s+=\s+(-?\d+)\s*$", line)
        if m and m.group(1).strip() in KEYS:
            out[sec + "." + KEYS[m.group(1).strip()]] = int(m.group(2))
    return out


def read_solution(workdir, nx, ny):
    """solution.dat: t, all u (row-major), all v  ->  (t, interleaved [u,v] array of 2*nx*ny)."""
    vals = np.array(open(os.path.join(workdir, "solution.dat")).read().split(), dtype=np.float64)
    y = np.empty(2 * nx * ny)
    y[0::2] = vals[1:1 + nx * ny]
    y[1::2] = vals[1 + nx * ny:]
    return vals[0], y


def run(binary, args, timeout=3600):
    workdir = tempfile.mkdtemp(prefix="adr_")
    r = subprocess.run([binary] + list(args), cwd=workdir, capture_output=True, text=True, timeout=timeout)
    if r.returncode:
        raise RuntimeError("%s failed rc=%d\n%s\n%s" % (binary, r.returncode, r.stdout[-3000:], r.stderr[-3000:]))
    return workdir, r.stdout


def get_arg(args, flag, default):
    return args[args.index(flag) + 1] if flag in args else default


def compare(args, verbose=True):
    args = list(args)
    nx, ny = int(get_arg(args, "--nx", 400)), int(get_arg(args, "--ny", 400))
    wd_g, out_g = run(B200_BIN, args)
    wd_c, out_c = run(REF_BIN, args)
    sg, sc = parse_stats(out_g), parse_stats(out_c)
    tg, yg = read_solution(wd_g, nx, ny)
    tc, yc = read_solution(wd_c, nx, ny)
    res = {"gpu": sg, "cpu": sc, "rel_l2": float(np.linalg.norm(yg - yc) / np.linalg.norm(yc)),
           "max_abs": float(np.max(np.abs(yg - yc))), "identical_15_digits": bool(np.array_equal(yg, yc))}
    if verbose:
        print("args:", " ".join(args))
        for k in sorted(set(sg) | set(sc)):
            print("  %-22s gpu=%-12s cpu=%-12s %s" % (k, sg.get(k), sc.get(k), "" if sg.get(k) == sc.get(k) else "<-- differs"))
        print("  rel_l2=%.3e max_abs=%.3e identical(15 digits)=%s" % (res["rel_l2"], res["max_abs"], res["identical_15_digits"]))
    return res


if __name__ == "__main__":
    argv = sys.argv[1:]
    if argv and argv[0] == "--":
        argv = argv[1:]
    compare(argv)
