#!/usr/bin/env python
"""Run the B200 driver and the reference CPU binary (oracle/_ref) on the same command
line and compare integrator statistics and the final state.

usage: python scripts/compare_runs.py [--np P] -- <diffusion_2D args...>

Used by the GPU parity tests (tests/test_parity_gpu.py) and by hand.
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200_BIN = os.path.join(ROOT, "ceda-demonstrations_b200", "bin", "diffusion_2D_b200")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "diffusion_2D_ref")

STAT_KEYS = {
    "Steps": "steps",
    "Step attempts": "attempts",
    "Error test fails": "err_fails",
    "RHS fn evals": "rhs_evals",
    "Explicit RHS fn evals": "rhs_evals_e",
    "Implicit RHS fn evals": "rhs_evals_i",
    "Number of dom_eig updates": "dom_eig_updates",
    "Max. num. of stages used": "max_stages",
    "Number of fe calls for DEE": "dee_evals",
    "Number of DEE iters": "dee_iters",
    "LS iters": "lin_iters",
    "NLS iters": "nls_iters",
    "Prec evals": "prec_evals",
    "Prec solves": "prec_solves",
}


def parse_stats(text):
    out = {}
    for line in text.splitlines():
        m = re.match(r"^\s*([A-Za-z_.\s]+?)\s+=\s+([-+0-9.eE]+)\s*$", line)
        if not m:
            continue
        key = m.group(1).strip()
        if key in STAT_KEYS:
            out[STAT_KEYS[key]] = float(m.group(2)) if "." in m.group(2) or "e" in m.group(2) else int(m.group(2))
        elif key == "Total simulation time":
            out["sim_time"] = float(m.group(2))
        elif key == "Max. spectral radius":
            out["sr_max"] = float(m.group(2))
    return out


def read_solution(workdir, nx, ny):
    """Assemble the global final state from the per-rank diffusion_2d_solution.NNNNN.txt files."""
    u = np.full((ny, nx), np.nan)
    t_final = None
    for name in sorted(os.listdir(workdir)):
        if not name.startswith("diffusion_2d_solution."):
            continue
        hdr = {}
        last = None
        with open(os.path.join(workdir, name)) as f:
            for line in f:
                if line.startswith("#"):
                    parts = line[1:].split()
                    if len(parts) >= 2:
                        hdr[parts[0]] = parts[1]
                elif line.strip():
                    last = line
        vals = np.array(last.split(), dtype=np.float64)
        t_final = vals[0]
        i0, i1, j0, j1 = int(hdr["is"]), int(hdr["ie"]), int(hdr["js"]), int(hdr["je"])
        u[j0 : j1 + 1, i0 : i1 + 1] = vals[1:].reshape(j1 - j0 + 1, i1 - i0 + 1)
    return t_final, u


def run(binary, args, np_ranks=1, env_extra=None, timeout=600):
    workdir = tempfile.mkdtemp(prefix="d2d_")
    env = dict(os.environ)
    if env_extra:
        env.update(env_extra)
    if binary == REF_BIN:
        env["MPISHIM_NP"] = str(np_ranks)
        procs = [subprocess.Popen([binary] + args, cwd=workdir, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)]
    else:
        procs = []
        idfile = os.path.join(workdir, "nccl_id")
        for r in range(np_ranks):
            e = dict(env)
            e.update({"B200_RANK": str(r), "B200_NP": str(np_ranks), "B200_DEVICE": str(r), "B200_NCCL_ID_FILE": idfile})
            e.pop("RANK", None)
            e.pop("WORLD_SIZE", None)
            e.pop("LOCAL_RANK", None)
            procs.append(subprocess.Popen([binary] + args, cwd=workdir, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    try:
        outs = [p.communicate(timeout=timeout)[0] for p in procs]
    except subprocess.TimeoutExpired:
        for p in procs:  # a hung rank must not outlive the test (exact PIDs we started)
            p.kill()
        raise
    rcs = [p.returncode for p in procs]
    if any(rcs):
        raise RuntimeError("%s failed rc=%s\n%s" % (binary, rcs, "\n".join(outs)[-4000:]))
    return workdir, outs[0]


def get_arg(args, flag, default):
    return args[args.index(flag) + 1] if flag in args else default


def compare(args, np_gpu=1, np_cpu=1, verbose=True):
    args = list(args)
    if "--output" not in args:
        args += ["--output", "2"]
    nx, ny = int(get_arg(args, "--nx", 64)), int(get_arg(args, "--ny", 64))
    wd_g, out_g = run(B200_BIN, args, np_gpu)
    wd_c, out_c = run(REF_BIN, args, np_cpu)
    sg, sc = parse_stats(out_g), parse_stats(out_c)
    tg, ug = read_solution(wd_g, nx, ny)
    tc, uc = read_solution(wd_c, nx, ny)
    rel_l2 = float(np.linalg.norm(ug - uc) / np.linalg.norm(uc))
    max_abs = float(np.max(np.abs(ug - uc)))
    identical = bool(np.array_equal(ug, uc))
    res = {"gpu": sg, "cpu": sc, "rel_l2": rel_l2, "max_abs": max_abs, "identical_16_digits": identical,
           "t_gpu": tg, "t_cpu": tc}
    if verbose:
        print("args:", " ".join(args))
        keys = sorted(set(sg) | set(sc))
        for k in keys:
            print("  %-18s gpu=%-22s cpu=%-22s %s" % (k, sg.get(k), sc.get(k), "" if sg.get(k) == sc.get(k) or k == "sim_time" else "<-- differs"))
        print("  rel_l2=%.3e max_abs=%.3e identical(16 digits)=%s t=%s/%s" % (rel_l2, max_abs, identical, tg, tc))
    return res


if __name__ == "__main__":
    argv = sys.argv[1:]
    npg = 1
    if argv and argv[0] == "--np":
        npg = int(argv[1])
        argv = argv[2:]
    if argv and argv[0] == "--":
        argv = argv[1:]
    compare(argv, np_gpu=npg, np_cpu=1)
