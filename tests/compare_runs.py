#!/usr/bin/env python
"""Run the B200 driver and the reference CPU binary (oracle/_ref) on the same command
line and compare integrator statistics and the final state.

usage: python scripts/compare_runs.py [--np P] -- <diffusion_2D args...>

Test infrastructure (it executes oracle/_ref): used by the GPU parity tests
(tests/test_parity_gpu.py) and by hand.
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


_text_lib = None


def text_lib():
    """oracle/liboracle_sts.so's text helpers (oracle/ref_text.c): the reference's only state output is
    16-digit text, one line per output time; at 10^7..10^8 values numpy parsing takes minutes."""
    global _text_lib
    if _text_lib is None:
        import ctypes

        lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle_sts.so"))
        lib.orc_text_last_line.restype = ctypes.c_long
        lib.orc_text_last_line.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        lib.orc_text_print_mismatches.restype = ctypes.c_long
        lib.orc_text_print_mismatches.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        _text_lib = lib
    return _text_lib


def read_solution_fast(workdir, nx, ny):
    """read_solution for large grids: header lines in Python, the last data line of every rank file in C."""
    import ctypes

    lib = text_lib()
    u = np.full((ny, nx), np.nan)
    t_final = None
    for name in sorted(os.listdir(workdir)):
        if not name.startswith("diffusion_2d_solution."):
            continue
        path = os.path.join(workdir, name)
        hdr = {}
        with open(path) as f:
            while True:
                line = f.readline(256)
                if not line.startswith("#"):
                    break
                parts = line[1:].split()
                if len(parts) >= 2:
                    hdr[parts[0]] = parts[1]
        i0, i1, j0, j1 = int(hdr["is"]), int(hdr["ie"]), int(hdr["js"]), int(hdr["je"])
        n = (j1 - j0 + 1) * (i1 - i0 + 1)
        vals = np.empty(n)
        t = ctypes.c_double()
        got = lib.orc_text_last_line(path.encode(), vals.ctypes.data, n, ctypes.byref(t))
        if got != n:
            raise RuntimeError("%s: expected %d values in the last line, found %d" % (path, n, got))
        t_final = t.value
        u[j0 : j1 + 1, i0 : i1 + 1] = vals.reshape(j1 - j0 + 1, i1 - i0 + 1)
    return t_final, u


def print_mismatches(ours, ref):
    """How many values of `ours` do not print ("%.15e", diffusion_2D.cpp:819-824) to the characters the reference
    printed (`ref` = its output parsed back)."""
    a = np.ascontiguousarray(ours, dtype=np.float64).ravel()
    b = np.ascontiguousarray(ref, dtype=np.float64).ravel()
    assert a.size == b.size
    return int(text_lib().orc_text_print_mismatches(a.ctypes.data, b.ctypes.data, a.size))


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
