#!/usr/bin/env python
"""The reference's evaluation matrix for diffusion_2D on the B200 driver.

Same matrix, same command-line flags and same result columns as
/root/reference/diffusion_2D/runtests-diffusion2d.py:72-145 (adaptive: solver x grid x kx x rtol;
fixed step: solver x grid x kx x h; common flags --inhomogeneous --atol 1e-11 --controller 2 --error
--nonlinear --msbp 1 --maxsteps 100000 --internaleig), but the executable is
ceda-demonstrations_b200/bin/diffusion_2D_b200 on one GPU (the reference launches 1..64 MPI ranks per
grid) and, with --cpu, the unmodified reference build oracle/_ref/diffusion_2D_ref beside it.
Writes one CSV per series (the reference writes .xlsx through pandas).

    python scripts/runtests_diffusion2d_b200.py --solvers rkc,rkl --out profiles/r01_sweep
"""
import argparse
import csv
import ctypes
import os
import re
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BIN = os.path.join(ROOT, "ceda-demonstrations_b200", "bin", "diffusion_2D_b200")
CPU_BIN = os.path.join(ROOT, "oracle", "_ref", "diffusion_2D_ref")

SOLVERS = {  # the hypre variants of the reference matrix need the un-vendored hypre: out of scope
    "dirk2-Jacobi": ["--integrator", "dirk", "--order", "2"],
    "dirk3-Jacobi": ["--integrator", "dirk", "--order", "3"],
    "erk2": ["--integrator", "erk", "--order", "-2"],
    "erk3": ["--integrator", "erk", "--order", "-3"],
    "erk4": ["--integrator", "erk", "--order", "-4"],
    "rkc": ["--integrator", "rkc"],
    "rkl": ["--integrator", "rkl"],
}
COMMON = ["--inhomogeneous", "--atol", "1.e-11", "--controller", "2", "--error", "--nonlinear", "--msbp", "1",
          "--maxsteps", "100000", "--internaleig"]
KX = [0.1, 1.0, 10.0]
GRIDS = [(1, 32), (4, 64), (16, 128), (64, 256)]  # (MPI ranks the reference uses, grid)
RTOLS = [1e-2, 1e-3, 1e-4, 1e-5, 1e-6]
HVALS = [1e-2 / d for d in (2.0, 4.0, 8.0, 16.0, 32.0, 64.0)]

PATTERNS = {
    "Steps": re.compile(r"^Steps\s+=\s+(\d+)", re.M),
    "Fails": re.compile(r"^Error test fails\s+=\s+(\d+)", re.M),
    "Accuracy": re.compile(r"^Maximum relative error\s+=\s+(\S+)", re.M),
    "Runtime": re.compile(r"^Total simulation time\s+=\s+(\S+)", re.M),
}
EVALS = re.compile(r"^(?:Explicit RHS fn evals|Implicit RHS fn evals|LS RHS fn evals|RHS fn evals)\s+=\s+(\d+)", re.M)


_INPROC = {}


def run_in_process(args):
    """The B200 driver's main() (b200_d2d_main, the reference main() sequence) called in THIS process through the C-ABI,
    its stdout captured through a temporary file: one CUDA context for the whole matrix instead of one process start
    (about 1.5 s) per row."""
    if "lib" not in _INPROC:
        sys.path.insert(0, ROOT)
        import importlib

        _INPROC["lib"] = importlib.import_module("ceda-demonstrations_b200").sundials_lib()
    lib = _INPROC["lib"]
    argv = [b"diffusion_2D_b200"] + [a.encode() for a in args]
    arr = (ctypes.c_char_p * len(argv))(*argv)
    sys.stdout.flush()
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        saved = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            rc = lib.b200_d2d_main(len(argv), arr)
            libc = ctypes.CDLL(None)
            libc.fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return rc, tmp.read().decode(errors="replace")


def run_one(binary, solver, grid, rtol, h, kx, env=None, timeout=1800, inprocess=False):
    args = ["--nx", str(grid), "--ny", str(grid), "--rtol", "%e" % rtol, "--kx", "%e" % kx, "--ky", "%e" % 0.0]
    args += SOLVERS[solver] + COMMON + ["--output", "1", "--nout", "1"]
    if h > 0:
        args += ["--fixedstep", "%e" % h]
    row = {"method": solver, "grid": grid, "rtol": rtol, "h": h, "kx": kx, "ky": 0.0, "ReturnCode": 1,
           "Steps": None, "Fails": None, "Accuracy": None, "FEvals": None, "Runtime": None, "Wall": None}
    t0 = time.time()
    if inprocess:
        rc, out = run_in_process(args)
    else:
        try:
            res = subprocess.run([binary] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd="/tmp")
        except subprocess.TimeoutExpired:
            row["ReturnCode"] = -9
            return row
        rc, out = res.returncode, res.stdout
    row["Wall"] = time.time() - t0
    row["ReturnCode"] = rc
    if rc == 0:
        for k, pat in PATTERNS.items():
            m = pat.search(out)
            if m:
                row[k] = float(m.group(1)) if k in ("Accuracy", "Runtime") else int(m.group(1))
        row["FEvals"] = sum(int(v) for v in EVALS.findall(out))
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--solvers", default="rkc,rkl")
    ap.add_argument("--grids", default="32,64,128,256")
    ap.add_argument("--series", default="adaptive,fixed")
    ap.add_argument("--cpu", action="store_true", help="also run the unmodified reference build (oracle/_ref) on the host")
    ap.add_argument("--only-cpu", action="store_true", help="run ONLY the reference build (no GPU needed)")
    ap.add_argument("--cpu-ranks", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=1, help="reference rows run concurrently (one host core each)")
    ap.add_argument("--inprocess", action="store_true",
                    help="B200 arm: call the driver's main() in this process (one CUDA context for the whole matrix)")
    ap.add_argument("--rtols", default=",".join("%g" % r for r in RTOLS))
    ap.add_argument("--hdivs", default="2,4,8,16,32,64", help="fixed-step series: h = 1e-2 / d")
    ap.add_argument("--kx", default=",".join("%g" % k for k in KX))
    ap.add_argument("--timeout", type=int, default=1800)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep"))
    args = ap.parse_args()
    solvers = args.solvers.split(",")
    grids = [int(g) for g in args.grids.split(",")]
    rtols = [float(r) for r in args.rtols.split(",")]
    hvals = [1e-2 / float(d) for d in args.hdivs.split(",")]
    kxs = [float(k) for k in args.kx.split(",")]
    arms = [] if args.only_cpu else [("b200", GPU_BIN, None)]
    if args.cpu or args.only_cpu:
        arms.append(("reference", CPU_BIN, dict(os.environ, MPISHIM_NP=str(args.cpu_ranks))))
    for series in args.series.split(","):
        knob = rtols if series == "adaptive" else hvals
        jobs = []
        for kx in kxs:
            for val in knob:
                for grid in grids:
                    for solver in solvers:
                        for arm, binary, env in arms:
                            rtol, h = (val, 0.0) if series == "adaptive" else (1e-9, val)
                            jobs.append((arm, binary, solver, grid, rtol, h, kx, env))

        def work(job):
            arm, binary, solver, grid, rtol, h, kx, env = job
            row = run_one(binary, solver, grid, rtol, h, kx, env, timeout=args.timeout,
                          inprocess=(args.inprocess and arm == "b200"))
            row["arm"] = arm
            return row

        gpu_jobs = [j for j in jobs if j[0] == "b200"]
        cpu_jobs = [j for j in jobs if j[0] != "b200"]
        rows = [work(j) for j in gpu_jobs]
        if cpu_jobs:
            with ThreadPoolExecutor(max_workers=max(1, args.jobs)) as ex:
                rows += list(ex.map(work, cpu_jobs))
        path = "%s_%s.csv" % (args.out, series)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=["arm", "method", "grid", "rtol", "h", "kx", "ky", "ReturnCode", "Steps", "Fails",
                                              "Accuracy", "FEvals", "Runtime", "Wall"])
            w.writeheader()
            w.writerows(rows)
        ok = sum(1 for r in rows if r["ReturnCode"] == 0)
        print("%s: %d runs (%d ok) -> %s" % (series, len(rows), ok, path))


if __name__ == "__main__":
    main()
