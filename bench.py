#!/usr/bin/env python
"""bench.py -- cell-updates/s of the explicit STS hot path (diffusion_2D, RKC, 16384^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun ... bench.py --gpus N --steps K --warmup W        (one rank per GPU)

Workload (BASELINE.json configs[2], SURVEY.md section 8d "C3 throughput"):
    diffusion_2D --nx 16384 --ny 16384 --integrator rkc --fixedstep 1e-4
    => rho = 1.01*8/dx^2, s = 92 RKC stages per step, one "step" = one LSRKStep time step
       (92 stage evaluations + the closing RHS evaluation = 93 RHS evals).
Weak scaling: every GPU owns a 16384 x 16384 block; the global grid is
(16384*npx) x (16384*npy) with dims = MPI_Dims_create(N) (2->2x1, 4->2x2, 8->4x2) and the
domain enlarged so dx, dy and therefore the stage count stay fixed.

metric  cell-updates/s = RHS evaluations (counted by ARKODE) x global cells / device time.
value   state resident in HBM, timed with CUDA events on the launching stream, max over ranks.
e2e     every step: pinned-host state -> device (H2D), ARKodeReset, one time step, device -> pinned
        host (D2H), through the public session API (b200_d2d_*), host wall clock.
roofline dominant kernel = k_chain_march<K> (K temporally blocked RKC stages per launch: fused stencil
        + 5-term recurrence, 4 streamed reads + 2 writes per K cell-updates); achieved = modelled bytes of
        all launches in the timed region / device time.  The 40 B per cell-update one-pass-per-stage basis
        of SURVEY.md section 8d is reported beside it.
"""
import argparse
import ctypes
import importlib
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_UPDATE = 40.0
H_FIXED = 1.0e-4
DEFAULT_ARITH = "exact"
XL, XU0, YL, YU0 = -3.141592653589793, 3.141592653589793, -6.0, 6.0


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def dims_create(n):
    b, f = 1, 1
    while f * f <= n:
        if n % f == 0:
            b = f
        f += 1
    return n // b, b


def workload_args(local_n, npx, npy, method="rkc", base_n=None):
    """Reference-style flags for a (local_n*npx) x (local_n*npy) grid whose spacing equals that of
    the single-GPU base_n^2 grid on the default domain (so rho, hence the stage count, is fixed)."""
    base_n = base_n or local_n
    dx = (XU0 - XL) / (base_n - 1)
    dy = (YU0 - YL) / (base_n - 1)
    nx, ny = local_n * npx, local_n * npy
    xu = XL + dx * (nx - 1)
    yu = YL + dy * (ny - 1)
    return ["--nx", str(nx), "--ny", str(ny), "--xu", repr(xu), "--yu", repr(yu),
            "--integrator", method, "--fixedstep", repr(H_FIXED), "--tf", "1.0", "--nout", "1",
            "--output", "0", "--npx", str(npx), "--npy", str(npy)]


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md's clocks line).  Sampled through
    NVML in this process (nvidia_ml_py): spawning nvidia-smi several times inside a 0.3 s timed region stalled the
    launching thread for tens of milliseconds (round-2 measurement: 52 -> 83 ms per step with 4 nvidia-smi calls inside
    the region); nvidia-smi remains the fallback when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []  # (sm_mhz, sm_max_mhz, set of active reasons)
        self.extra = []    # (mem_mhz, gpu_temp_C, power_W) when NVML is available
        self._stop_evt = threading.Event()
        self.source = "nvidia-smi"
        self._nvml = self._handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = None
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml, self._handle, self.source = pynvml, handle, "nvml"
        except Exception:
            pass

    def _sample_nvml(self):
        nv, h = self._nvml, self._handle
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        self.samples.append((int(sm), int(mx), {n for n, b in self.BITS.items() if mask & b}))
        try:
            self.extra.append((int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_MEM)),
                               int(nv.nvmlDeviceGetTemperature(h, nv.NVML_TEMPERATURE_GPU)),
                               nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
        except Exception:
            pass

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 7 and parts[0].replace(".", "").isdigit() and parts[1].replace(".", "").isdigit():
            self.samples.append((int(float(parts[0])), int(float(parts[1])),
                                 {n for k, n in enumerate(self.NAMES) if parts[3 + k].lower().startswith("active")}))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.1 if self._nvml is not None else 0.5)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples]
        reasons = [n for n in self.NAMES if any(n in s[2] for s in self.samples)]
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(self.samples), "source": self.source}
        if self.extra:
            out["mem_mhz"] = sorted(e[0] for e in self.extra)[len(self.extra) // 2]
            out["gpu_temp_c"] = max(e[1] for e in self.extra)
            out["power_w"] = round(sorted(e[2] for e in self.extra)[len(self.extra) // 2], 1)
        return out


# ------------------------------------------------------------------------------ reference arm
def parse_ref(out):
    t = float(re.search(r"Total simulation time\s*=\s*([-+0-9.eE]+)", out).group(1))
    evals = int(re.search(r"RHS fn evals\s*=\s*(\d+)", out).group(1))
    return t, evals


def run_reference_sample(nsteps, sample_n, cores, base_n=16384):
    """The reference's own MPI CPU path (oracle/_ref/diffusion_2D_ref: unmodified sources +
    the in-tree MPI shim) on `cores` ranks, on a sample_n^2 block of the workload (same dx, dy,
    h => same 92 stages)."""
    binary = os.path.join(ROOT, "oracle", "_ref", "diffusion_2D_ref")
    if not os.path.exists(binary):
        raise RuntimeError("oracle/_ref/diffusion_2D_ref missing (built by __graft_entry__.build())")
    args = workload_args(sample_n, 1, 1, base_n=base_n)
    args[args.index("--tf") + 1] = repr(nsteps * H_FIXED)
    args = [a for a in args]
    # let the reference pick its own process grid for `cores` ranks
    for flag in ("--npx", "--npy"):
        k = args.index(flag)
        del args[k:k + 2]
    env = dict(os.environ, MPISHIM_NP=str(cores))
    out = subprocess.run([binary] + args, env=env, capture_output=True, text=True, timeout=3000).stdout
    t, evals = parse_ref(out)
    return evals * sample_n * sample_n / t, t, evals


def cpu_sample_plan(ranks, target_s=10.0):
    """(sample_n, nsteps) so that the reference's CPU path works for about target_s seconds:
    ~7.5e7 cell-updates/s per core was measured for it on this pool's hosts (profiles/)."""
    sample_n = 4096 if ranks >= 8 else 2048
    per_step = 93.0 * sample_n * sample_n
    nsteps = int(round(target_s * 7.5e7 * ranks / per_step))
    return sample_n, max(1, min(nsteps, 40))


# ------------------------------------------------------------------------ host placement (N > 1)
def gpu_numa_node(index):
    """NUMA node the GPU's PCIe root hangs off (sysfs), or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:  # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node):
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:
        return set()


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs of its GPU's NUMA node, so that the pinned staging buffers it allocates afterwards
    are first-touched there (Linux's default local allocation) and its copies do not cross the socket link.
    Returns what was done (reported in the JSON line)."""
    node = gpu_numa_node(index)
    if node is None:
        return {"node": None, "why": "GPU NUMA node unknown"}
    allowed = os.sched_getaffinity(0)
    cpus = node_cpus(node) & allowed
    if not cpus:
        return {"node": node, "why": "no allowed CPU on that node (affinity mask %d CPUs)" % len(allowed)}
    os.sched_setaffinity(0, cpus)
    return {"node": node, "cpus": len(cpus)}


def pages_numa_node(tensor):
    """NUMA node of the first page of a host tensor (move_pages query), or None."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        page = ctypes.c_void_p(tensor.data_ptr() & ~4095)
        status = ctypes.c_int(-1)
        rc = libc.syscall(279, 0, ctypes.c_ulong(1), ctypes.byref(page), None, ctypes.byref(status), 0)  # move_pages
        return status.value if rc == 0 and status.value >= 0 else None
    except Exception:
        return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    ranks = 1
    while ranks * 2 <= min(cores, 64):
        ranks *= 2
    sample_n = 4096 if ranks >= 8 else 2048
    if os.environ.get("B200_BENCH_REF_SAMPLE_N"):  # tests/test_bench_contract.py: a seconds-long sample
        sample_n = int(os.environ["B200_BENCH_REF_SAMPLE_N"])
    if args.warmup > 0:
        run_reference_sample(1, sample_n, ranks)
    t0 = time.time()
    value, t, evals = run_reference_sample(max(args.steps, 1), sample_n, ranks)
    npx, npy = dims_create(args.gpus)
    sample = ("%d^2 block of the workload grid (same dx, dy, h=1e-4 => 92 RKC stages/step), %d steps, "
              "%d shared-memory MPI ranks of the unmodified reference build" % (sample_n, max(args.steps, 1), ranks))
    line = {
        "impl": "reference", "metric": "cell-updates/s (RHS evals x cells / s), diffusion_2D 16384^2 RKC",
        "value": value, "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "diffusion_2D 16384x16384 per GPU, LSRKStep RKC, fixedstep 1e-4 (92 stages/step)",
                   "global_grid": [16384 * npx, 16384 * npy], "parallelism": "%dx%d blocks" % (npx, npy)},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": ranks, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    emit(line)


# ------------------------------------------------------------------ the other BASELINE configs (--config)
# c3 (default) is the headline line above.  c2 / c4 / c5 are BASELINE.json configs[1] / [3] / [4] with the same line
# shape: `value` = state resident in HBM, CUDA events on the launching stream; `e2e` = every batch starts and ends in
# pinned host memory; `roofline` = modelled bytes of every launch in the timed region (b200_algorithmic_bytes: 8 B x
# entries x full vectors read + written) / device time against the measured HBM copy peak; `cpu_baseline` = the
# unmodified reference build on the host cores on a bounded sample of the same workload.
CONFIGS = {
    "c2": {
        "kind": "d2d", "steps": 25, "e2e_batch_steps": 5, "cells": 4096 * 4096,
        "args": ["--nx", "4096", "--ny", "4096", "--integrator", "rkl", "--kx", "1", "--ky", "0.1", "--inhomogeneous",
                 "--internaleig", "--tf", "1.0", "--nout", "1", "--output", "0"],
        "metric": "cell-updates/s ((RHS + DEE evals) x cells / s), diffusion_2D 4096^2 RKL2 anisotropic inhomogeneous, power-iteration eigenvalue",
        "workload": "diffusion_2D 4096x4096 FP64, kx=1 ky=0.1 inhomogeneous, LSRKStep RKL2 adaptive (rtol 1e-5), --internaleig "
                    "(power iteration every 25 steps: one falls into the default 25 timed steps)",
        "kernel": "k_chain_march<4> (table-driven coefficients) + k_dq_march (power-iteration difference quotients)",
        "ref": {"bin": "diffusion_2D_ref", "args": ["--nx", "4096", "--ny", "4096", "--integrator", "rkl", "--kx", "1", "--ky", "0.1",
                                                     "--inhomogeneous", "--internaleig", "--tf", "0.001", "--nout", "1", "--output", "0"],
                "cells": 4096 * 4096, "sample": "the full 4096^2 problem to tf = 1e-3 (19 steps)"},
    },
    "c4": {
        "kind": "adr", "steps": 5, "cells": 2048 * 2048,
        "args": ["--nx", "2048", "--ny", "2048", "--integrator", "3", "--sts_method", "0", "--fixed_h", "0.001", "--tf", "10.0",
                 "--nout", "1", "--output", "0"],
        "metric": "grid-point-updates/s (RHS evals x grid points / s, 2 species per point), adr 2D Brusselator 2048^2, Strang + RKC",
        "workload": "adr 2D advection-diffusion-reaction (Brusselator) 2048x2048 x 2 species, Strang splitting: RKC diffusion "
                    "half steps (17 stages) + explicit advection-reaction, fixed h = 1e-3",
        "kernel": "k_adr_march<2> / k_adr_chain (diffusion stages) + k_adr_march<5> (advection-reaction)",
        # --output 1: the reference prints its statistics and "Total solve time" (2 significant digits) only then
        "ref": {"bin": "adr2d_ref", "args": ["--nx", "2048", "--ny", "2048", "--integrator", "3", "--sts_method", "0", "--fixed_h", "0.001",
                                              "--nout", "1", "--output", "1"],
                "cells": 2048 * 2048, "sample": "the full 2048^2 problem, %d Strang steps (the reference adr driver is serial: 1 core)"},
    },
    "c5": {
        "kind": "d2d", "steps": 5, "e2e_batch_steps": 3, "cells": 8192 * 8192,
        "args": ["--nx", "8192", "--ny", "8192", "--integrator", "dirk", "--order", "3", "--tf", "1.0", "--nout", "1", "--output", "0"],
        "metric": "cell-updates/s ((implicit RHS + linear-solver RHS evals) x cells / s), diffusion_2D 8192^2 DIRK3 + PCG + Jacobi",
        "workload": "diffusion_2D 8192x8192 FP64, ARKStep DIRK order 3 adaptive (rtol 1e-5), matrix-free PCG (<= 20 iterations) "
                    "with Jacobi preconditioner",
        "kernel": "k_dq_march (A*p = p - gamma*J p by difference quotient, fused with <Ap,p>) + k_lin2_wsqr / k_prod_dot (PCG vector work)",
        "ref": {"bin": "diffusion_2D_ref", "args": ["--nx", "8192", "--ny", "1024", "--integrator", "dirk", "--order", "3", "--tf", "0.0008",
                                                     "--nout", "1", "--output", "0"],
                "cells": 8192 * 1024, "sample": "an 8192 x 1024 block (same dx, hence the same stiffness) to tf = 8e-4 (8 steps)"},
    },
}


def config_evals(kind, st):
    if kind == "adr":
        return st["lsrk_rhs_evals"] + st["ark_rhs_evals"] + st["rhs_evals_explicit"]
    return st["rhs_evals"] + st["dee_rhs_evals"] + st["lin_rhs_evals"]


def reference_config_sample(cfg, ranks, reps=1):
    """(value, seconds, evals) of the unmodified reference build on the host cores for this config's sample."""
    ref = cfg["ref"]
    binary = os.path.join(ROOT, "oracle", "_ref", ref["bin"])
    if not os.path.exists(binary):
        raise RuntimeError("oracle/_ref/%s missing (built by __graft_entry__.build())" % ref["bin"])
    rargs = list(ref["args"])
    if cfg["kind"] == "adr":
        rargs += ["--tf", repr(0.001 * max(reps, 1))]
        ranks = 1
    env = dict(os.environ, MPISHIM_NP=str(ranks))
    import tempfile

    t0 = time.time()
    out = subprocess.run([binary] + rargs, env=env, capture_output=True, text=True, timeout=3000, cwd=tempfile.mkdtemp(prefix="ref_")).stdout
    wall = time.time() - t0
    if cfg["kind"] == "adr":
        m = re.search(r"Total solve time\s*=\s*([-+0-9.eE]+)", out)
        t = float(m.group(1)) if m else wall
        evals = sum(int(v) for v in re.findall(r"RHS fn evals\s*=\s*(\d+)", out))
    else:
        t = float(re.search(r"Total simulation time\s*=\s*([-+0-9.eE]+)", out).group(1))
        evals = sum(int(v) for v in re.findall(r"(?:^|\n)\s*(?:RHS fn evals|Implicit RHS fn evals|LS RHS fn evals|Number of fe calls for DEE)\s*=\s*(\d+)", out))
    return evals * ref["cells"] / t, t, evals, ranks


def config_bench(args, cfgname):
    cfg = CONFIGS[cfgname]
    rank = int(os.environ.get("RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.gpus > 1:
        # these configurations are single-GPU measurements (the adr driver is serial in the reference; c2 is quoted on 1 B200);
        # c5's 8-GPU leg is the weak-scaling line of the default workload's process grid and is not built here
        if rank == 0:
            emit({"metric": cfg["metric"], "config": {"workload": cfg["workload"]}, "n_gpus": args.gpus,
                  "unavailable": "--config %s is measured on one GPU" % cfgname})
        return
    cores = host_cores()
    ranks = 1
    while ranks * 2 <= min(cores, 64):
        ranks *= 2
    steps = args.steps if args.steps_given else cfg["steps"]
    warmup = max(args.warmup, 3)
    base = {"metric": cfg["metric"], "unit": "cell-updates/s" if cfg["kind"] == "d2d" else "grid-point-updates/s", "n_gpus": 1,
            "steps": steps, "warmup": warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": cfg["workload"], "name": cfgname}}
    if args.impl == "reference":
        if rank != 0:
            return
        v, t, ev, used = reference_config_sample(cfg, ranks, reps=max(steps, 1))
        sample = (cfg["ref"]["sample"] % max(steps, 1)) if "%d" in cfg["ref"]["sample"] else cfg["ref"]["sample"]
        emit(dict(base, impl="reference", value=v, ms_per_step=1e3 * t / max(steps, 1), gpu_launches=0,
                  cpu_baseline={"value": v, "unit": base["unit"], "cores": used, "kind": "reference", "sample": sample + ", %d RHS evals, %.1f s" % (ev, t)},
                  e2e={"value": v, "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}))
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    b200 = importlib.import_module("ceda-demonstrations_b200")
    torch.cuda.set_device(0)
    placement = {"gpu_numa": gpu_numa_node(0), "bound": bind_to_gpu_numa_node(0) if os.environ.get("B200_BENCH_NO_BIND") is None else None}
    klib = b200.kernel_lib()
    klib.b200_algorithmic_bytes.restype = ctypes.c_uint64
    extra = list(args.config_args or [])
    prob = (b200.Adr2D if cfg["kind"] == "adr" else b200.Diffusion2D)(cfg["args"] + extra, device=0)
    nvals = cfg["cells"] * (2 if cfg["kind"] == "adr" else 1)

    prob.step(warmup)
    torch.cuda.synchronize()
    s0 = prob.stats()
    ab0 = klib.b200_algorithmic_bytes()
    sampler = ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    prob.step(steps)
    ev1.record()
    torch.cuda.synchronize()
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    s1 = prob.stats()
    alg_bytes = klib.b200_algorithmic_bytes() - ab0
    evals = config_evals(cfg["kind"], s1) - config_evals(cfg["kind"], s0)
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    value = evals * cfg["cells"] / (ms * 1e-3)

    # ---- e2e: every batch starts from a state in pinned host memory and ends with its result there
    e2e = None
    if not args.no_e2e:
        hin = [torch.empty(nvals, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        hout = [torch.empty(nvals, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        placement["pinned_pages_numa"] = pages_numa_node(hin[0])
        prob.get_state(hin[0])
        hin[1].copy_(hin[0])
        t_cur = prob.stats()["t"]
        bsteps = cfg.get("e2e_batch_steps", 1)
        nb = max(2, steps // bsteps) if cfg["kind"] == "d2d" else max(2, steps)
        if cfg["kind"] == "d2d":
            prob.run_batches([hin[0], hin[1]], [hout[0], hout[1]], t_cur, 1)  # staging buffers, copy streams
            e0 = prob.stats()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            prob.run_batches([hin[i % 2] for i in range(nb)], [hout[i % 2] for i in range(nb)], t_cur, bsteps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            mode = "pipelined: independent batches of %d steps, uploads / downloads of neighbouring batches overlap the integration" % bsteps
        else:
            bsteps = 1
            prob.set_state(hin[0], t_cur)
            e0 = prob.stats()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(nb):
                prob.set_state(hin[i % 2], t_cur)  # H2D + ARKodeReset
                prob.step(1)
                prob.get_state(hout[i % 2])  # D2H (synchronises)
            dt = time.perf_counter() - t0
            mode = "sequential: upload, one Strang step, download (the adr session has no batch pipeline)"
        e1 = prob.stats()
        e_evals = config_evals(cfg["kind"], e1) - config_evals(cfg["kind"], e0)
        e2e = {"value": e_evals * cfg["cells"] / dt, "unit": base["unit"], "h2d_bytes_per_step": 8 * nvals, "d2h_bytes_per_step": 8 * nvals,
               "ms_per_step": 1e3 * dt / (nb * bsteps), "batches": nb, "steps_per_batch": bsteps, "mode": mode, "host_placement": placement}
        del hin, hout
    stats_final = prob.stats()
    prob.close()

    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": cfg["kernel"], "modelled_bytes_timed": int(alg_bytes),
                "bytes_per_update": alg_bytes / max(evals * cfg["cells"], 1), "rhs_evals_timed": evals,
                "one_pass_per_stage_basis": {"bytes_per_update": ALG_BYTES_PER_UPDATE * (2 if cfg["kind"] == "adr" else 1),
                                             "achieved": ALG_BYTES_PER_UPDATE * (2 if cfg["kind"] == "adr" else 1) * evals * cfg["cells"] / (ms * 1e-3) / 1e9}}
    roofline["one_pass_per_stage_basis"]["frac"] = roofline["one_pass_per_stage_basis"]["achieved"] / peak
    cpu_baseline = None
    if not args.no_cpu_baseline:
        try:
            reps = 3 if cfg["kind"] == "adr" else 1
            v, t, ev, used = reference_config_sample(cfg, ranks, reps=reps)
            sample = (cfg["ref"]["sample"] % reps) if "%d" in cfg["ref"]["sample"] else cfg["ref"]["sample"]
            cpu_baseline = {"value": v, "unit": base["unit"], "cores": used, "kind": "reference",
                            "sample": sample + ", %d RHS evals, %.1f s" % (ev, t)}
        except Exception as exc:
            cpu_baseline = {"value": None, "unit": base["unit"], "cores": 0, "kind": "reference", "sample": "unavailable: %s" % exc}
    keep = ("steps", "step_attempts", "err_test_fails", "rhs_evals", "dee_rhs_evals", "lin_iters", "lin_rhs_evals", "max_stages",
            "lsrk_rhs_evals", "lsrk_max_stages", "ark_rhs_evals", "dq_fused", "ew_fused", "chain_launches", "chain_stages")
    emit(dict(base, value=value, ms_per_step=ms / steps, e2e=e2e, gpu_launches=int(launches), roofline=roofline,
              cpu_baseline=cpu_baseline, clocks=sampler.summary(),
              integrator_stats={k: stats_final[k] for k in keep if k in stats_final},
              launches_per_rhs_eval=launches / max(evals, 1)))


# ------------------------------------------------------------------------------------ our arm
def emit(line):
    """The contract is ONE JSON line on stdout: libraries (NCCL prints its version banner to stdout)
    write to the process' fd 1, which main() points at stderr; the result goes to the real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--local-n", type=int, default=16384, help="per-GPU block edge (default: the headline 16384)")
    ap.add_argument("--method", default="rkc")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chain", type=int, default=0, help="temporal-blocking depth (0 = library default)")
    ap.add_argument("--chain-variant", type=int, default=-1, help="0 = k_chain_march (default), 1 = k_chain_quad")
    ap.add_argument("--arith", default=DEFAULT_ARITH, choices=["exact", "fma"],
                    help="exact = bit-identical to the reference's baseline x86-64 build; fma = contracted multiply-adds")
    ap.add_argument("--e2e-mode", default="pipelined", choices=["pipelined", "sequential"])
    ap.add_argument("--chain-rows", type=int, default=0, help="rows per block of the chain kernel (0 = library default)")
    ap.add_argument("--config", default="c3", choices=["c3", "c2", "c4", "c5"],
                    help="c3 (default) = the headline 16384^2 RKC line; c2 / c4 / c5 = BASELINE configs[1] / [3] / [4]")
    ap.add_argument("--config-args", nargs=argparse.REMAINDER, help="extra driver flags for --config c2/c4/c5 (e.g. --sts_chain 6)")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 5
    if args.config != "c3":
        config_bench(args, args.config)
        return

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    b200 = importlib.import_module("ceda-demonstrations_b200")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    # one rank per GPU: keep the rank (and the pinned staging buffers it first-touches) on its GPU's NUMA node
    placement = {"gpu_numa": gpu_numa_node(local_rank), "bound": None}
    if os.environ.get("B200_BENCH_NO_BIND") is None:
        placement["bound"] = bind_to_gpu_numa_node(local_rank)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(b200.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    npx, npy = dims_create(world)
    n = args.local_n
    wargs = workload_args(n, npx, npy, method=args.method, base_n=n) + ["--arith", args.arith]
    if args.chain > 0:
        wargs += ["--chain", str(args.chain)]
    if args.chain_variant >= 0:
        wargs += ["--chain-variant", str(args.chain_variant)]
    prob = b200.Diffusion2D(wargs, rank=rank, nranks=world, nccl_id=nccl_id, device=local_rank)
    if args.chain_rows > 0:
        b200.kernel_lib().b200_set_chain_rows(args.chain_rows)
    ncell_global = (n * npx) * (n * npy)
    ncell_local = n * n

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: state resident in HBM -------------------------------------------------------
    prob.step(max(args.warmup, 3))
    barrier()
    s0 = prob.stats()
    klib = b200.kernel_lib()
    klib.b200_algorithmic_bytes.restype = ctypes.c_uint64
    ab0 = klib.b200_algorithmic_bytes()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    prob.step(args.steps)
    ev1.record()
    barrier()
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    alg_bytes = klib.b200_algorithmic_bytes() - ab0
    s1 = prob.stats()
    evals = s1["rhs_evals"] - s0["rhs_evals"]
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = evals * ncell_global / (ms_max * 1e-3)
    # per-rank view of the timed region (device time, SM clock, board power of every GPU): the job runs in lockstep
    # through the halo exchanges, so one slow GPU (a lower power-capped clock) sets everybody's time
    per_rank = None
    if dist is not None:
        cs = sampler.summary()
        mine = torch.tensor([ms, float(cs.get("sm_mhz") or 0), float(cs.get("power_w") or 0), float(cs.get("gpu_temp_c") or 0)],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rows = [r.cpu().tolist() for r in allr]
        per_rank = {"ms": [round(r[0], 2) for r in rows], "sm_mhz": [int(r[1]) for r in rows],
                    "power_w": [round(r[2], 1) for r in rows], "gpu_temp_c": [int(r[3]) for r in rows]}

    # diagnostic (B200_BENCH_PER_STEP=1, outside the timed region): the same steps once more, one call and one event
    # pair per step, so that a periodic hiccup or a slow first step shows up
    step_ms = None
    if os.environ.get("B200_BENCH_PER_STEP"):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        evs[0].record()
        for i in range(args.steps):
            prob.step(1)
            evs[i + 1].record()
        barrier()
        step_ms = [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(args.steps)]

    # ---- e2e: host buffers, H2D + step + D2H every step ---------------------------------------
    # Every timed step starts from a state in pinned HOST memory and ends with its result back in
    # pinned host memory, through the problem layer's public calls.
    #   pipelined (default): the steps are independent batches (b200_d2d_run_batches): the upload of
    #     batch i+1 and the download of batch i-1 run on copy streams while batch i is integrated;
    #     every batch's 2 x 8 B x cells cross PCIe inside the timed region.
    #   sequential: one dependent chain, set_state -> step -> get_state, nothing overlaps.
    e2e = None
    hin = hout = None
    if not args.no_e2e:
        # the pinned staging buffers (4 x 8 B x cells per rank) are the only host allocation of size; whether they
        # could be had is decided collectively so that no rank waits in a halo exchange for one that gave up
        hin = hout = None
        try:
            hin = [torch.empty(ncell_local, dtype=torch.float64, pin_memory=True) for _ in range(2)]
            hout = [torch.empty(ncell_local, dtype=torch.float64, pin_memory=True) for _ in range(2)]
            have = 1
        except Exception as exc:  # noqa: BLE001 -- reported in the JSON line
            have = 0
            sys.stderr.write("bench.py: pinned host allocation failed on rank %d: %s\n" % (rank, exc))
        flag = torch.tensor([have], dtype=torch.int32, device="cuda")
        if dist is not None:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            e2e = {"value": None, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "error": "pinned host memory for the staging buffers could not be allocated on every rank"}
            hin = hout = None
    if hin is not None:
        placement["pinned_pages_numa"] = pages_numa_node(hin[0])
        prob.get_state(hin[0])
        hin[1].copy_(hin[0])
        t_cur = prob.stats()["t"]
        if args.e2e_mode == "pipelined":
            prob.run_batches([hin[0], hin[1]], [hout[0], hout[1]], t_cur, 1)  # warm-up: staging buffers, streams
        e_s0 = prob.stats()
        barrier()
        t0 = time.perf_counter()
        if args.e2e_mode == "pipelined":
            prob.run_batches([hin[i % 2] for i in range(args.steps)], [hout[i % 2] for i in range(args.steps)], t_cur, 1)
        else:
            h_in, h_out = hin[0], hout[0]
            for _ in range(args.steps):
                prob.set_state(h_in, t_cur)  # H2D + ARKodeReset
                prob.step(1)
                prob.get_state(h_out)  # D2H (synchronises)
                h_in, h_out = h_out, h_in
                t_cur += H_FIXED
        barrier()
        dt = time.perf_counter() - t0
        e_s1 = prob.stats()
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e_evals = e_s1["rhs_evals"] - e_s0["rhs_evals"]
        e2e = {"value": e_evals * ncell_global / float(tt.item()), "unit": "cell-updates/s",
               "h2d_bytes_per_step": 8 * ncell_local * world, "d2h_bytes_per_step": 8 * ncell_local * world,
               "ms_per_step": 1e3 * float(tt.item()) / args.steps, "mode": args.e2e_mode,
               "host_placement_rank0": placement,
               "pcie_gbs_per_gpu_per_direction": 8 * ncell_local / 1e9 / (float(tt.item()) / args.steps),
               "note": ("independent batches, double-buffered: copies of batches i-1 / i+1 overlap the integration of batch i"
                        if args.e2e_mode == "pipelined" else "dependent steps: upload, step, download, nothing overlaps")}
        del hin, hout

    stats_final = prob.stats()
    prob.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline (HBM bound) -----------------------------------------------------------------
    # achieved = modelled bytes of every kernel launched in the timed region (the library counts
    # 8 B x cells x full vectors read + written per launch, b200_algorithmic_bytes) / device time.
    # >= 97 % of those bytes belong to the dominant kernel, the temporally blocked stage kernel
    # k_chain_march<K>: 4 streamed reads + 2 writes per K cell-updates (48/K B per update), so the
    # figure is that kernel's achieved bandwidth (slightly under-stated by the small per-step ops).
    # The SURVEY 8(d) basis of 40 B per cell-update (one HBM pass per stage) is reported beside it:
    # on that basis temporal blocking exceeds the one-pass-per-stage ceiling.
    peak, peak_src = measured_peak()
    klib.b200_last_chain_kernel.restype = ctypes.c_char_p
    chain_kernel = (klib.b200_last_chain_kernel() or b"").decode() or "k_chain_march"
    chain_l = s1["chain_launches"] - s0["chain_launches"]
    chain_s = s1["chain_stages"] - s0["chain_stages"]
    depth = (chain_s / chain_l) if chain_l else 1.0
    achieved = alg_bytes / (ms_max * 1e-3) / 1e9
    one_pass = ALG_BYTES_PER_UPDATE * evals * ncell_local / (ms_max * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("chain_dram_bytes_per_launch_16384" if chain_l else "dram_bytes_per_launch_16384")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": ("%s<K=%d> (temporally blocked STS stages)" % (chain_kernel, round(depth))) if chain_l
                          else "k_stage_march<5,PAT5(S,V,V,C,V)>",
                "modelled_bytes_timed": int(alg_bytes), "bytes_per_cell_update": alg_bytes / max(evals * ncell_local, 1),
                "stages_per_launch": depth, "chain_launches_timed": chain_l, "rhs_evals_timed": evals,
                "one_pass_per_stage_basis": {"bytes_per_cell_update": ALG_BYTES_PER_UPDATE, "achieved": one_pass,
                                             "frac": one_pass / peak,
                                             "note": "40 B/update x updates/s: above 1.0 = faster than any kernel "
                                                     "that makes one HBM pass per stage can be"},
                "frac_of_nominal_8TBs": achieved / 8000.0}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        ranks = 1
        while ranks * 2 <= min(cores, 64):
            ranks *= 2
        try:
            sample_n, sample_steps = cpu_sample_plan(ranks)
            v, t, ev = run_reference_sample(sample_steps, sample_n, ranks)
            cpu_baseline = {"value": v, "unit": "cell-updates/s", "cores": ranks, "kind": "reference",
                            "sample": "%d^2 block of the workload (same dx, dy, h => 92 RKC stages), %d steps = %d RHS evals, "
                                      "%d shared-memory MPI ranks of the unmodified reference build, %.1f s"
                                      % (sample_n, sample_steps, ev, ranks, t)}
        except Exception as exc:  # the reference binary is test infrastructure; report, do not hide
            cpu_baseline = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference",
                            "sample": "unavailable: %s" % exc}

    line = {
        "metric": "cell-updates/s (RHS evals x cells / s), diffusion_2D 16384^2 RKC",
        "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "diffusion_2D %dx%d per GPU, LSRKStep %s, fixedstep 1e-4 (%d stages/step)"
                               % (n, n, args.method.upper(), stats_final["max_stages"]),
                   "global_grid": [n * npx, n * npy], "parallelism": "%dx%d blocks, NCCL halo exchange" % (npx, npy),
                   "l2_note": "per-stage working set %.1f GiB >> 126 MB L2 (inputs larger than L2, no flush needed)"
                              % (5 * 8 * ncell_local / 2**30),
                   "rhs_evals_timed": evals,
                   "arith": ("exact: every multiply / add rounded separately, bit-identical to the reference's baseline build"
                             if args.arith == "exact" else
                             "fma: multiply-adds of the chained stages contracted (agrees with the reference to rounding, <= 1e-10 bar)")},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
        "clocks": sampler.summary(),
        "stage_chain_depth": depth,
        "operand_ring": ("cp.async.bulk + mbarrier (one 512-byte copy per warp, operand and row)"
                         if b200.kernel_lib().b200_get_chain_bulk() and args.arith == "exact" and depth == 4.0
                         else "cp.async (16 bytes per thread, operand and row)"),
    }
    if per_rank is not None:
        line["per_rank"] = per_rank
    if step_ms is not None:
        line["step_ms_rank0_diagnostic"] = step_ms
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
