"""Shared test plumbing.

`-m "not gpu"` (run in the build container, no GPU): oracle vs golden vectors, host logic,
C-ABI symbol checks, world_size-2 gloo tests.  `-m gpu` (run on a B200): parity tests proper,
all through the C-ABI / the drop-in N_Vector + callbacks.

oracle/ is test infrastructure: it is loaded only from here (and smoke() / bench.py's
cpu_baseline), never from the product.
"""
import ctypes
import importlib
import json
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped, not failed, on a box without a CUDA device (plain `pytest tests`)."""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on a B200 with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _ensure_built():
    """The CPU suite needs the C oracle and the host libraries; build them if missing."""
    need = [os.path.join(ROOT, "oracle", "liboracle_sts.so"),
            os.path.join(ROOT, "ceda-demonstrations_b200", "lib", "libb200sts.so")]
    if not all(os.path.exists(p) for p in need):
        import subprocess

        subprocess.run(["make", "-s", "-C", ROOT, "product"], check=True)
        subprocess.run(["make", "-s", "-C", ROOT, "oracle"], check=True)


class OrcGrid(ctypes.Structure):
    _fields_ = [("kx", ctypes.c_double), ("ky", ctypes.c_double), ("inhomogeneous", ctypes.c_int),
                ("xl", ctypes.c_double), ("yl", ctypes.c_double), ("dx", ctypes.c_double), ("dy", ctypes.c_double),
                ("nx_loc", ctypes.c_int64), ("ny_loc", ctypes.c_int64), ("is_", ctypes.c_int64), ("js", ctypes.c_int64)]


class OrcAdr(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int64), ("ny", ctypes.c_int64)] + [(n, ctypes.c_double) for n in
                ("dx", "dy", "cux", "cuy", "cvx", "cvy", "d", "A", "B")]


class OrcStepWs(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("nglobal", ctypes.c_int64)] + [(n, ctypes.c_void_p) for n in
                ("yn", "fn", "ycur", "tempv1", "tempv2", "tempv3", "ewt")] + [("fixedstep", ctypes.c_int), ("nfe", ctypes.c_long)]


RHS_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                          ctypes.POINTER(ctypes.c_double), ctypes.c_void_p)
ATIMES_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def make_grid(nx, ny, kx=1.0, ky=1.0, inhom=False, xl=-math.pi, xu=math.pi, yl=-6.0, yu=6.0,
              nx_loc=None, ny_loc=None, is_=0, js=0):
    """orc_grid for the default domain (diffusion_2D.hpp:79-93); dx=(xu-xl)/(nx-1)."""
    return OrcGrid(kx, ky, int(inhom), xl, yl, (xu - xl) / (nx - 1), (yu - yl) / (ny - 1),
                   nx if nx_loc is None else nx_loc, ny if ny_loc is None else ny_loc, is_, js)


@pytest.fixture(scope="session")
def orc():
    _ensure_built()
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle_sts.so"))
    for name in ("orc_dot", "orc_maxnorm", "orc_wsqrsum", "orc_wrmsnorm", "orc_min", "orc_l1norm", "orc_coeff_x",
                 "orc_coeff_y", "orc_dom_eig", "orc_adr_domeig"):
        getattr(lib, name).restype = ctypes.c_double
    lib.orc_diffusion_fixed_run.restype = ctypes.c_long
    return lib


@pytest.fixture(scope="session")
def b200():
    _ensure_built()
    return importlib.import_module("ceda-demonstrations_b200")


def load_golden(name):
    with open(os.path.join(GOLDEN, "d2d_%s.json" % name)) as f:
        meta = json.load(f)
    state = np.load(os.path.join(GOLDEN, "d2d_%s.npy" % name))
    return meta, state


def fmt16(a):
    """The reference writes states with 16 significant digits (diffusion_2D.cpp:760-761); compare
    on that basis: two doubles are 'equal to the reference's output' if they print identically."""
    return np.array(["%.15e" % v for v in np.asarray(a).ravel()])


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
