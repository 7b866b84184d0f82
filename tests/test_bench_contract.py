"""bench.py's host-side contract, on CPU: the reference arm (`--impl reference` needs no GPU: it times the
unmodified reference build oracle/_ref/diffusion_2D_ref on the host cores) must print exactly one JSON line
with the keys the driver reads, and the workload helpers must keep dx, dy -- hence rho and the 92 stages --
fixed under weak scaling."""
import importlib.util
import json
import math
import os
import subprocess
import sys

import pytest
from conftest import ROOT

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "diffusion_2D_ref")


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/diffusion_2D_ref not built")
def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, B200_BENCH_REF_SAMPLE_N="256")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["dtype"] == "f64"
    assert "16384" in d["metric"] and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "256^2" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/diffusion_2D_ref not built")
def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", B200_BENCH_REF_SAMPLE_N="256")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_weak_scaling_workload_keeps_the_grid_spacing():
    b = _bench()
    assert [b.dims_create(n) for n in (1, 2, 4, 8)] == [(1, 1), (2, 1), (2, 2), (4, 2)]  # MPI_Dims_create
    n = 16384
    dx0 = (b.XU0 - b.XL) / (n - 1)
    for world in (1, 2, 4, 8):
        npx, npy = b.dims_create(world)
        a = b.workload_args(n, npx, npy, base_n=n)
        get = lambda f: a[a.index(f) + 1]  # noqa: E731
        nx, ny = int(get("--nx")), int(get("--ny"))
        assert (nx, ny) == (n * npx, n * npy)
        dx = (float(get("--xu")) - b.XL) / (nx - 1)
        dy = (float(get("--yu")) - b.YL) / (ny - 1)
        assert dx == pytest.approx(dx0, rel=1e-14) and dy == pytest.approx((b.YU0 - b.YL) / (n - 1), rel=1e-14)
        # rho = 1.01 * 8 / dx^2 (dx < dy), s = ceil(sqrt(1.54 h rho)) = 92 (arkode_lsrkstep.c:563-565)
        rho = 1.01 * 8.0 / (dx * dx)
        assert math.ceil(math.sqrt(1.54 * b.H_FIXED * rho)) == 92


def test_cpu_sample_plan_is_bounded():
    b = _bench()
    for ranks in (1, 2, 4, 8, 16, 32, 64):
        n, steps = b.cpu_sample_plan(ranks)
        assert n in (2048, 4096) and 1 <= steps <= 40
        assert 5.0 <= 93.0 * n * n * steps / (7.5e7 * ranks) <= 30.0  # ~10 s of CPU work by the measured rate


@pytest.mark.parametrize("fail_pin", ["0", "1"], ids=["pinned_ok", "pinned_alloc_fails"])
def test_measured_arm_control_flow_and_contract_keys(fail_pin):
    """bench.py's measured arm needs a B200; its host-side control flow does not.  tests/mock_bench_main.py runs
    main() with torch.cuda and the package mocked: the line must carry every key of the contract, and a failing
    pinned allocation must turn into an e2e error entry, not into an exception or a rank that stops taking part."""
    env = dict(os.environ, MOCK_FAIL_PIN=fail_pin)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mock_bench_main.py"), "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in d, key
    assert d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and "arith" in d["config"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"):
        assert key in d["roofline"], key
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["unit"] == "GB/s"
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in d["cpu_baseline"], key
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in d["e2e"], key
    if fail_pin == "1":
        assert d["e2e"]["value"] is None and "error" in d["e2e"]
    else:
        assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 8 * 16384 * 16384
